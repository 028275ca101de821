// rdoq_impl.cuh -- rate-distortion optimised quantisation of one TU, the per-TU body of rdoq.cu.
//
// Replaces TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1974-2520) with xGetCodedLevel (:2660), xGetICRate (:2725),
// xGetRateLast (:2815), getSigCtxInc (:2548), calcPatternSigCtx (:2521), getSigCoeffGroupCtxInc (:2872) for square TUs of
// 4:2:0 pictures without scaling lists, extended precision or Golomb-Rice adaptation (every BASELINE cfg).
//
// The reference walks the coefficients of a TU once, backwards, with the state of the level coder (context set, greater-1 /
// greater-2 counters, Rice parameter) carried from coefficient to coefficient, and sums IEEE doubles in that order; the order
// of the additions is part of the result.  What does NOT depend on that state is split off and done by all lanes of the TU's
// lane group: the scaled magnitudes, the cost of quantising to zero, the distortion of the two candidate levels, the position
// of the first non-zero magnitude, the signs, the sign-bit hiding of the coefficient groups (independent of one another) and the
// stores.  The group's leader runs the state machine on those precomputed terms:
//
//   phase A  rq_prepass      every lane   q, cost of level 0, distortion of levels max and max-1 per scan position; last position
//   phase B  rq_decide       leader       level decisions, coefficient-group zero-out, last-position search      (sequential)
//   phase C  rq_finish       every lane   signs, zeros above the chosen last position, partial absolute sum
//   phase D  rq_hide_signs   every lane   sign-bit hiding, one coefficient group per lane at a time
//
// The functions are __host__ __device__ so that tests/rdoq_emul.cpp can run the same lines lane by lane on the CPU where there is
// no GPU (test infrastructure; the product launches them from rdoq.cu only).  Every double operation that the reference performs
// as a separate multiplication / addition is written with RQ_MUL / RQ_ADD / RQ_SUB (__dmul_rn / __dadd_rn on the device: never
// contracted into a fused multiply-add).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "../../include/hmgpu.h"

#if defined(__CUDACC__)
#define RQ_HD __host__ __device__ __forceinline__
#else
#define RQ_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define RQ_MUL(a, b) __dmul_rn((a), (b))
#define RQ_ADD(a, b) __dadd_rn((a), (b))
#define RQ_SUB(a, b) __dadd_rn((a), -(b))
#define RQ_LD(p) __ldg(p)
#else
#define RQ_MUL(a, b) ((a) * (b))
#define RQ_ADD(a, b) ((a) + (b))
#define RQ_SUB(a, b) ((a) - (b))
#define RQ_LD(p) (*(p))
#endif

#define RQ_SIGN_BITS   32768      // one bypass bin, 15-bit fixed point (xGetIEPRate)
#define RQ_MAX_LEVEL   32767      // entropyCodingMaximum of a 15-bit dynamic range
#define RQ_DBL_MAX     1.7976931348623157e308

// ---- scan tables (TComRom.cpp:53-220: SCAN_GROUPED_4x4 per TU size and scan type, SCAN_UNGROUPED over the groups) -------------
// one table for all (scan type t = 0 diagonal / 1 horizontal / 2 vertical, size class s = log2 - 2):
//   positions of the coefficients:        tab[t * 1360 + {0, 16, 80, 336}[s] + k]           k < 4^log2
//   raster index of the k-th group:       tab[4080 + t * 85 + {0, 1, 5, 21}[s] + k]         k < 4^(log2 - 2)
#define RQ_SCAN_WORDS 4335
RQ_HD int rq_scan_base(int t, int s) { return t * 1360 + (s == 0 ? 0 : s == 1 ? 16 : s == 2 ? 80 : 336); }
RQ_HD int rq_cg_base(int t, int s) { return 4080 + t * 85 + (s == 0 ? 0 : s == 1 ? 1 : s == 2 ? 5 : 21); }

// (host side: the table is built once and travels with every call)
// k-th sample of a w x w block in scan order -> (x, y)
static inline void rq_scan_xy(int type, int w, int k, int* x, int* y)
{
  if (type == 1) { *y = k / w; *x = k % w; return; }
  if (type == 2) { *x = k / w; *y = k % w; return; }
  // up-right diagonal: walk the anti-diagonals, each from its lower-left end
  int d = 0, first = 0;
  for (;; d++)
  {
    const int lo = d < w ? 0 : d - w + 1, hi = d < w ? d : w - 1, len = hi - lo + 1;     // x runs lo..hi on diagonal d
    if (k < first + len) { *x = lo + (k - first); *y = d - *x; return; }
    first += len;
  }
}
static inline void rq_build_scan_table(uint16_t* tab)
{
  for (int t = 0; t < 3; t++)
    for (int s = 0; s < 4; s++)
    {
      const int n = 4 << s, g = n >> 2;
      uint16_t* sc = tab + rq_scan_base(t, s);
      uint16_t* cg = tab + rq_cg_base(t, s);
      for (int i = 0; i < g * g; i++)
      {
        int gx, gy;
        rq_scan_xy(t, g, i, &gx, &gy);
        cg[i] = (uint16_t)(gy * g + gx);
        for (int k = 0; k < 16; k++)
        {
          int x, y;
          rq_scan_xy(t, 4, k, &x, &y);
          sc[16 * i + k] = (uint16_t)((4 * gy + y) * n + 4 * gx + x);
        }
      }
    }
}

// ---- the workspace of one TU (shared memory on the device) -------------------------------------------------------------------
// indexed by scan position except lv (raster).  44 bytes per coefficient.
struct RqWork
{
  double*  cz;     // cost of level 0
  double*  cc;     // A: distortion of level max            B: cost of the chosen level (pdCostCoeff)
  double*  cs;     // A: distortion of level max - 1        B: cost of its significance flag (pdCostSig)
  int32_t* qv;     // A: scaled magnitude (lLevelDouble)    B: deltaU
  int32_t* ru;     // rateIncUp
  int32_t* rd;     // rateIncDown
  int32_t* sd;     // sigRateDelta
  int32_t* lv;     // levels, raster
};
#define RQ_WORK_BYTES_PER_COEF 44
RQ_HD RqWork rq_carve(void* base, int n_coef)
{
  RqWork w;
  w.cz = (double*)base; w.cc = w.cz + n_coef; w.cs = w.cc + n_coef;
  w.qv = (int32_t*)(w.cs + n_coef); w.ru = w.qv + n_coef; w.rd = w.ru + n_coef; w.sd = w.rd + n_coef; w.lv = w.sd + n_coef;
  return w;
}

RQ_HD int rq_quant_scale(int rem) { return rem == 0 ? 26214 : rem == 1 ? 23302 : rem == 2 ? 20560 : rem == 3 ? 18396 : rem == 4 ? 16384 : 14564; }   // g_quantScales
RQ_HD int rq_inv_quant_scale(int rem) { return rem == 0 ? 40 : rem == 1 ? 45 : rem == 2 ? 51 : rem == 3 ? 57 : rem == 4 ? 64 : 72; }                    // g_invQuantScales
RQ_HD int rq_abs(int v) { return v < 0 ? -v : v; }

// ---- phase A: lane `lane` of `lanes` -----------------------------------------------------------------------------------------
// returns the highest scan position this lane saw with a non-zero rounded magnitude (-1: none)
RQ_HD int rq_prepass(const hmgpu_rdoq_job& j, const uint16_t* scan, const int32_t* coef, RqWork w, int lane, int lanes)
{
  const int n_coef = 1 << (2 * j.log2_size), qbits = j.qbits, qscale = rq_quant_scale(j.qp_rem);
  const long long cap = 0x7fffffffLL - (1LL << (qbits - 1));
  const double es = j.err_scale;
  int last = -1;
  for (int sp = lane; sp < n_coef; sp += lanes)
  {
    const int pos = scan[sp];
    const long long wide = (long long)rq_abs(RQ_LD(coef + pos)) * qscale;
    const int q = (int)(wide < cap ? wide : cap);
    int max_lvl = (q + (1 << (qbits - 1))) >> qbits;
    if (max_lvl > RQ_MAX_LEVEL) max_lvl = RQ_MAX_LEVEL;
    const double e0 = (double)q;
    w.cz[sp] = RQ_MUL(RQ_MUL(e0, e0), es);
    w.qv[sp] = q;
    if (max_lvl > 0)
    {
      last = sp;                                                   // sp ascends: the lane's highest
      const double e1 = (double)(q - (int)((unsigned)max_lvl << qbits));
      w.cc[sp] = RQ_MUL(RQ_MUL(e1, e1), es);
      if (max_lvl > 1)
      {
        const double e2 = (double)(q - (int)((unsigned)(max_lvl - 1) << qbits));
        w.cs[sp] = RQ_MUL(RQ_MUL(e2, e2), es);
      }
    }
    w.lv[sp] = 0;                                                  // (every raster entry once: sp runs over all of them)
  }
  return last;
}

// ---- rates -----------------------------------------------------------------------------------------------------------------
struct RqCoder { int ctx_set, c1, c2, c1_idx, c2_idx, rice; };

// bits of coding the absolute level lvl after its significance flag (xGetICRate; 8 greater-1 flags and 1 greater-2 flag per group,
// Rice prefix of at most 3 before the escape, no limited prefix length)
RQ_HD int rq_level_rate(const hmgpu_rdoq_bits* eb, int lvl, int ctx_one, int ctx_abs, const RqCoder& c)
{
  if (lvl <= 0) return 0;
  const bool g1 = c.c1_idx < 8, g2 = g1 && c.c2_idx < 1;
  const int base = g1 ? (g2 ? 3 : 2) : 1;
  int rate = RQ_SIGN_BITS;
  if (lvl >= base)
  {
    unsigned sym = (unsigned)(lvl - base);
    const int rice = c.rice;
    if (sym < (3u << rice)) rate += (int)((sym >> rice) + 1 + rice) << 15;
    else
    {
      sym -= 3u << rice;
      int len = rice;
      while (sym >= (1u << len)) { sym -= 1u << len; len++; }
      rate += (3 + len + 1 - rice + len) << 15;
    }
    if (g1)
    {
      rate += RQ_LD(&eb->greater_one[ctx_one][1]);
      if (g2) rate += RQ_LD(&eb->level_abs[ctx_abs][1]);
    }
  }
  else if (lvl == 1) rate += RQ_LD(&eb->greater_one[ctx_one][0]);
  else rate += RQ_LD(&eb->greater_one[ctx_one][1]) + RQ_LD(&eb->level_abs[ctx_abs][0]);       // lvl == 2 below a base of 3
  return rate;
}

RQ_HD int rq_last_group(int v) { return v < 4 ? v : v < 6 ? 4 : v < 8 ? 5 : v < 12 ? 6 : v < 16 ? 7 : v < 24 ? 8 : 9; }          // g_uiGroupIdx

// lambda x bits of (x, y) as the last significant position (xGetRateLast)
RQ_HD double rq_last_cost(const hmgpu_rdoq_bits* eb, double lambda, int ch, int x, int y)
{
  const int cx = rq_last_group(x), cy = rq_last_group(y);
  double bits = (double)(RQ_LD(&eb->last_x[ch][cx]) + RQ_LD(&eb->last_y[ch][cy]));
  if (cx > 3) bits = RQ_ADD(bits, 32768.0 * ((cx - 2) >> 1));
  if (cy > 3) bits = RQ_ADD(bits, 32768.0 * ((cy - 2) >> 1));
  return RQ_MUL(lambda, bits);
}

// significant_coeff_flag context of raster position pos (getSigCtxInc), with the chroma table offset
RQ_HD int rq_sig_ctx(int pattern, int first_ctx, int pos, int log2, int ch)
{
  const int y = pos >> log2, x = pos - (y << log2);
  const int table = ch ? 28 : 0;
  if (pos == 0) return table;
  if (log2 == 2)
  {
    const unsigned long long map = 0x8877886654325410ULL;          // ctxIndMap4x4, a nibble per position
    return table + first_ctx + (int)((map >> (4 * pos)) & 15);
  }
  const int xs = x & 3, ys = y & 3;
  int cnt;
  if (pattern == 0) cnt = xs + ys >= 3 ? 0 : (xs + ys >= 1 ? 1 : 2);
  else if (pattern == 1) cnt = ys >= 2 ? 0 : (ys >= 1 ? 1 : 2);
  else if (pattern == 2) cnt = xs >= 2 ? 0 : (xs >= 1 ? 1 : 2);
  else cnt = 2;
  return table + first_ctx + ((ch == 0 && ((x | y) >> 2)) ? 3 : 0) + cnt;
}

// ---- phase B: the leader -----------------------------------------------------------------------------------------------------
// last_pos = highest scan position with a non-zero rounded magnitude (>= 0).  Returns iBestLastIdxP1: the levels of the scan
// positions below it stand (in w.lv, magnitudes), everything from it upwards is zero.
RQ_HD int rq_decide(const hmgpu_rdoq_job& j, const hmgpu_rdoq_bits* eb, const uint16_t* scan, const uint16_t* scan_cg, RqWork w, int last_pos)
{
  const int log2 = j.log2_size, n_coef = 1 << (2 * log2), g = 1 << (log2 - 2), n_cg = g * g;
  const int ch = j.channel, qbits = j.qbits, rice0 = j.go_rice_init;
  const double lambda = j.lambda;
  const int set0 = ch ? 4 : 0;
  const int first_ctx = log2 == 2 ? 0 : log2 == 3 ? 9 + ((ch == 0 && j.scan != 0) ? 6 : 0) : (ch == 0 ? 21 : 12);
  const int last_cg = last_pos >> 4;
  unsigned long long cg_mask = 0;                                  // coefficient groups marked significant, bit = raster index
  double cg_cost[64];                                              // cost of the group flag as decided (pdCostCoeffGroupSig)

  // everything above the last position costs its level-0 distortion, summed from the top as the reference does
  double base = 0.0;
  for (int sp = n_coef - 1; sp > last_pos; sp--) base = RQ_ADD(base, w.cz[sp]);
  double uncoded = base;

  RqCoder lc;
  lc.ctx_set = set0 + ((ch == 0 && last_cg > 0) ? 2 : 0); lc.c1 = 1; lc.c2 = 0; lc.c1_idx = 0; lc.c2_idx = 0; lc.rice = rice0;

  for (int cg = last_cg; cg >= 0; cg--)
  {
    const int blk = scan_cg[cg], gy = blk >> (log2 - 2), gx = blk - (gy << (log2 - 2));
    const int right = gx < g - 1 ? (int)((cg_mask >> (blk + 1)) & 1) : 0;
    const int below = gy < g - 1 ? (int)((cg_mask >> (blk + g)) & 1) : 0;
    const int pattern = n_cg > 1 ? right + 2 * below : 0;
    double s_sig = 0.0, s_sig_first = 0.0, s_coded = 0.0, s_uncoded = 0.0;
    int nz_above_first = 0, any = 0;
    for (int k = (cg == last_cg ? (last_pos & 15) : 15); k >= 0; k--)
    {
      const int sp = cg * 16 + k, pos = scan[sp];
      const int q = w.qv[sp];
      int max_lvl = (q + (1 << (qbits - 1))) >> qbits;
      if (max_lvl > RQ_MAX_LEVEL) max_lvl = RQ_MAX_LEVEL;
      const double c_zero = w.cz[sp];
      uncoded = RQ_ADD(uncoded, c_zero);
      const int ctx_one = 4 * lc.ctx_set + lc.c1, ctx_abs = lc.ctx_set + lc.c2;
      const bool is_last = sp == last_pos;
      int sig0 = 0, sig1 = 0;
      if (!is_last)
      {
        const int ctx_sig = rq_sig_ctx(pattern, first_ctx, pos, log2, ch);
        sig0 = RQ_LD(&eb->sig[ctx_sig][0]); sig1 = RQ_LD(&eb->sig[ctx_sig][1]);
      }
      // xGetCodedLevel: level 0 competes only below 3 and not at the last position; then max, then max - 1 (strict <)
      int best = 0;
      double c_best = RQ_DBL_MAX, c_sig_best = 0.0;
      if (!is_last && max_lvl < 3)
      {
        c_sig_best = RQ_MUL(lambda, (double)sig0);
        c_best = RQ_ADD(c_zero, c_sig_best);
      }
      if (max_lvl > 0)
      {
        const double c_sig_one = is_last ? 0.0 : RQ_MUL(lambda, (double)sig1);
        double c = RQ_ADD(RQ_ADD(w.cc[sp], RQ_MUL(lambda, (double)rq_level_rate(eb, max_lvl, ctx_one, ctx_abs, lc))), c_sig_one);
        if (c < c_best) { best = max_lvl; c_best = c; c_sig_best = c_sig_one; }
        if (max_lvl > 1)
        {
          c = RQ_ADD(RQ_ADD(w.cs[sp], RQ_MUL(lambda, (double)rq_level_rate(eb, max_lvl - 1, ctx_one, ctx_abs, lc))), c_sig_one);
          if (c < c_best) { best = max_lvl - 1; c_best = c; c_sig_best = c_sig_one; }
        }
      }
      w.cc[sp] = c_best;
      w.cs[sp] = c_sig_best;
      w.sd[sp] = sig1 - sig0;
      w.qv[sp] = (q - (int)((unsigned)best << qbits)) >> (qbits - 8);
      if (best > 0)
      {
        const int now = rq_level_rate(eb, best, ctx_one, ctx_abs, lc);
        w.ru[sp] = rq_level_rate(eb, best + 1, ctx_one, ctx_abs, lc) - now;
        w.rd[sp] = rq_level_rate(eb, best - 1, ctx_one, ctx_abs, lc) - now;
      }
      else { w.ru[sp] = RQ_LD(&eb->greater_one[ctx_one][0]); w.rd[sp] = 0; }
      w.lv[pos] = best;
      base = RQ_ADD(base, c_best);

      // the level coder after this coefficient
      const int base_lvl = lc.c1_idx < 8 ? (lc.c2_idx < 1 ? 3 : 2) : 1;
      if (best >= base_lvl && best > (3 << lc.rice) && lc.rice < 4) lc.rice++;
      if (best >= 1) lc.c1_idx++;
      if (best > 1) { lc.c1 = 0; if (lc.c2 < 2) lc.c2++; lc.c2_idx++; }
      else if (best == 1 && lc.c1 > 0 && lc.c1 < 3) lc.c1++;
      if (k == 0 && sp > 0)
      {
        lc.ctx_set = set0 + ((ch == 0 && cg > 1) ? 2 : 0) + (lc.c1 == 0);
        lc.c1 = 1; lc.c2 = 0; lc.c1_idx = 0; lc.c2_idx = 0; lc.rice = rice0;
      }

      s_sig = RQ_ADD(s_sig, c_sig_best);
      if (k == 0) s_sig_first = c_sig_best;
      if (best)
      {
        any = 1;
        s_coded = RQ_ADD(s_coded, RQ_SUB(c_best, c_sig_best));
        s_uncoded = RQ_ADD(s_uncoded, c_zero);
        if (k != 0) nz_above_first++;
      }
    }
    cg_cost[cg] = 0.0;
    if (any) cg_mask |= 1ULL << blk;
    if (cg == 0) { cg_mask |= 1ULL; continue; }                    // the flag of the DC group is inferred
    const int ctx_grp = (right | below) != 0;                      // (the neighbours were decided before this group)
    if (!any)
    {
      const double flag0 = RQ_MUL(lambda, (double)RQ_LD(&eb->sig_group[ctx_grp][0]));
      base = RQ_ADD(base, RQ_SUB(flag0, s_sig));
      cg_cost[cg] = flag0;
    }
    else if (cg < last_cg)                                         // (the group of the last position is settled with that position)
    {
      if (nz_above_first == 0) { base = RQ_SUB(base, s_sig_first); s_sig = RQ_SUB(s_sig, s_sig_first); }
      const double flag0 = RQ_MUL(lambda, (double)RQ_LD(&eb->sig_group[ctx_grp][0]));
      const double flag1 = RQ_MUL(lambda, (double)RQ_LD(&eb->sig_group[ctx_grp][1]));
      double zeroed = RQ_ADD(base, flag0);
      base = RQ_ADD(base, flag1);
      cg_cost[cg] = flag1;
      zeroed = RQ_ADD(zeroed, s_uncoded);
      zeroed = RQ_SUB(zeroed, s_coded);
      zeroed = RQ_SUB(zeroed, s_sig);
      if (zeroed < base)
      {
        cg_mask &= ~(1ULL << blk);
        base = zeroed;
        cg_cost[cg] = flag0;
        for (int k = 15; k >= 0; k--)
        {
          const int sp = cg * 16 + k, pos = scan[sp];
          if (w.lv[pos]) { w.lv[pos] = 0; w.cc[sp] = w.cz[sp]; w.cs[sp] = 0.0; }
        }
      }
    }
  }

  // where to put the last significant position, or nothing coded at all
  double best_cost = RQ_ADD(uncoded, RQ_MUL(lambda, (double)j.cbf_bits[0]));
  base = RQ_ADD(base, RQ_MUL(lambda, (double)j.cbf_bits[1]));
  int best_end = 0;
  bool stop = false;
  for (int cg = last_cg; cg >= 0 && !stop; cg--)
  {
    base = RQ_SUB(base, cg_cost[cg]);
    if (!((cg_mask >> scan_cg[cg]) & 1)) continue;
    for (int k = (cg == last_cg ? (last_pos & 15) : 15); k >= 0; k--)
    {
      const int sp = cg * 16 + k, pos = scan[sp];
      const int l = w.lv[pos];
      if (l)
      {
        const int y = pos >> log2, x = pos - (y << log2);
        const double c_last = j.scan == 2 ? rq_last_cost(eb, lambda, ch, y, x) : rq_last_cost(eb, lambda, ch, x, y);
        const double total = RQ_SUB(RQ_ADD(base, c_last), w.cs[sp]);
        if (total < best_cost) { best_end = sp + 1; best_cost = total; }
        if (l > 1) { stop = true; break; }
        base = RQ_SUB(base, w.cc[sp]);
        base = RQ_ADD(base, w.cz[sp]);
      }
      else base = RQ_SUB(base, w.cs[sp]);
    }
  }
  return best_end;
}

// ---- phase C: signs below best_end, zeros from it up to last_pos; returns the lane's part of uiAbsSum ----------------------------
RQ_HD int rq_finish(const hmgpu_rdoq_job& j, const uint16_t* scan, const int32_t* coef, RqWork w, int best_end, int last_pos, int lane, int lanes)
{
  int sum = 0;
  for (int sp = lane; sp <= last_pos; sp += lanes)
  {
    const int pos = scan[sp];
    if (sp < best_end)
    {
      const int l = w.lv[pos];
      sum += l;
      if (RQ_LD(coef + pos) < 0) w.lv[pos] = -l;
    }
    else w.lv[pos] = 0;
  }
  (void)j;
  return sum;
}

// ---- phase D: sign-bit hiding (TComTrQuant.cpp:2380-2517), coefficient groups cg = lane, lane + lanes, ... ------------------------
// The parity of the sum of a group's levels must tell the sign of its first non-zero coefficient when its first and last non-zero
// are at least 4 scan positions apart; where it does not, the cheapest +-1 change in the group is applied.  Groups are independent
// but for one rule: in the first non-empty group from the top (`top_cg`) candidates above its last non-zero are not considered.
RQ_HD void rq_hide_signs(const hmgpu_rdoq_job& j, const uint16_t* scan, const int32_t* coef, RqWork w, int best_end, int lane, int lanes)
{
  const int top_cg = (best_end - 1) >> 4;
  const double inv = (double)rq_inv_quant_scale(j.qp_rem);
  // (Int64)(invQuant^2 * 2^(2 per) / lambda / 16 / 2^(2 (bitDepth - 8)) + 0.5), the reference's order of operations
  const double f = (RQ_MUL(RQ_MUL(inv, inv), (double)(1 << (2 * j.qp_per))) / j.lambda) / 16.0 / (double)(1 << (2 * (j.bit_depth - 8)));
  const long long rd_factor = (long long)RQ_ADD(f, 0.5);
  for (int cg = lane; cg <= top_cg; cg += lanes)
  {
    const uint16_t* s = scan + cg * 16;
    int first_nz = 16, last_nz = -1, sum = 0;
    for (int k = 0; k < 16; k++)
    {
      const int l = w.lv[s[k]];
      if (l) { if (first_nz == 16) first_nz = k; last_nz = k; }
    }
    if (last_nz - first_nz < 4) continue;
    for (int k = first_nz; k <= last_nz; k++) sum += w.lv[s[k]];
    const int sign = w.lv[s[first_nz]] > 0 ? 0 : 1;
    if (sign == (sum & 1)) continue;
    const bool top = cg == top_cg;
    long long min_cost = INT64_MAX;
    int min_pos = -1, final_change = 0;
    for (int k = top ? last_nz : 15; k >= 0; k--)
    {
      const int sp = cg * 16 + k, pos = s[k], l = w.lv[pos];
      long long cur;
      int change;
      if (l != 0)
      {
        const bool one = l == 1 || l == -1;
        const long long up = rd_factor * (long long)(-w.qv[sp]) + w.ru[sp];
        long long down = rd_factor * (long long)w.qv[sp] + w.rd[sp] - (one ? w.sd[sp] : 0);
        if (top && last_nz == k && one) down -= 4 << 15;
        if (up < down) { cur = up; change = 1; }
        else { change = -1; cur = (k == first_nz && one) ? INT64_MAX : down; }
      }
      else
      {
        cur = rd_factor * (long long)(-rq_abs(w.qv[sp])) + (1 << 15) + w.ru[sp] + w.sd[sp];
        change = 1;
        if (k < first_nz && (RQ_LD(coef + pos) >= 0 ? 0 : 1) != sign) cur = INT64_MAX;
      }
      if (cur < min_cost) { min_cost = cur; final_change = change; min_pos = pos; }
    }
    if (w.lv[min_pos] == RQ_MAX_LEVEL || w.lv[min_pos] == -RQ_MAX_LEVEL - 1) final_change = -1;
    if (RQ_LD(coef + min_pos) >= 0) w.lv[min_pos] += final_change; else w.lv[min_pos] -= final_change;
  }
}

// =================================================================================================================================
// One THREAD per TU (rdoq.cu, rdoq_tu_kernel): the 32 lanes of a warp carry 32 TUs of the same size through the reference's loop
// in LOCKSTEP -- every lane is at the same scan position at the same time, so the workspace of the warp, laid out
// [scan position][lane] in global memory, is read and written in full 128-byte lines -- and each lane runs the state machine of its
// own TU.  Where the lane-group kernel above keeps one lane of 32 busy during the sequential phase, this one keeps all of them.
//
// Per coefficient the workspace holds 24 bytes: the scaled magnitude with the coefficient's sign in bit 31 (qw), the decision
// (st: level, the level coder's state at that moment, the group's pattern), the cost of the decided level (cc) and of its
// significance flag (cs).  Everything else the reference stores per coefficient (cost of level 0, deltaU, rateIncUp / Down,
// sigRateDelta) is a function of those and is recomputed where it is needed: the same operations on the same operands, so the
// same bits.
//
// The loops have warp-uniform bounds (RQ_WARP_MAX over the lanes) and lane predicates inside.  On the host (tests/rdoq_emul.cpp)
// one lane runs alone; rq_ghost_top >= 0 stands for a neighbouring lane that is longer in every respect, so that the predicates
// are exercised too.
// =================================================================================================================================
#if defined(__CUDA_ARCH__)
#define RQ_WARP_MAX(v) __reduce_max_sync(0xffffffffu, (v))
#define RQ_WARP_ANY(p) (__any_sync(0xffffffffu, (p)) != 0)
#define RQ_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" :: "l"(p))
#define RQ_UNROLL _Pragma("unroll")
#define RQ2_STRIDE 32
#else
#define RQ_PREFETCH(p) ((void)(p))
#define RQ_UNROLL
#ifndef RQ_WARP_MAX      // (tests/rdoq_emul_warp.cpp brings its own: 32 host threads as the lanes of a warp, the device's strides)
static int rq_ghost_top = -1;
#define RQ_WARP_MAX(v) ((v) > rq_ghost_top ? (v) : rq_ghost_top)
#define RQ_WARP_ANY(p) ((p) || rq_ghost_top >= 0)
#define RQ2_STRIDE 1
#define RQ2_EB_STRIDE 1
#endif
#endif
#define RQ2_BYTES_PER_COEF 24

// The bit estimates as this half reads them: the 192 words of hmgpu_rdoq_bits, word i of the lane's set at p[i * RQ2_EB_STRIDE].
// On the device the warp keeps the 32 sets of its TUs in shared memory as [word][lane]: whatever word a lane asks for, it sits in
// the lane's own bank.
#if defined(__CUDA_ARCH__)
#define RQ2_EB_STRIDE 32
#endif
struct Rq2Bits { const int32_t* p; };
#define RQ2_EB(eb, i) ((eb).p[(i) * RQ2_EB_STRIDE])
enum { RQ2_SIG_GROUP = 0, RQ2_SIG = 4, RQ2_LAST_X = 92, RQ2_LAST_Y = 112, RQ2_GREATER_ONE = 132, RQ2_LEVEL_ABS = 180, RQ2_BITS_WORDS = 192 };
static_assert(sizeof(hmgpu_rdoq_bits) == 4 * RQ2_BITS_WORDS && offsetof(hmgpu_rdoq_bits, sig) == 4 * RQ2_SIG && offsetof(hmgpu_rdoq_bits, last_x) == 4 * RQ2_LAST_X
              && offsetof(hmgpu_rdoq_bits, last_y) == 4 * RQ2_LAST_Y && offsetof(hmgpu_rdoq_bits, greater_one) == 4 * RQ2_GREATER_ONE
              && offsetof(hmgpu_rdoq_bits, level_abs) == 4 * RQ2_LEVEL_ABS, "layout of hmgpu_rdoq_bits");

// lambda x bits of (x, y) as the last significant position (xGetRateLast; rq_last_cost on this half's bit estimates)
RQ_HD double rq2_last_cost(Rq2Bits eb, double lambda, int ch, int x, int y)
{
  const int cx = rq_last_group(x), cy = rq_last_group(y);
  double bits = (double)(RQ2_EB(eb, RQ2_LAST_X + 10 * ch + cx) + RQ2_EB(eb, RQ2_LAST_Y + 10 * ch + cy));
  if (cx > 3) bits = RQ_ADD(bits, 32768.0 * ((cx - 2) >> 1));
  if (cy > 3) bits = RQ_ADD(bits, 32768.0 * ((cy - 2) >> 1));
  return RQ_MUL(lambda, bits);
}

struct Rq2Work { int32_t* qw; uint32_t* st; double* cc; double* cs; };      // this lane's element sp: [sp * RQ2_STRIDE]
// slot = the workspace of one warp (32 TUs of n_coef coefficients), lane = this thread's place in it
RQ_HD Rq2Work rq2_carve(void* slot, int n_coef, int lane)
{
  Rq2Work w;
  const size_t n = (size_t)n_coef * RQ2_STRIDE;
  w.cc = (double*)slot + lane; w.cs = w.cc + n;
  w.qw = (int32_t*)((double*)slot + 2 * n) + lane; w.st = (uint32_t*)(w.qw - lane + n) + lane;
  return w;
}

// st: bits 0..14 level (as decided: a group zeroed afterwards is known by its bit in cg_mask), 16..18 context set, 19..20 c1, 21..22 c2, 23 greater-1 flags left,
// 24 greater-2 flag left, 25..27 Rice parameter, 28..29 pattern of the group, 30 the last position
RQ_HD uint32_t rq2_pack(int best, const RqCoder& c, int pattern, bool is_last)
{
  return (uint32_t)best | ((uint32_t)c.ctx_set << 16) | ((uint32_t)c.c1 << 19) | ((uint32_t)c.c2 << 21) | ((uint32_t)(c.c1_idx < 8) << 23)
       | ((uint32_t)(c.c2_idx < 1) << 24) | ((uint32_t)c.rice << 25) | ((uint32_t)pattern << 28) | ((uint32_t)is_last << 30);
}
RQ_HD int rq2_level(uint32_t st) { return (int)(st & 0x7fffu); }
RQ_HD double rq2_cost_zero(int q, double es) { const double e0 = (double)q; return RQ_MUL(RQ_MUL(e0, e0), es); }
RQ_HD int rq2_first_ctx(const hmgpu_rdoq_job& j, int log2)
{
  return log2 == 2 ? 0 : log2 == 3 ? 9 + ((j.channel == 0 && j.scan != 0) ? 6 : 0) : (j.channel == 0 ? 21 : 12);
}

// ---- the rates without branches: the lanes of a warp carry different TUs, so every branch of the per-coefficient step that two
// lanes take differently is walked twice.  The bit estimates of the level coder's contexts are loaded once per coefficient and the
// rate of a level is arithmetic on them (same sums as rq_level_rate, the escape length in closed form).
#if defined(__CUDA_ARCH__)
#define RQ_CLZ(v) __clz((int)(v))
#else
#define RQ_CLZ(v) __builtin_clz((unsigned)(v))
#endif
struct RqRateCtx { int go0, go1, la0, la1, base_lvl, rice; bool g1, g2; };
RQ_HD RqRateCtx rq_rate_ctx(Rq2Bits eb, const RqCoder& c)
{
  RqRateCtx r;
  const int ctx_one = 4 * c.ctx_set + c.c1;
  r.g1 = c.c1_idx < 8; r.g2 = r.g1 && c.c2_idx < 1;
  r.base_lvl = r.g1 ? (r.g2 ? 3 : 2) : 1; r.rice = c.rice;
  r.go0 = RQ2_EB(eb, RQ2_GREATER_ONE + 2 * ctx_one); r.go1 = RQ2_EB(eb, RQ2_GREATER_ONE + 2 * ctx_one + 1);
  const int ctx_abs = r.g2 ? c.ctx_set + c.c2 : 0;               // (only read while no greater-2 flag was coded in the group: c2 = 0)
  r.la0 = RQ2_EB(eb, RQ2_LEVEL_ABS + 2 * ctx_abs); r.la1 = RQ2_EB(eb, RQ2_LEVEL_ABS + 2 * ctx_abs + 1);
  return r;
}
RQ_HD int rq_rate_alu(int lvl, const RqRateCtx& r)
{
  const unsigned sym = (unsigned)(lvl - r.base_lvl), thr = 3u << r.rice;
  const int prefix = (int)((sym >> r.rice) + 1 + r.rice);
  const int len = 31 - RQ_CLZ((sym - thr + (1u << r.rice)) | 1u);  // Exp-Golomb of order rice: 2^len <= sym - thr + 2^rice
  const int escape = 3 + len + 1 - r.rice + len;
  const int at_base = ((sym < thr ? prefix : escape) << 15) + (r.g1 ? r.go1 : 0) + (r.g2 ? r.la1 : 0);
  const int below = lvl == 1 ? r.go0 : r.go1 + r.la0;
  return lvl > 0 ? RQ_SIGN_BITS + (lvl >= r.base_lvl ? at_base : below) : 0;
}

// getSigCtxInc with the group's part taken out of the per-coefficient step: grp_off = table + first_ctx (+ 3 in a luma group other
// than the DC group); the count of the neighbouring groups' pattern comes from a table of 64 two-bit entries [pattern][ys][xs]
constexpr unsigned long long rq_cnt_lut(int half)
{
  unsigned long long v = 0;
  for (int i = 0; i < 32; i++)
  {
    const int idx = half * 32 + i, p = idx >> 4, ys = (idx >> 2) & 3, xs = idx & 3;
    const int c = p == 0 ? (xs + ys >= 3 ? 0 : (xs + ys >= 1 ? 1 : 2)) : p == 1 ? (ys >= 2 ? 0 : (ys >= 1 ? 1 : 2)) : p == 2 ? (xs >= 2 ? 0 : (xs >= 1 ? 1 : 2)) : 2;
    v |= (unsigned long long)c << (2 * i);
  }
  return v;
}
constexpr unsigned long long RQ_CNT_LUT0 = rq_cnt_lut(0), RQ_CNT_LUT1 = rq_cnt_lut(1);
RQ_HD int rq_sig_ctx_lut(int pattern, int table, int first_ctx, int grp_off, int pos, int log2)
{
  const int idx = pattern * 16 + ((pos >> log2) & 3) * 4 + (pos & 3);
  const int cnt = (int)(((idx & 32 ? RQ_CNT_LUT1 : RQ_CNT_LUT0) >> (2 * (idx & 31))) & 3);
  const int in4 = table + first_ctx + (int)((0x8877886654325410ULL >> (4 * (pos & 15))) & 15);      // ctxIndMap4x4
  return pos == 0 ? table : (log2 == 2 ? in4 : grp_off + cnt);
}
RQ_HD int rq_sig_grp_off(int table, int first_ctx, int ch, int blk) { return table + first_ctx + ((ch == 0 && blk != 0) ? 3 : 0); }

// what the reference noted down beside a decision for sign-bit hiding: deltaU, rateIncUp, rateIncDown, sigRateDelta
RQ_HD void rq2_side(const hmgpu_rdoq_job& j, Rq2Bits eb, uint32_t st, int q, int pos, int log2, int* d_u, int* r_up, int* r_down, int* sig_delta)
{
  const int best = (int)(st & 0x7fffu);
  RqCoder c;
  c.ctx_set = (st >> 16) & 7; c.c1 = (st >> 19) & 3; c.c2 = (st >> 21) & 3;
  c.c1_idx = ((st >> 23) & 1) ? 0 : 8; c.c2_idx = ((st >> 24) & 1) ? 0 : 1; c.rice = (st >> 25) & 7;
  const RqRateCtx r = rq_rate_ctx(eb, c);
  *d_u = (q - (int)((unsigned)best << j.qbits)) >> (j.qbits - 8);
  const int now = rq_rate_alu(best, r);
  *r_up = best > 0 ? rq_rate_alu(best + 1, r) - now : r.go0;
  *r_down = best > 0 ? rq_rate_alu(best - 1, r) - now : 0;
  const int table = j.channel ? 28 : 0, first_ctx = rq2_first_ctx(j, log2);
  const int not_dc = (pos >> (log2 + 2)) | ((pos & ((1 << log2) - 1)) >> 2);     // the group is not the DC group
  const int ctx_sig = rq_sig_ctx_lut((st >> 28) & 3, table, first_ctx, rq_sig_grp_off(table, first_ctx, j.channel, not_dc), pos, log2);
  *sig_delta = ((st >> 30) & 1) ? 0 : RQ2_EB(eb, RQ2_SIG + 2 * ctx_sig + 1) - RQ2_EB(eb, RQ2_SIG + 2 * ctx_sig);
}

// The whole of xRateDistOptQuant for this lane's TU.  has_tu: the lane carries a TU (the last warp of a size class may not be
// full); log2: the size class, the same for every lane of the warp.  Levels that come out zero are NOT stored: the caller's level
// buffer starts out as zeros.  Returns uiAbsSum.
RQ_HD int rq2_tu(const hmgpu_rdoq_job& j, bool has_tu, int log2, Rq2Bits eb, const uint16_t* scan, const uint16_t* scan_cg,
                 const int32_t* coef, int32_t* level, Rq2Work w)
{
  const int n_coef = 1 << (2 * log2), g = 1 << (log2 - 2), n_cg = g * g;
  const int qbits = j.qbits, half = 1 << (qbits - 1), ch = j.channel;
  const double lambda = j.lambda, es = j.err_scale;

  // A: scaled magnitudes, the highest position whose rounded magnitude is not zero
  int last_pos = -1;
  if (has_tu)
  {
    const int qscale = rq_quant_scale(j.qp_rem);
    const long long cap = 0x7fffffffLL - (1LL << (qbits - 1));
#if defined(__CUDA_ARCH__)
#pragma unroll 16
#endif
    for (int sp = 0; sp < n_coef; sp++)
    {
      const int c = RQ_LD(coef + RQ_LD(scan + sp));
      const long long wide = (long long)rq_abs(c) * qscale;
      const int q = (int)(wide < cap ? wide : cap);
      if (((q + half) >> qbits) > 0) last_pos = sp;
      w.qw[(size_t)sp * RQ2_STRIDE] = (int32_t)((uint32_t)q | (c < 0 ? 0x80000000u : 0u));
    }
  }
  const bool live = last_pos >= 0;
  if (!RQ_WARP_ANY(live)) return 0;

  // B: level decisions and group zero-out, from the top of the scan down
  const int last_cg = last_pos >> 4;                               // (-1 for a lane that is not live)
  const int rice0 = j.go_rice_init, set0 = ch ? 4 : 0, first_ctx = rq2_first_ctx(j, log2), table = ch ? 28 : 0;
  unsigned long long cg_mask = 0;
  double cg_cost[64];
  double base = 0.0, uncoded = 0.0;
  int sum_all = 0;                                                 // sum of the levels as they stand (zeroed groups taken out again)
  RqCoder lc;
  lc.ctx_set = set0 + ((ch == 0 && last_cg > 0) ? 2 : 0); lc.c1 = 1; lc.c2 = 0; lc.c1_idx = 0; lc.c2_idx = 0; lc.rice = rice0;
  int q_next = live ? w.qw[(size_t)(n_coef - 1) * RQ2_STRIDE] : 0;
  for (int cg = n_cg - 1; cg >= 0; cg--)
  {
    const bool in = live && cg <= last_cg;
    int blk = 0, right = 0, below = 0, pattern = 0, grp_off = 0;
    if (in)
    {
      blk = RQ_LD(scan_cg + cg);
      const int gy = blk >> (log2 - 2), gx = blk - (gy << (log2 - 2));
      right = gx < g - 1 ? (int)((cg_mask >> (blk + 1)) & 1) : 0;
      below = gy < g - 1 ? (int)((cg_mask >> (blk + g)) & 1) : 0;
      pattern = n_cg > 1 ? right + 2 * below : 0;
      grp_off = rq_sig_grp_off(table, first_ctx, ch, blk);
    }
    double s_sig = 0.0, s_sig_first = 0.0, s_coded = 0.0, s_uncoded = 0.0;
    int nz_above_first = 0, any = 0, grp_sum = 0;
    for (int k = 15; k >= 0; k--)
    {
      const int sp = cg * 16 + k;
      if (!live) continue;
      const int q = q_next & 0x7fffffff;
      if (sp > 0) q_next = w.qw[(size_t)(sp - 1) * RQ2_STRIDE];     // (one position ahead: the load is off the critical path)
      if (sp >= 12) RQ_PREFETCH(&w.qw[(size_t)(sp - 12) * RQ2_STRIDE]);
      const double c_zero = rq2_cost_zero(q, es);
      uncoded = RQ_ADD(uncoded, c_zero);
      if (sp > last_pos) { base = uncoded; continue; }             // above the last position only the cost of level 0 adds up
      const int pos = RQ_LD(scan + sp);
      int max_lvl = (q + half) >> qbits;
      if (max_lvl > RQ_MAX_LEVEL) max_lvl = RQ_MAX_LEVEL;
      const bool is_last = sp == last_pos;
      // the bit estimates this coefficient can meet (independent loads), then both candidate levels side by side
      const RqRateCtx r = rq_rate_ctx(eb, lc);
      const int ctx_sig = rq_sig_ctx_lut(pattern, table, first_ctx, grp_off, pos, log2);
      const int sig0 = is_last ? 0 : RQ2_EB(eb, RQ2_SIG + 2 * ctx_sig), sig1 = is_last ? 0 : RQ2_EB(eb, RQ2_SIG + 2 * ctx_sig + 1);
      const double c_sig_zero = RQ_MUL(lambda, (double)sig0), c_sig_one = RQ_MUL(lambda, (double)sig1);    // (0.0 at the last position)
      const double e1 = (double)(q - (int)((unsigned)max_lvl << qbits)), e2 = (double)(q - (int)((unsigned)(max_lvl - 1) << qbits));
      const double c1 = RQ_ADD(RQ_ADD(RQ_MUL(RQ_MUL(e1, e1), es), RQ_MUL(lambda, (double)rq_rate_alu(max_lvl, r))), c_sig_one);
      const double c2 = RQ_ADD(RQ_ADD(RQ_MUL(RQ_MUL(e2, e2), es), RQ_MUL(lambda, (double)rq_rate_alu(max_lvl - 1, r))), c_sig_one);
      // xGetCodedLevel: level 0 competes only below 3 and not at the last position; then max, then max - 1 (strict <)
      int best = 0;
      double c_best = RQ_DBL_MAX, c_sig_best = 0.0;
      if (!is_last && max_lvl < 3) { c_sig_best = c_sig_zero; c_best = RQ_ADD(c_zero, c_sig_zero); }
      if (max_lvl > 0 && c1 < c_best) { best = max_lvl; c_best = c1; c_sig_best = c_sig_one; }
      if (max_lvl > 1 && c2 < c_best) { best = max_lvl - 1; c_best = c2; c_sig_best = c_sig_one; }
      w.cc[(size_t)sp * RQ2_STRIDE] = c_best;
      w.cs[(size_t)sp * RQ2_STRIDE] = c_sig_best;
      w.st[(size_t)sp * RQ2_STRIDE] = rq2_pack(best, lc, pattern, is_last);
      base = RQ_ADD(base, c_best);

      // the level coder after this coefficient
      const int base_lvl = lc.c1_idx < 8 ? (lc.c2_idx < 1 ? 3 : 2) : 1;
      if (best >= base_lvl && best > (3 << lc.rice) && lc.rice < 4) lc.rice++;
      if (best >= 1) lc.c1_idx++;
      if (best > 1) { lc.c1 = 0; if (lc.c2 < 2) lc.c2++; lc.c2_idx++; }
      else if (best == 1 && lc.c1 > 0 && lc.c1 < 3) lc.c1++;
      if (k == 0 && sp > 0)
      {
        lc.ctx_set = set0 + ((ch == 0 && cg > 1) ? 2 : 0) + (lc.c1 == 0);
        lc.c1 = 1; lc.c2 = 0; lc.c1_idx = 0; lc.c2_idx = 0; lc.rice = rice0;
      }

      s_sig = RQ_ADD(s_sig, c_sig_best);
      if (k == 0) s_sig_first = c_sig_best;
      if (best)
      {
        any = 1;
        grp_sum += best;
        s_coded = RQ_ADD(s_coded, RQ_SUB(c_best, c_sig_best));
        s_uncoded = RQ_ADD(s_uncoded, c_zero);
        if (k != 0) nz_above_first++;
      }
    }
    if (!in) continue;
    sum_all += grp_sum;
    cg_cost[cg] = 0.0;
    if (any) cg_mask |= 1ULL << blk;
    if (cg == 0) { cg_mask |= 1ULL; continue; }                    // the flag of the DC group is inferred
    const int ctx_grp = (right | below) != 0;
    if (!any)
    {
      const double flag0 = RQ_MUL(lambda, (double)RQ2_EB(eb, RQ2_SIG_GROUP + 2 * ctx_grp));
      base = RQ_ADD(base, RQ_SUB(flag0, s_sig));
      cg_cost[cg] = flag0;
    }
    else if (cg < last_cg)                                         // (the group of the last position is settled with that position)
    {
      if (nz_above_first == 0) { base = RQ_SUB(base, s_sig_first); s_sig = RQ_SUB(s_sig, s_sig_first); }
      const double flag0 = RQ_MUL(lambda, (double)RQ2_EB(eb, RQ2_SIG_GROUP + 2 * ctx_grp));
      const double flag1 = RQ_MUL(lambda, (double)RQ2_EB(eb, RQ2_SIG_GROUP + 2 * ctx_grp + 1));
      double zeroed = RQ_ADD(base, flag0);
      base = RQ_ADD(base, flag1);
      cg_cost[cg] = flag1;
      zeroed = RQ_ADD(zeroed, s_uncoded);
      zeroed = RQ_SUB(zeroed, s_coded);
      zeroed = RQ_SUB(zeroed, s_sig);
      if (zeroed < base)
      {
        cg_mask &= ~(1ULL << blk);
        base = zeroed;
        cg_cost[cg] = flag0;
        sum_all -= grp_sum;                                        // (nothing reads the decisions of a group that is not in cg_mask)
      }
    }
  }

  // where to put the last significant position, or nothing coded at all.  The walk goes down from the last position until the
  // first level above 1; the decisions of a group are fetched at once (the loads do not depend on the running cost).
  // uiAbsSum (before sign-bit hiding, as the reference returns it) = the levels as they stand minus those the walk passed before it
  // found its best position: every one of them was a 1.
  int best_end = 0, sum = 0;
  {
    double best_cost = RQ_ADD(uncoded, RQ_MUL(lambda, (double)j.cbf_bits[0]));
    base = RQ_ADD(base, RQ_MUL(lambda, (double)j.cbf_bits[1]));
    bool stop = !live;
    int walked = 0, above_best = 0;
    for (int cg = RQ_WARP_MAX(last_cg); cg >= 0; cg--)
    {
      if (!RQ_WARP_ANY(!stop)) break;
      bool act = !stop && cg <= last_cg;
      if (act) { base = RQ_SUB(base, cg_cost[cg]); act = ((cg_mask >> RQ_LD(scan_cg + cg)) & 1) != 0; }
      if (!RQ_WARP_ANY(act)) continue;
      const size_t at0 = (size_t)cg * 16 * RQ2_STRIDE;
      uint32_t stv[16];
      double csv[16];
      RQ_UNROLL
      for (int k = 0; k < 16; k++)
        if (act) { stv[k] = w.st[at0 + (size_t)k * RQ2_STRIDE]; csv[k] = w.cs[at0 + (size_t)k * RQ2_STRIDE]; }
        else { stv[k] = 0; csv[k] = 0.0; }
      RQ_UNROLL
      for (int k = 0; k < 16; k++)
        if (rq2_level(stv[k])) { RQ_PREFETCH(&w.cc[at0 + (size_t)k * RQ2_STRIDE]); RQ_PREFETCH(&w.qw[at0 + (size_t)k * RQ2_STRIDE]); }
      RQ_UNROLL
      for (int k = 15; k >= 0; k--)
      {
        const int sp = cg * 16 + k;
        if (!act || sp > last_pos) continue;
        const size_t at = (size_t)sp * RQ2_STRIDE;
        const int l = rq2_level(stv[k]);
        if (l)
        {
          const int pos = RQ_LD(scan + sp), y = pos >> log2, x = pos - (y << log2);
          const double c_last = j.scan == 2 ? rq2_last_cost(eb, lambda, ch, y, x) : rq2_last_cost(eb, lambda, ch, x, y);
          const double total = RQ_SUB(RQ_ADD(base, c_last), csv[k]);
          if (total < best_cost) { best_end = sp + 1; best_cost = total; above_best = walked; }
          if (l > 1) { stop = true; act = false; continue; }
          walked += l;
          base = RQ_SUB(base, w.cc[at]);
          base = RQ_ADD(base, rq2_cost_zero(w.qw[at] & 0x7fffffff, es));
        }
        else base = RQ_SUB(base, csv[k]);
      }
    }
    if (best_end) sum = sum_all - above_best;
  }

  // D: sign-bit hiding group by group (TComTrQuant.cpp:2380-2517), then the signed levels of the group go out
  const bool hide = (j.flags & HMGPU_RDOQ_SIGN_HIDE) && sum >= 2;
  const int top_cg = (best_end - 1) >> 4;                          // (-1: nothing coded)
  long long rd_factor = 0;
  if (hide)
  {
    const double inv = (double)rq_inv_quant_scale(j.qp_rem);
    const double f = (RQ_MUL(RQ_MUL(inv, inv), (double)(1 << (2 * j.qp_per))) / lambda) / 16.0 / (double)(1 << (2 * (j.bit_depth - 8)));
    rd_factor = (long long)RQ_ADD(f, 0.5);
  }
  const int top_top_cg = RQ_WARP_MAX(top_cg);
  for (int cg = 0; cg <= top_top_cg; cg++)
  {
    if (cg > top_cg || !((cg_mask >> RQ_LD(scan_cg + cg)) & 1)) continue;        // zeroed or empty groups: every level is zero
    const uint16_t* s = scan + cg * 16;
    const size_t at0 = (size_t)cg * 16 * RQ2_STRIDE;
    const int in_end = best_end - cg * 16;                         // positions k < in_end of this group are coded
    int lv[16], qwv[16];
    RQ_UNROLL
    for (int k = 0; k < 16; k++) { lv[k] = rq2_level(w.st[at0 + (size_t)k * RQ2_STRIDE]); qwv[k] = w.qw[at0 + (size_t)k * RQ2_STRIDE]; }
    if (cg < top_cg)                                               // the next group's lines, while this one is worked on
    {
      RQ_UNROLL
      for (int k = 0; k < 16; k++) { RQ_PREFETCH(&w.st[at0 + (size_t)(16 + k) * RQ2_STRIDE]); RQ_PREFETCH(&w.qw[at0 + (size_t)(16 + k) * RQ2_STRIDE]); }
    }
    RQ_UNROLL
    for (int k = 0; k < 16; k++) if (k >= in_end) lv[k] = 0;
    int chg_k = -1, chg = 0;
    if (hide)
    {
      int first_nz = 16, last_nz = -1, gsum = 0, sign = 0;
      RQ_UNROLL
      for (int k = 0; k < 16; k++)
        if (lv[k]) { if (first_nz == 16) { first_nz = k; sign = qwv[k] < 0 ? 1 : 0; } last_nz = k; gsum += lv[k]; }
      if (last_nz - first_nz >= 4 && sign != (gsum & 1))
      {
        const bool top = cg == top_cg;
        long long min_cost = INT64_MAX;
        bool chg_at_max = false;
        for (int k = top ? last_nz : 15; k >= 0; k--)             // (the group's lines are in L1 now)
        {
          const size_t at = at0 + (size_t)k * RQ2_STRIDE;
          const uint32_t st = w.st[at];
          const int qword = w.qw[at], l = k < in_end ? rq2_level(st) : 0;
          int d_u, r_up, r_down, sig_delta;
          rq2_side(j, eb, st, qword & 0x7fffffff, RQ_LD(s + k), log2, &d_u, &r_up, &r_down, &sig_delta);
          long long cur;
          int change;
          if (l != 0)
          {
            const bool one = l == 1;
            const long long up = rd_factor * (long long)(-d_u) + r_up;
            long long down = rd_factor * (long long)d_u + r_down - (one ? sig_delta : 0);
            if (top && last_nz == k && one) down -= 4 << 15;
            if (up < down) { cur = up; change = 1; }
            else { change = -1; cur = (k == first_nz && one) ? INT64_MAX : down; }
          }
          else
          {
            cur = rd_factor * (long long)(-rq_abs(d_u)) + (1 << 15) + r_up + sig_delta;
            change = 1;
            if (k < first_nz && ((qword < 0) ? 1 : 0) != sign) cur = INT64_MAX;
          }
          if (cur < min_cost) { min_cost = cur; chg = change; chg_k = k; chg_at_max = l == RQ_MAX_LEVEL && qword >= 0; }
        }
        if (chg_at_max) chg = -1;
      }
    }
    RQ_UNROLL
    for (int k = 0; k < 16; k++)
    {
      int l = lv[k];
      if (k == chg_k) l += chg;
      if (l) level[RQ_LD(s + k)] = qwv[k] < 0 ? -l : l;
    }
  }
  return sum;
}
