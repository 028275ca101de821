/* oracle/hm_rdoq.c -- TEST INFRASTRUCTURE ONLY (part of liboracle.so, see hm_oracle.c).
 *
 * CPU restatement of HM-16.2's rate-distortion optimised quantiser, TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1974-2520)
 * with its helpers xGetCodedLevel (:2660), xGetICRate (:2725), xGetRateLast (:2815), getSigCtxInc (:2548), calcPatternSigCtx
 * (:2521), getSigCoeffGroupCtxInc (:2872), the context-set rule getContextSetIndex (TComChromaFormat.h:243) and the scan tables of
 * initROM (TComRom.cpp:140-220), for the configurations the BASELINE cfgs use: square TUs of 4:2:0 pictures, no scaling lists,
 * no extended precision, no Golomb-Rice adaptation, sign-bit hiding on or off.
 *
 * PARITY: pinned.  tests/golden/rdoq_golden.npz holds calls of the reference's own function -- inputs, CABAC bit estimates and
 * returned levels -- dumped by the instrumented reference encoder oracle/_ref/TAppEncoderRdoq (oracle/Makefile target `rdoq`,
 * hooks in oracle/rdoq_dump.inc); tests/test_golden.py compares this file against every one of them and
 * tests/test_oracle_vs_ref.py against a fresh dump where the instrumented encoder exists.
 *
 * The costs are IEEE doubles and the decisions are comparisons of sums of them, so the ORDER of the floating-point operations
 * is part of the algorithm: every sum below is accumulated in the order the reference accumulates it (no re-association, no
 * fused multiply-add -- compile without -ffast-math / -ffp-contract=fast).
 */
#include "hm_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define RQ_MAX_COEF 1024
#define RQ_MAX_CG 64
#define RQ_SIGN_BIT_RATE 32768          /* one bypass bin in the 15-bit fixed point of the estimates (xGetIEPRate) */
#define RQ_C1_FLAGS 8                   /* greater-than-1 flags coded per coefficient group (C1FLAG_NUMBER) */
#define RQ_C2_FLAGS 1                   /* greater-than-2 flags coded per coefficient group (C2FLAG_NUMBER) */
#define RQ_RICE_ESCAPE 3                /* COEF_REMAIN_BIN_REDUCTION */
#define RQ_MAX_LEVEL 32767              /* entropyCodingMaximum for a 15-bit dynamic range */

static const int k_quant_scale[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };   /* g_quantScales, TComRom.cpp:321 */
static const int k_inv_quant_scale[6] = { 40, 45, 51, 57, 64, 72 };                 /* g_invQuantScales, TComRom.cpp:326 */
static const uint8_t k_last_group[32] = { 0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7,
                                          8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9 };   /* g_uiGroupIdx, TComRom.cpp:578 */
static const uint8_t k_sig_ctx_4x4[16] = { 0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8 }; /* ctxIndMap4x4, TComRom.cpp:569 */

/* ---- scan tables (TComRom.cpp:53-137 ScanGenerator, :140-220) ----------------------------------------------------------------
 * order[k] = raster position (pitch `pitch`) of the k-th sample of a bw x bh block whose top-left sample is (x0, y0) */
static void scan_block(int type, int bw, int bh, int x0, int y0, int pitch, uint16_t* order)
{
  int k = 0;
  if (type == 1)                                   /* horizontal: row by row */
    for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) order[k++] = (uint16_t)((y0 + y) * pitch + x0 + x);
  else if (type == 2)                              /* vertical: column by column */
    for (int x = 0; x < bw; x++) for (int y = 0; y < bh; y++) order[k++] = (uint16_t)((y0 + y) * pitch + x0 + x);
  else                                             /* diagonal: every anti-diagonal from its bottom-left end to its top-right end */
    for (int d = 0; d < bw + bh - 1; d++)
      for (int y = d < bh ? d : bh - 1; y >= 0 && d - y < bw; y--) order[k++] = (uint16_t)((y0 + y) * pitch + x0 + d - y);
}

void hmo_scan_order(int log2_size, int scan_type, uint16_t* scan, uint16_t* scan_cg)
{
  const int n = 1 << log2_size, g = n >> 2;
  scan_block(scan_type, g, g, 0, 0, g, scan_cg);                         /* the coefficient groups, SCAN_UNGROUPED over g x g */
  for (int i = 0; i < g * g; i++)                                        /* SCAN_GROUPED_4x4 */
    scan_block(scan_type, 4, 4, 4 * (scan_cg[i] % g), 4 * (scan_cg[i] / g), n, scan + 16 * i);
}

/* ---- contexts ---------------------------------------------------------------------------------------------------------------- */
/* first significance-map context of the TU (getTUEntropyCodingParameters, TComChromaFormat.cpp; ContextTables.h:85-87) */
static int first_sig_ctx(int log2_size, int channel, int scan_type)
{
  if (log2_size == 2) return 0;
  if (log2_size == 3) return 9 + ((channel == 0 && scan_type != 0) ? 6 : 0);
  return channel == 0 ? 21 : 12;
}

/* significant_coeff_flag context of raster position pos, without the luma / chroma table offset (getSigCtxInc, :2548) */
static int sig_ctx_inc(int pattern, int first_ctx, int pos, int log2_size, int channel)
{
  const int y = pos >> log2_size, x = pos - (y << log2_size);
  if (x + y == 0) return 0;
  if (log2_size == 2) return first_ctx + k_sig_ctx_4x4[4 * y + x];
  const int xs = x & 3, ys = y & 3;
  int cnt;
  if (pattern == 0) cnt = xs + ys >= 3 ? 0 : (xs + ys >= 1 ? 1 : 2);
  else if (pattern == 1) cnt = ys >= 2 ? 0 : (ys >= 1 ? 1 : 2);
  else if (pattern == 2) cnt = xs >= 2 ? 0 : (xs >= 1 ? 1 : 2);
  else cnt = 2;
  const int other_group = ((x >> 2) + (y >> 2)) > 0;
  return first_ctx + ((other_group && channel == 0) ? 3 : 0) + cnt;
}

/* right / below coefficient groups already marked significant */
static void cg_neighbours(const uint8_t* cg_sig, int gx, int gy, int g, int* right, int* below)
{
  *right = gx < g - 1 ? cg_sig[gy * g + gx + 1] != 0 : 0;
  *below = gy < g - 1 ? cg_sig[(gy + 1) * g + gx] != 0 : 0;
}

/* ---- rates -------------------------------------------------------------------------------------------------------------------- */
/* bits (15-bit fixed point) of coding absolute level `lvl` after its significance flag (xGetICRate, :2725; no limited prefix) */
static int level_rate(const hmo_rdoq_bits* eb, uint32_t lvl, int ctx_one, int ctx_abs, int rice, int c1_idx, int c2_idx)
{
  const uint32_t base = c1_idx < RQ_C1_FLAGS ? (2u + (c2_idx < RQ_C2_FLAGS)) : 1u;
  int rate = RQ_SIGN_BIT_RATE;
  if (lvl >= base)
  {
    uint32_t sym = lvl - base;
    if (sym < ((uint32_t)RQ_RICE_ESCAPE << rice)) rate += (int)((sym >> rice) + 1 + rice) << 15;
    else
    {
      int len = rice;
      sym -= (uint32_t)RQ_RICE_ESCAPE << rice;
      while (sym >= (1u << len)) sym -= 1u << (len++);
      rate += (RQ_RICE_ESCAPE + len + 1 - rice + len) << 15;
    }
    if (c1_idx < RQ_C1_FLAGS)
    {
      rate += eb->greater_one[ctx_one][1];
      if (c2_idx < RQ_C2_FLAGS) rate += eb->level_abs[ctx_abs][1];
    }
  }
  else if (lvl == 1) rate += eb->greater_one[ctx_one][0];
  else if (lvl == 2) rate += eb->greater_one[ctx_one][1] + eb->level_abs[ctx_abs][0];
  else rate = 0;
  return rate;
}

/* lambda-weighted bits of signalling (x, y) as the last significant position (xGetRateLast, :2815) */
static double last_pos_cost(const hmo_rdoq_bits* eb, double lambda, int channel, int x, int y)
{
  const int cx = k_last_group[x], cy = k_last_group[y];
  double bits = eb->last_x[channel][cx] + eb->last_y[channel][cy];
  if (cx > 3) bits += 32768.0 * ((cx - 2) >> 1);
  if (cy > 3) bits += 32768.0 * ((cy - 2) >> 1);
  return lambda * bits;
}

/* state of the greater-than-1 / greater-than-2 / Rice machinery while a coefficient group is walked backwards */
typedef struct { int ctx_set, c1, c2, c1_idx, c2_idx, rice; } LevelCtx;

int hmo_rdoq(const hmo_rdoq_tu* tu, const hmo_rdoq_bits* eb, const int32_t* coef, int32_t* level)
{
  const int log2 = tu->log2_size, n = 1 << log2, n_coef = n * n, g = n >> 2, n_cg = g * g;
  const int ch = tu->channel, qbits = tu->qbits;
  const double lambda = tu->lambda, err_scale = tu->err_scale;
  const int qscale = k_quant_scale[tu->qp_rem];
  const int sig_table = ch == 0 ? 0 : 28;                      /* getSignificanceMapContextOffset: chroma contexts follow luma's 28 */
  const int set_table = ch == 0 ? 0 : 4;                       /* contextSetStartTable */
  const int first_ctx = first_sig_ctx(log2, ch, tu->scan);
  static __thread uint16_t scan[RQ_MAX_COEF], scan_cg[RQ_MAX_CG];
  static __thread double cost_coded[RQ_MAX_COEF], cost_zero[RQ_MAX_COEF], cost_sig[RQ_MAX_COEF];
  static __thread int rate_up[RQ_MAX_COEF], rate_down[RQ_MAX_COEF], sig_delta[RQ_MAX_COEF], delta_u[RQ_MAX_COEF];
  double cg_sig_cost[RQ_MAX_CG];
  uint8_t cg_sig[RQ_MAX_CG];
  hmo_scan_order(log2, tu->scan, scan, scan_cg);
  memset(cost_coded, 0, sizeof(double) * (size_t)n_coef);
  memset(cost_sig, 0, sizeof(double) * (size_t)n_coef);
  memset(rate_up, 0, sizeof(int) * (size_t)n_coef);
  memset(rate_down, 0, sizeof(int) * (size_t)n_coef);
  memset(sig_delta, 0, sizeof(int) * (size_t)n_coef);
  memset(delta_u, 0, sizeof(int) * (size_t)n_coef);
  memset(cg_sig_cost, 0, sizeof cg_sig_cost);
  memset(cg_sig, 0, sizeof cg_sig);

  double uncoded_total = 0, base = 0;
  int last_pos = -1, last_cg = -1;
  LevelCtx lc = { 0, 1, 0, 0, 0, tu->go_rice_init };

  /* ---- pass 1, backwards over the coefficient groups: best level of every coefficient, then "drop the whole group?" ---- */
  for (int cg = n_cg - 1; cg >= 0; cg--)
  {
    const int cg_blk = scan_cg[cg], gy = cg_blk / g, gx = cg_blk - gy * g;
    int right, below;
    cg_neighbours(cg_sig, gx, gy, g, &right, &below);
    const int pattern = n_cg > 1 ? right + 2 * below : 0;       /* calcPatternSigCtx */
    double s_sig = 0, s_sig_first = 0, s_coded = 0, s_uncoded = 0;
    int nz_above_first = 0;
    for (int k = 15; k >= 0; k--)
    {
      const int sp = cg * 16 + k, pos = scan[sp];
      const int64_t wide = (int64_t)abs(coef[pos]) * qscale;
      const int64_t cap = (int64_t)0x7fffffff - ((int64_t)1 << (qbits - 1));
      const int32_t q = (int32_t)(wide < cap ? wide : cap);                       /* scaled magnitude, qbits fractional bits */
      uint32_t max_lvl = (uint32_t)((q + (1 << (qbits - 1))) >> qbits);
      if (max_lvl > RQ_MAX_LEVEL) max_lvl = RQ_MAX_LEVEL;
      const double e0 = (double)q;
      cost_zero[sp] = e0 * e0 * err_scale;
      uncoded_total += cost_zero[sp];
      level[pos] = (int32_t)max_lvl;
      if (max_lvl > 0 && last_pos < 0)
      {
        last_pos = sp;
        last_cg = cg;
        lc.ctx_set = set_table + ((ch == 0 && (sp >> 4) > 0) ? 2 : 0);
      }
      if (last_pos >= 0)
      {
        const int ctx_one = 4 * lc.ctx_set + lc.c1, ctx_abs = lc.ctx_set + lc.c2;
        const int is_last = sp == last_pos;
        int ctx_sig = sig_table;
        if (!is_last) ctx_sig += sig_ctx_inc(pattern, first_ctx, pos, log2, ch);
        /* xGetCodedLevel: the cheapest of {0 (only when max_lvl < 3 and not the last), max_lvl, max_lvl - 1} */
        uint32_t best = 0;
        double c_best, c_sig_best = cost_sig[sp], c_sig_one = 0;
        int decided = 0;
        if (!is_last && max_lvl < 3)
        {
          c_sig_best = lambda * (double)eb->sig[ctx_sig][0];
          c_best = cost_zero[sp] + c_sig_best;
          if (max_lvl == 0) decided = 1;
        }
        else c_best = DBL_MAX;
        if (!decided)
        {
          if (!is_last) c_sig_one = lambda * (double)eb->sig[ctx_sig][1];
          const uint32_t min_lvl = max_lvl > 1 ? max_lvl - 1 : 1;
          for (int l = (int)max_lvl; l >= (int)min_lvl; l--)
          {
            const double e = (double)(q - (int32_t)((uint32_t)l << qbits));
            double c = e * e * err_scale + lambda * (double)level_rate(eb, (uint32_t)l, ctx_one, ctx_abs, lc.rice, lc.c1_idx, lc.c2_idx);
            c += c_sig_one;
            if (c < c_best) { best = (uint32_t)l; c_best = c; c_sig_best = c_sig_one; }
          }
        }
        cost_coded[sp] = c_best;
        cost_sig[sp] = c_sig_best;
        if (!is_last) sig_delta[pos] = eb->sig[ctx_sig][1] - eb->sig[ctx_sig][0];
        delta_u[pos] = (q - (int32_t)(best << qbits)) >> (qbits - 8);
        if (best > 0)
        {
          const int now = level_rate(eb, best, ctx_one, ctx_abs, lc.rice, lc.c1_idx, lc.c2_idx);
          rate_up[pos] = level_rate(eb, best + 1, ctx_one, ctx_abs, lc.rice, lc.c1_idx, lc.c2_idx) - now;
          rate_down[pos] = level_rate(eb, best - 1, ctx_one, ctx_abs, lc.rice, lc.c1_idx, lc.c2_idx) - now;
        }
        else rate_up[pos] = eb->greater_one[ctx_one][0];
        level[pos] = (int32_t)best;
        base += cost_coded[sp];

        /* the coder's state after this coefficient */
        const uint32_t base_lvl = lc.c1_idx < RQ_C1_FLAGS ? (2u + (lc.c2_idx < RQ_C2_FLAGS)) : 1u;
        if (best >= base_lvl && best > (3u << lc.rice)) lc.rice = lc.rice + 1 < 4 ? lc.rice + 1 : 4;
        if (best >= 1) lc.c1_idx++;
        if (best > 1) { lc.c1 = 0; lc.c2 += lc.c2 < 2; lc.c2_idx++; }
        else if (lc.c1 < 3 && lc.c1 > 0 && best) lc.c1++;
        if (k == 0 && sp > 0)                                   /* entering the next group */
        {
          lc.ctx_set = set_table + ((ch == 0 && ((sp - 1) >> 4) > 0) ? 2 : 0) + (lc.c1 == 0);
          lc.c1 = 1; lc.c2 = 0; lc.c1_idx = 0; lc.c2_idx = 0; lc.rice = tu->go_rice_init;
        }
      }
      else base += cost_zero[sp];
      s_sig += cost_sig[sp];
      if (k == 0) s_sig_first = cost_sig[sp];
      if (level[pos])
      {
        cg_sig[cg_blk] = 1;
        s_coded += cost_coded[sp] - cost_sig[sp];
        s_uncoded += cost_zero[sp];
        if (k != 0) nz_above_first++;
      }
    }
    if (last_cg < 0) continue;
    if (cg == 0) { cg_sig[cg_blk] = 1; continue; }              /* the DC group's flag is inferred */
    cg_neighbours(cg_sig, gx, gy, g, &right, &below);
    const int ctx_grp = (right + below) != 0;                    /* getSigCoeffGroupCtxInc */
    if (!cg_sig[cg_blk])
    {
      base += lambda * (double)eb->sig_group[ctx_grp][0] - s_sig;
      cg_sig_cost[cg] = lambda * (double)eb->sig_group[ctx_grp][0];
    }
    else if (cg < last_cg)                                      /* (the group of the last position is settled with that position) */
    {
      if (nz_above_first == 0) { base -= s_sig_first; s_sig -= s_sig_first; }
      double zeroed = base;
      base += lambda * (double)eb->sig_group[ctx_grp][1];
      zeroed += lambda * (double)eb->sig_group[ctx_grp][0];
      cg_sig_cost[cg] = lambda * (double)eb->sig_group[ctx_grp][1];
      zeroed += s_uncoded;
      zeroed -= s_coded;
      zeroed -= s_sig;
      if (zeroed < base)
      {
        cg_sig[cg_blk] = 0;
        base = zeroed;
        cg_sig_cost[cg] = lambda * (double)eb->sig_group[ctx_grp][0];
        for (int k = 15; k >= 0; k--)
        {
          const int sp = cg * 16 + k, pos = scan[sp];
          if (level[pos]) { level[pos] = 0; cost_coded[sp] = cost_zero[sp]; cost_sig[sp] = 0; }
        }
      }
    }
  }
  if (last_pos < 0) return 0;

  /* ---- pass 2: where to put the last significant position (or code nothing at all) ---- */
  double best_cost = uncoded_total + lambda * (double)tu->cbf_bits[0];
  base += lambda * (double)tu->cbf_bits[1];
  int best_end = 0, stop = 0;
  for (int cg = last_cg; cg >= 0 && !stop; cg--)
  {
    base -= cg_sig_cost[cg];
    if (!cg_sig[scan_cg[cg]]) continue;
    for (int k = 15; k >= 0; k--)
    {
      const int sp = cg * 16 + k;
      if (sp > last_pos) continue;
      const int pos = scan[sp];
      if (level[pos])
      {
        const int y = pos >> log2, x = pos - (y << log2);
        const double c_last = tu->scan == 2 ? last_pos_cost(eb, lambda, ch, y, x) : last_pos_cost(eb, lambda, ch, x, y);
        const double total = base + c_last - cost_sig[sp];
        if (total < best_cost) { best_end = sp + 1; best_cost = total; }
        if (level[pos] > 1) { stop = 1; break; }
        base -= cost_coded[sp];
        base += cost_zero[sp];
      }
      else base -= cost_sig[sp];
    }
  }

  int abs_sum = 0;
  for (int sp = 0; sp < best_end; sp++)
  {
    const int pos = scan[sp], l = level[pos];
    abs_sum += l;
    level[pos] = coef[pos] < 0 ? -l : l;
  }
  for (int sp = best_end; sp <= last_pos; sp++) level[scan[sp]] = 0;

  /* ---- pass 3: sign-bit hiding -- make the parity of every eligible group agree with the sign of its first coefficient ---- */
  if (tu->sign_hide && abs_sum >= 2)
  {
    const double inv = (double)k_inv_quant_scale[tu->qp_rem];
    const int64_t rd_factor = (int64_t)(inv * inv * (1 << (2 * tu->qp_per)) / lambda / 16 / (1 << (2 * (tu->bit_depth - 8))) + 0.5);
    int first_group_seen = -1;                                   /* lastCG: -1 before the first non-empty group, 1 in it, 0 after */
    for (int cg = n_cg - 1; cg >= 0; cg--)
    {
      const uint16_t* s = scan + cg * 16;
      int first_nz = 16, last_nz = -1, sum = 0;
      for (int k = 15; k >= 0; k--) if (level[s[k]]) { last_nz = k; break; }
      for (int k = 0; k < 16; k++) if (level[s[k]]) { first_nz = k; break; }
      for (int k = first_nz; k <= last_nz; k++) sum += level[s[k]];
      if (last_nz >= 0 && first_group_seen == -1) first_group_seen = 1;
      if (last_nz - first_nz >= 4)
      {
        const int sign = level[s[first_nz]] > 0 ? 0 : 1;
        if (sign != (sum & 1))
        {
          int64_t min_cost = INT64_MAX, cur = INT64_MAX;
          int min_pos = -1, final_change = 0, change = 0;
          for (int k = first_group_seen == 1 ? last_nz : 15; k >= 0; k--)
          {
            const int pos = s[k];
            if (level[pos] != 0)
            {
              const int64_t up = rd_factor * (-delta_u[pos]) + rate_up[pos];
              int64_t down = rd_factor * delta_u[pos] + rate_down[pos] - (abs(level[pos]) == 1 ? sig_delta[pos] : 0);
              if (first_group_seen == 1 && last_nz == k && abs(level[pos]) == 1) down -= 4 << 15;
              if (up < down) { cur = up; change = 1; }
              else
              {
                change = -1;
                cur = (k == first_nz && abs(level[pos]) == 1) ? INT64_MAX : down;
              }
            }
            else
            {
              cur = rd_factor * (-(int64_t)abs(delta_u[pos])) + (1 << 15) + rate_up[pos] + sig_delta[pos];
              change = 1;
              if (k < first_nz && (coef[pos] >= 0 ? 0 : 1) != sign) cur = INT64_MAX;
            }
            if (cur < min_cost) { min_cost = cur; final_change = change; min_pos = pos; }
          }
          if (level[min_pos] == RQ_MAX_LEVEL || level[min_pos] == -RQ_MAX_LEVEL - 1) final_change = -1;
          if (coef[min_pos] >= 0) level[min_pos] += final_change; else level[min_pos] -= final_change;
        }
      }
      if (first_group_seen == 1) first_group_seen = 0;
    }
  }
  return abs_sum;
}


/* ---- xDeQuant (TComTrQuant.cpp:1203-1313), the branch without scaling lists (:1276-1311) ------------------------------------
   rightShift = IQUANT_SHIFT - (transform shift + per) with transform shift = 15 - bitDepth - log2 (getTransformShift,
   TComTrQuant.h); the level is clipped to the input range the 32-bit intermediate can carry (:1284-1286), multiplied by
   g_invQuantScales[rem], rounded and shifted right -- or shifted left when rightShift <= 0 -- and clipped to [-32768, 32767].
   Intermediate_Int is a 32-bit int in this build (TypeDef.h:696). */
void hmo_dequant(const int32_t* level, int n_coef, int log2_size, int qp_per, int qp_rem, int bit_depth, int32_t* coef)
{
  static const int inv_scale[6] = { 40, 45, 51, 57, 64, 72 };              /* g_invQuantScales (TComRom.cpp) */
  const int transform_shift = 15 - bit_depth - log2_size;
  const int right_shift = 6 - (transform_shift + qp_per);                 /* IQUANT_SHIFT = 6 */
  const int scale = inv_scale[qp_rem], scale_bits = 6 + 1;
  int target = 32 + right_shift - scale_bits;
  if (target > 16) target = 16;
  const int32_t in_min = -(1 << (target - 1)), in_max = (1 << (target - 1)) - 1;
  for (int i = 0; i < n_coef; i++)
  {
    int32_t q = level[i] < in_min ? in_min : level[i] > in_max ? in_max : level[i];
    int32_t v;
    if (right_shift > 0) v = (q * scale + (1 << (right_shift - 1))) >> right_shift;
    else v = (int32_t)((uint32_t)(q * scale) << -right_shift);
    coef[i] = v < -32768 ? -32768 : v > 32767 ? 32767 : v;
  }
}


/* many TUs in one call: TU i reads coef + coef_offset[i], writes level + coef_offset[i] and uses bits[bits_index[i]]; returns the
   sum of the uiAbsSum values.  (bench.py's cpu_baseline of the rdoq leg: the restatement timed without the binding's per-call cost) */
long long hmo_rdoq_batch(const hmo_rdoq_tu* tus, const int32_t* bits_index, const uint32_t* coef_offset, int n_tus,
                         const hmo_rdoq_bits* bits, const int32_t* coef, int32_t* level)
{
  long long total = 0;
  for (int i = 0; i < n_tus; i++)
    total += hmo_rdoq(&tus[i], &bits[bits_index[i]], coef + coef_offset[i], level + coef_offset[i]);
  return total;
}
