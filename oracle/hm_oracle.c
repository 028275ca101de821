/* oracle/hm_oracle.c
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C, CPU restatement of the data-parallel inter-search
 * hot path of the HM-16.2 HEVC reference encoder.  It is the checker
 * the CUDA path (libhmgpu) is compared against; nothing in the product path links, loads
 * or calls it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * Parity status: PINNED.  The reference has no golden vectors of its own, so
 * every function below is checked against the real reference compiled from /root/reference
 * (oracle/_ref/libhmref.so via oracle/ref_harness.cpp) by tests/test_oracle_vs_ref.py, and
 * against the golden vectors generated from that same library (tests/golden/, made by
 * tests/golden/make_golden.py).
 *
 * Written from the behaviour of the reference, not from its text: each function cites the
 * reference file:line whose results it must reproduce bit-for-bit.
 */
#include "hm_oracle.h"
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>

static int iabs(int v) { return v < 0 ? -v : v; }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* =====================================================================================
 * Distortion family
 * ===================================================================================== */

/* SAD of a w x h block.
 * Follows TComRdCost.cpp:465-964 (xGetSAD, xGetSAD4..64, xGetSAD12/24/48):
 *  - the width-specialised functions visit rows 0, 2^s, 2*2^s, ... (s = iSubShift), then
 *    scale the sum by << s, then >> (bitDepth-8)  (DISTORTION_PRECISION_ADJUSTMENT,
 *    TypeDef.h:266-270);
 *  - the generic xGetSAD (:465-491) ignores iSubShift entirely (generic != 0 here).
 * Which one is dispatched is decided by setDistParam (TComRdCost.cpp:294-396). */
uint32_t hmo_sad(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                 int w, int h, int sub_shift, int bit_depth, int generic)
{
  uint32_t sum = 0;
  const int step = generic ? 1 : (1 << sub_shift);
  for (int y = 0; y < h; y += step)
    for (int x = 0; x < w; x++)
      sum += (uint32_t)iabs((int)org[y * org_stride + x] - (int)cur[y * cur_stride + x]);
  if (!generic) sum <<= sub_shift;
  return sum >> (bit_depth - 8);
}

/* In-place 1-D Hadamard of length n (n = 2, 4, 8) with stride `st` over int32 data. */
static void hadamard1d(int32_t* v, int n, int st)
{
  for (int len = 1; len < n; len <<= 1)
    for (int i = 0; i < n; i += 2 * len)
      for (int j = i; j < i + len; j++)
      {
        const int32_t a = v[j * st], b = v[(j + len) * st];
        v[j * st] = a + b;
        v[(j + len) * st] = a - b;
      }
}

/* Sum of |H d H^T| over one n x n tile.  The reference's butterflies (TComRdCost.cpp:
 * 1321-1534) produce the 2-D Hadamard coefficients in a permuted order with some signs
 * flipped; the sum of absolute values is invariant to both, so a textbook Hadamard is
 * an exact restatement.  Rounding per tile: 8x8 (s+2)>>2 (:1531), 4x4 (s+1)>>1 (:1434),
 * 2x2 none (:1341). */
static uint32_t had_tile(const int16_t* org, int os, const int16_t* cur, int cs, int n)
{
  int32_t d[64];
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++)
      d[y * n + x] = (int32_t)org[y * os + x] - (int32_t)cur[y * cs + x];
  for (int y = 0; y < n; y++) hadamard1d(d + y * n, n, 1);
  for (int x = 0; x < n; x++) hadamard1d(d + x, n, n);
  uint32_t s = 0;
  for (int i = 0; i < n * n; i++) s += (uint32_t)iabs(d[i]);
  if (n == 8) return (s + 2) >> 2;
  if (n == 4) return (s + 1) >> 1;
  return s;
}

/* SATD.  Follows xGetHADs (TComRdCost.cpp:1537-1604) / calcHAD (:398-431): 8x8 tiles iff
 * both dimensions are multiples of 8, else 4x4 tiles, else 2x2; per-tile rounding before
 * the sum; total >> (bitDepth-8). */
uint32_t hmo_hads(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                  int w, int h, int bit_depth)
{
  int n;
  if ((w % 8) == 0 && (h % 8) == 0) n = 8;
  else if ((w % 4) == 0 && (h % 4) == 0) n = 4;
  else n = 2;
  uint32_t sum = 0;
  for (int y = 0; y < h; y += n)
    for (int x = 0; x < w; x += n)
      sum += had_tile(org + y * org_stride + x, org_stride, cur + y * cur_stride + x, cur_stride, n);
  return sum >> (bit_depth - 8);
}

/* SSE.  Follows xGetSSE* (TComRdCost.cpp:970-1315): every squared term is shifted by
 * 2*(bitDepth-8) before accumulation into a uint32. */
uint32_t hmo_sse(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                 int w, int h, int bit_depth)
{
  const int sh = (bit_depth - 8) << 1;
  uint32_t sum = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int32_t t = (int32_t)org[y * org_stride + x] - (int32_t)cur[y * cur_stride + x];
      sum += (uint32_t)((t * t) >> sh);
    }
  return sum;
}

/* =====================================================================================
 * MV rate cost
 * ===================================================================================== */

/* Exp-Golomb length of a signed MV difference.  TComRdCost::xGetComponentBits
 * (TComRdCost.cpp:278-292): 2*floor(log2(v<=0 ? -2v+1 : 2v)) + 1. */
uint32_t hmo_component_bits(int v)
{
  uint32_t t = (v <= 0) ? (uint32_t)((-v << 1) + 1) : (uint32_t)(v << 1);
  uint32_t len = 1;
  while (t != 1) { t >>= 1; len += 2; }
  return len;
}

/* TComRdCost::getBits (TComRdCost.h:184-188). */
uint32_t hmo_mv_bits(int pred_x, int pred_y, int scale, int x, int y)
{
  return hmo_component_bits((x << scale) - pred_x) + hmo_component_bits((y << scale) - pred_y);
}

/* TComRdCost::getCost(x,y) (TComRdCost.h:171-178): uint32 wrap-around product >> 16. */
uint32_t hmo_mv_cost(uint32_t ui_cost, int pred_x, int pred_y, int scale, int x, int y)
{
  return (uint32_t)(ui_cost * hmo_mv_bits(pred_x, pred_y, scale, x, y)) >> 16;
}

/* TComRdCost::getCost(bits) (TComRdCost.h:182). */
uint32_t hmo_bits_cost(uint32_t ui_cost, uint32_t bits) { return (uint32_t)(ui_cost * bits) >> 16; }

/* m_uiLambdaMotionSAD[0] = floor(65536*sqrt(lambda)) (TComRdCost.cpp:196-209), selected by
 * getMotionCost(true,0,false) (TComRdCost.h:165). */
uint32_t hmo_lambda_to_cost(double lambda) { return (uint32_t)floor(65536.0 * sqrt(lambda)); }

/* calcRdCost(bits, dist, false, DF_SAD) in COST_STANDARD_LOSSY mode (TComRdCost.cpp:62-122):
 * floor(dist + floor(bits*lambdaSAD + 0.5)/65536) with lambdaSAD the uint32 above. */
double hmo_calc_rd_cost_sad(double lambda, uint32_t bits, uint32_t dist)
{
  const double l = (double)hmo_lambda_to_cost(lambda);
  return floor((double)dist + floor((double)bits * l + 0.5) / 65536.0);
}

/* =====================================================================================
 * Search-window derivation
 * ===================================================================================== */

/* Bounds of TComDataCU::clipMv (TComDataCU.cpp:2917-2929): relative to the CU origin,
 * CTU size 64, offset 8, in quarter-pel. */
void hmo_clip_bounds(int pic_w, int pic_h, int cu_x, int cu_y, int bounds[4])
{
  bounds[0] = (-64 - 8 - cu_x + 1) * 4;
  bounds[1] = (pic_w + 8 - cu_x - 1) * 4;
  bounds[2] = (-64 - 8 - cu_y + 1) * 4;
  bounds[3] = (pic_h + 8 - cu_y - 1) * 4;
}

static void clip_with(const int bd[4], int mv[2])
{
  mv[0] = imin(bd[1], imax(bd[0], mv[0]));
  mv[1] = imin(bd[3], imax(bd[2], mv[1]));
}

void hmo_clip_mv(int pic_w, int pic_h, int cu_x, int cu_y, int mv[2])
{
  int bd[4];
  hmo_clip_bounds(pic_w, pic_h, cu_x, cu_y, bd);
  clip_with(bd, mv);
}

/* int16 wrap + arithmetic shift, as TComMv stores Short (TComMv.h:53-55,112-124). */
static int s16(int v) { return (int)(int16_t)v; }

static void search_range_with(const int bd[4], int pred_x, int pred_y, int srch_rng, int ltrb[4])
{
  int p[2] = { pred_x, pred_y };
  clip_with(bd, p);
  int lt[2] = { s16(p[0] - (srch_rng << 2)), s16(p[1] - (srch_rng << 2)) };
  int rb[2] = { s16(p[0] + (srch_rng << 2)), s16(p[1] + (srch_rng << 2)) };
  clip_with(bd, lt);
  clip_with(bd, rb);
  ltrb[0] = lt[0] >> 2; ltrb[1] = lt[1] >> 2; ltrb[2] = rb[0] >> 2; ltrb[3] = rb[1] >> 2;
}

/* TEncSearch::xSetSearchRange (TEncSearch.cpp:3911-3927). */
void hmo_set_search_range(int pic_w, int pic_h, int cu_x, int cu_y, int pred_x, int pred_y,
                          int srch_rng, int ltrb[4])
{
  int bd[4];
  hmo_clip_bounds(pic_w, pic_h, cu_x, cu_y, bd);
  search_range_with(bd, pred_x, pred_y, srch_rng, ltrb);
}

/* =====================================================================================
 * Interpolation
 * ===================================================================================== */

/* Tap tables: TComInterpolationFilter.cpp:57-75 (these are the H.265 8.5.3.3.3 tables). */
static const int kLuma[4][8] = {
  {  0, 0,   0, 64,  0,   0, 0,  0 },
  { -1, 4, -10, 58, 17,  -5, 1,  0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 },
  {  0, 1,  -5, 17, 58, -10, 4, -1 } };
static const int kChroma[8][4] = {
  {  0, 64,  0,  0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
  { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

#define IF_PREC 14
#define IF_FILT 6
#define IF_OFFS (1 << (IF_PREC - 1))

/* One separable pass.  Reproduces filter<N,isVertical,isFirst,isLast> and filterCopy
 * (TComInterpolationFilter.cpp:94-251) including the int16 store of the un-clipped value. */
static void filter_pass(int ntaps, const int* c, int frac_is_zero, int vertical,
                        const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                        int w, int h, int is_first, int is_last, int bit_depth)
{
  const int head = imax(2, IF_PREC - bit_depth);
  const int max_val = (1 << bit_depth) - 1;
  if (frac_is_zero)
  {
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++)
      {
        const int v = src[y * src_stride + x];
        int o;
        if (is_first == is_last) o = v;                                   /* :98-110 */
        else if (is_first) o = (int16_t)(v << head) - IF_OFFS;            /* :111-126 */
        else                                                              /* :127-147 */
        {
          o = (int16_t)((v + IF_OFFS + (1 << (head - 1))) >> head);
          o = imin(max_val, imax(0, o));
        }
        dst[y * dst_stride + x] = (int16_t)o;
      }
    return;
  }
  const int cs = vertical ? src_stride : 1;
  int shift = IF_FILT, offset;
  if (is_last)
  {
    shift += is_first ? 0 : head;
    offset = 1 << (shift - 1);
    offset += is_first ? 0 : (IF_OFFS << IF_FILT);
  }
  else
  {
    shift -= is_first ? head : 0;
    offset = is_first ? -(IF_OFFS << shift) : 0;
  }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int16_t* p = src + y * src_stride + x - (ntaps / 2 - 1) * cs;
      int sum = 0;
      for (int k = 0; k < ntaps; k++) sum += p[k * cs] * c[k];
      int16_t val = (int16_t)((sum + offset) >> shift);
      if (is_last) { if (val < 0) val = 0; if (val > max_val) val = (int16_t)max_val; }
      dst[y * dst_stride + x] = val;
    }
}

/* Public filterHor (TComInterpolationFilter.cpp:331-350); chroma frac is in 1/8 units (4:2:0). */
void hmo_filter_hor(int chroma, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                    int w, int h, int frac, int is_last, int bit_depth)
{
  filter_pass(chroma ? 4 : 8, chroma ? kChroma[frac] : kLuma[frac], frac == 0, 0,
              src, src_stride, dst, dst_stride, w, h, 1, is_last, bit_depth);
}

/* Public filterVer (TComInterpolationFilter.cpp:364-381). */
void hmo_filter_ver(int chroma, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                    int w, int h, int frac, int is_first, int is_last, int bit_depth)
{
  filter_pass(chroma ? 4 : 8, chroma ? kChroma[frac] : kLuma[frac], frac == 0, 1,
              src, src_stride, dst, dst_stride, w, h, is_first, is_last, bit_depth);
}

/* Replicate padding of TComPicYuv::extendPicBorder (TComPicYuv.cpp:171-215); margin is
 * g_uiMaxCUWidth + 16 = 80 luma samples (TComPicYuv.cpp:87-88).  dst is (w+2m) x (h+2m). */
void hmo_extend_border(const int16_t* src, int w, int h, int margin, int16_t* dst)
{
  const int pw = w + 2 * margin, ph = h + 2 * margin;
  for (int y = 0; y < ph; y++)
  {
    const int sy = imin(h - 1, imax(0, y - margin));
    for (int x = 0; x < pw; x++)
    {
      const int sx = imin(w - 1, imax(0, x - margin));
      dst[y * pw + x] = src[sy * w + sx];
    }
  }
}

/* One interpolated luma block at quarter-pel phase (fx, fy) with the ME two-pass order
 * (horizontal first, not last -> vertical, last), which is what xExtDIFUpSamplingH/Q
 * (TEncSearch.cpp:5565-5766) build, and which equals the uni-prediction MC output
 * (TComPrediction.cpp:680-697). `ref` = top-left integer sample. */
static void interp_block_luma(const int16_t* ref, int ref_stride, int fx, int fy, int w, int h,
                              int bit_depth, int16_t* dst, int dst_stride)
{
  int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)w * (size_t)(h + 8));
  hmo_filter_hor(0, ref - 3 * ref_stride, ref_stride, tmp, w, w, h + 7, fx, 0, bit_depth);
  hmo_filter_ver(0, tmp + 3 * w, w, dst, dst_stride, w, h, fy, 0, 1, bit_depth);
  free(tmp);
}

void hmo_phase_planes(const int16_t* padded, int pw, int ph, int bit_depth, int16_t* planes)
{
  /* clamp-extended copy so that every 8-tap support is readable */
  const int e = 4, ew = pw + 2 * e, eh = ph + 2 * e;
  int16_t* ext = (int16_t*)malloc(sizeof(int16_t) * (size_t)ew * (size_t)eh);
  for (int y = 0; y < eh; y++)
    for (int x = 0; x < ew; x++)
      ext[y * ew + x] = padded[imin(ph - 1, imax(0, y - e)) * pw + imin(pw - 1, imax(0, x - e))];
  for (int fy = 0; fy < 4; fy++)
    for (int fx = 0; fx < 4; fx++)
      interp_block_luma(ext + e * ew + e, ew, fx, fy, pw, ph, bit_depth,
                        planes + (size_t)(fy * 4 + fx) * (size_t)pw * (size_t)ph, pw);
  free(ext);
}

/* =====================================================================================
 * Motion compensation
 * ===================================================================================== */

/* TComPrediction::xPredInterBlk (TComPrediction.cpp:660-698): one component of one block.
 * Luma mv in quarter-pel; chroma (4:2:0) mv in eighth-pel; w,h in component samples.
 * bi != 0 keeps the 14-bit intermediate (isLast = !bi). */
void hmo_pred_inter_blk(int chroma, const int16_t* ref, int ref_stride, int mvx, int mvy,
                        int w, int h, int bi, int bit_depth, int16_t* dst, int dst_stride)
{
  const int sh = chroma ? 3 : 2, ntaps = chroma ? 4 : 8, half = ntaps >> 1;
  const int16_t* r = ref + (mvx >> sh) + (mvy >> sh) * ref_stride;
  const int fx = mvx & ((1 << sh) - 1), fy = mvy & ((1 << sh) - 1);
  if (fy == 0) hmo_filter_hor(chroma, r, ref_stride, dst, dst_stride, w, h, fx, !bi, bit_depth);
  else if (fx == 0) hmo_filter_ver(chroma, r, ref_stride, dst, dst_stride, w, h, fy, 1, !bi, bit_depth);
  else
  {
    int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)w * (size_t)(h + ntaps));
    hmo_filter_hor(chroma, r - (half - 1) * ref_stride, ref_stride, tmp, w, w, h + ntaps - 1, fx, 0, bit_depth);
    hmo_filter_ver(chroma, tmp + (half - 1) * w, w, dst, dst_stride, w, h, fy, 0, !bi, bit_depth);
    free(tmp);
  }
}

/* TComYuv::addAvg (TComYuv.cpp:336-392): bi-prediction average of two 14-bit predictions. */
void hmo_add_avg(const int16_t* s0, int st0, const int16_t* s1, int st1, int w, int h,
                 int bit_depth, int16_t* dst, int dst_stride)
{
  const int shift = imax(2, IF_PREC - bit_depth) + 1;
  const int offset = (1 << (shift - 1)) + 2 * IF_OFFS;
  const int max_val = (1 << bit_depth) - 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int v = (s0[y * st0 + x] + s1[y * st1 + x] + offset) >> shift;
      dst[y * dst_stride + x] = (int16_t)imin(max_val, imax(0, v));
    }
}

/* Bi-pred key pattern 2*org - otherPred, unclipped (TComYuv::removeHighFreq,
 * TComYuv.cpp:393-424 with DISABLING_CLIP_FOR_BIPREDME, TypeDef.h:117). */
void hmo_bipred_key(const int16_t* org, int org_stride, const int16_t* other_pred, int pred_stride,
                    int w, int h, int16_t* dst, int dst_stride)
{
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      dst[y * dst_stride + x] = (int16_t)(2 * org[y * org_stride + x] - other_pred[y * pred_stride + x]);
}

/* =====================================================================================
 * Integer searches
 * ===================================================================================== */

static int sub_shift_for(const hmo_search_t* s)
{
  /* FEN: sub-sampled SAD when rows > 8 (TEncSearch.cpp:347-353, 3950-3956) */
  return (s->fen && s->h > 8) ? 1 : 0;
}

static uint32_t int_cost(hmo_search_t* s, int x, int y)
{
  s->n_cand++;
  /* integer ME always dispatches a width-specialised SAD (AMP override, TComRdCost.cpp:320-333) */
  const int generic = !(s->w == 4 || s->w == 8 || s->w == 16 || s->w == 32 || s->w == 64 ||
                        s->w == 12 || s->w == 24 || s->w == 48);
  return hmo_sad(s->org, s->org_stride, s->ref + y * s->ref_stride + x, s->ref_stride,
                 s->w, s->h, sub_shift_for(s), s->bit_depth, generic)
       + hmo_mv_cost(s->ui_cost, s->pred_x, s->pred_y, 2, x, y);
}

/* Exhaustive raster search, first strict minimum in (y outer, x inner) order.
 * TEncSearch::xPatternSearch (TEncSearch.cpp:3932-3989). */
void hmo_pattern_search(hmo_search_t* s)
{
  uint32_t best = UINT_MAX;
  int bx = 0, by = 0;
  s->n_cand = 0;
  for (int y = s->t; y <= s->b; y++)
    for (int x = s->l; x <= s->r; x++)
    {
      const uint32_t c = int_cost(s, x, y);
      if (c < best) { best = c; bx = x; by = y; }
    }
  s->mv_x = bx; s->mv_y = by;
  s->sad = best - hmo_mv_cost(s->ui_cost, s->pred_x, s->pred_y, 2, bx, by);
}

/* ---- TZ search --------------------------------------------------------------------- */

typedef struct
{
  uint32_t best_cost;
  int best_x, best_y;
  unsigned best_dist, best_round;
  int point_nr;
  int selective;                        /* FastSearch == SELECTIVE: xTZSearchHelp takes its progressive branch */
} tz_state;

/* DistFunc of the selective branch: SAD over rows row0, row0 + 2^sh, ... (H >> sh of them), "<< sh", then the bit-depth
 * shift -- what xGetSAD* return with pOrg/pCur moved down by row0 rows and iSubShift = sh (TComRdCost.cpp:493-964). */
static uint32_t sel_sad(const hmo_search_t* s, int x, int y, int row0, int sh)
{
  const int16_t* cur = s->ref + y * s->ref_stride + x;
  uint32_t sum = 0;
  for (int k = 0; k < (s->h >> sh); k++)
  {
    const int r = row0 + (k << sh);
    for (int i = 0; i < s->w; i++) sum += (uint32_t)iabs(s->org[r * s->org_stride + i] - cur[r * s->ref_stride + i]);
  }
  return (sum << sh) >> (s->bit_depth - 8);
}

/* xTZSearchHelp (TEncSearch.cpp:333-424), selective branch (:360-406): a coarse row-sampled SAD first, refined level by level,
 * abandoned as soon as the running estimate cannot beat the best cost. */
static void tz_try_selective(hmo_search_t* s, tz_state* z, int x, int y, int point_nr, unsigned dist)
{
  s->n_cand++;
  const uint32_t bit_cost = hmo_mv_cost(s->ui_cost, s->pred_x, s->pred_y, 2, x, y);
  int sh = s->h > 32 ? 4 : (s->h > 16 ? 3 : (s->h > 8 ? 2 : 1));
  uint32_t sad = 0;
  uint32_t tmp = sel_sad(s, x, y, 0, sh);
  if (tmp + bit_cost < z->best_cost)
  {
    sad += tmp >> sh;
    while (sh > 0)
    {
      const int is = sh - 1;
      tmp = sel_sad(s, x, y, 1 << is, sh);
      sad += tmp >> sh;
      if (((sad << is) + bit_cost) > z->best_cost) break;
      sh--;
    }
    if (sh == 0)
    {
      sad += bit_cost;
      if (sad < z->best_cost)
      {
        z->best_cost = sad; z->best_x = x; z->best_y = y;
        z->best_dist = dist; z->best_round = 0; z->point_nr = point_nr;
      }
    }
  }
}

/* xTZSearchHelp (TEncSearch.cpp:333-424), non-selective branch: strict '<' update. */
static void tz_try(hmo_search_t* s, tz_state* z, int x, int y, int point_nr, unsigned dist)
{
  if (z->selective) { tz_try_selective(s, z, x, y, point_nr, dist); return; }
  const uint32_t c = int_cost(s, x, y);
  if (c < z->best_cost)
  {
    z->best_cost = c; z->best_x = x; z->best_y = y;
    z->best_dist = dist; z->best_round = 0; z->point_nr = point_nr;
  }
}

/* xTZ8PointDiamondSearch (TEncSearch.cpp:616-791).  The point emission order and the
 * per-point window tests are what matter; the "fully inside" fast path (:668-679, :721-738)
 * emits the same points in the same order as the per-point tests, so one code path serves. */
static void tz_diamond(hmo_search_t* s, tz_state* z, const int win[4], int cx, int cy, int dist)
{
  const int L = win[0], T = win[1], R = win[2], B = win[3];
  const int top = cy - dist, bot = cy + dist, lef = cx - dist, rig = cx + dist;
  z->best_round += 1;
  if (dist == 1)
  {
    if (top >= T) tz_try(s, z, cx, top, 2, dist);
    if (lef >= L) tz_try(s, z, lef, cy, 4, dist);
    if (rig <= R) tz_try(s, z, rig, cy, 5, dist);
    if (bot <= B) tz_try(s, z, cx, bot, 7, dist);
  }
  else if (dist <= 8)
  {
    const int h = dist >> 1;
    const int top2 = cy - h, bot2 = cy + h, lef2 = cx - h, rig2 = cx + h;
    if (top >= T) tz_try(s, z, cx, top, 2, dist);
    if (top2 >= T)
    {
      if (lef2 >= L) tz_try(s, z, lef2, top2, 1, h);
      if (rig2 <= R) tz_try(s, z, rig2, top2, 3, h);
    }
    if (lef >= L) tz_try(s, z, lef, cy, 4, dist);
    if (rig <= R) tz_try(s, z, rig, cy, 5, dist);
    if (bot2 <= B)
    {
      if (lef2 >= L) tz_try(s, z, lef2, bot2, 6, h);
      if (rig2 <= R) tz_try(s, z, rig2, bot2, 8, h);
    }
    if (bot <= B) tz_try(s, z, cx, bot, 7, dist);
  }
  else
  {
    if (top >= T) tz_try(s, z, cx, top, 0, dist);
    if (lef >= L) tz_try(s, z, lef, cy, 0, dist);
    if (rig <= R) tz_try(s, z, rig, cy, 0, dist);
    if (bot <= B) tz_try(s, z, cx, bot, 0, dist);
    for (int i = 1; i < 4; i++)
    {
      const int q = (dist >> 2) * i;
      const int yt = top + q, yb = bot - q, xl = cx - q, xr = cx + q;
      if (yt >= T)
      {
        if (xl >= L) tz_try(s, z, xl, yt, 0, dist);
        if (xr <= R) tz_try(s, z, xr, yt, 0, dist);
      }
      if (yb <= B)
      {
        if (xl >= L) tz_try(s, z, xl, yb, 0, dist);
        if (xr <= R) tz_try(s, z, xr, yb, 0, dist);
      }
    }
  }
}

/* xTZ2PointSearch (TEncSearch.cpp:429-557): the two untested neighbours of the best point,
 * both relative to the best position at entry. */
static void tz_two_point(hmo_search_t* s, tz_state* z, const int win[4])
{
  const int L = win[0], T = win[1], R = win[2], B = win[3];
  const int x = z->best_x, y = z->best_y;
  /* per point-nr (1..8 = position of the best point on the last diamond): the two
   * neighbours not yet tested, in the reference's order.  Each is evaluated iff the
   * coordinate(s) that moved stay inside the window on the side they moved towards. */
  static const int kTwo[9][2][2] = {
    { {0,0}, {0,0} },
    { {-1, 0}, { 0,-1} }, { {-1,-1}, { 1,-1} }, { { 0,-1}, { 1, 0} },
    { {-1, 1}, {-1,-1} }, { { 1,-1}, { 1, 1} },
    { {-1, 0}, { 0, 1} }, { {-1, 1}, { 1, 1} }, { { 1, 0}, { 0, 1} } };
  if (z->point_nr < 1 || z->point_nr > 8) return;      /* the reference asserts here (:551-555) */
  const int nr = z->point_nr;
  for (int i = 0; i < 2; i++)
  {
    const int dx = kTwo[nr][i][0], dy = kTwo[nr][i][1];
    if (dx < 0 && x + dx < L) continue;
    if (dx > 0 && x + dx > R) continue;
    if (dy < 0 && y + dy < T) continue;
    if (dy > 0 && y + dy > B) continue;
    tz_try(s, z, x + dx, y + dy, 0, 2);
  }
}

/* TEncSearch::xTZSearch (TEncSearch.cpp:4027-4228) with TZ_SEARCH_CONFIGURATION (:298-314):
 * diamond first search (stop after 3 rounds without improvement), zero-MV test, optional
 * 2Nx2N integer-MV test + window re-centre (local bounds only),
 * 2-point fill, raster step 5 when best distance > 5, star refinement. */
void hmo_tz_search(hmo_search_t* s)
{
  const int raster = 5;
  int bd[4];
  hmo_clip_bounds(s->pic_w, s->pic_h, s->cu_x, s->cu_y, bd);
  const int win[4] = { s->l, s->t, s->r, s->b };          /* what the pattern helpers see */
  int rwin[4] = { s->l, s->t, s->r, s->b };               /* what the raster scan sees */
  tz_state z; memset(&z, 0, sizeof z); z.best_cost = UINT_MAX;
  s->n_cand = 0;

  int st[2] = { s->start_x, s->start_y };
  clip_with(bd, st);
  tz_try(s, &z, st[0] >> 2, st[1] >> 2, 0, 0);             /* :4054 */
  tz_try(s, &z, 0, 0, 0, 0);                               /* :4071 */
  if (s->has_2nx2n)                                        /* :4074-4093 */
  {
    int m[2] = { s16(s->i2n_x << 2), s16(s->i2n_y << 2) };
    clip_with(bd, m);
    tz_try(s, &z, m[0] >> 2, m[1] >> 2, 0, 0);
    search_range_with(bd, s16(z.best_x << 2), s16(z.best_y << 2), s->search_range, rwin);
  }

  int cx = z.best_x, cy = z.best_y;
  for (int d = 1; d <= s->search_range; d *= 2)            /* :4101-4116 */
  {
    tz_diamond(s, &z, win, cx, cy, d);
    if (z.best_round >= 3) break;
  }
  if (z.best_dist == 1)                                    /* :4137-4141 */
  {
    z.best_dist = 0;
    tz_two_point(s, &z, win);
  }
  if ((int)z.best_dist > raster)                           /* :4144-4154 */
  {
    z.best_dist = raster;
    for (int y = rwin[1]; y <= rwin[3]; y += raster)
      for (int x = rwin[0]; x <= rwin[2]; x += raster)
        tz_try(s, &z, x, y, 0, raster);
  }
  while (z.best_dist > 0)                                  /* :4189-4223 */
  {
    cx = z.best_x; cy = z.best_y;
    z.best_dist = 0; z.point_nr = 0;
    for (int d = 1; d < s->search_range + 1; d *= 2) tz_diamond(s, &z, win, cx, cy, d);
    if (z.best_dist == 1)
    {
      z.best_dist = 0;
      if (z.point_nr != 0) tz_two_point(s, &z, win);
    }
  }
  s->mv_x = z.best_x; s->mv_y = z.best_y;
  s->sad = z.best_cost - hmo_mv_cost(s->ui_cost, s->pred_x, s->pred_y, 2, z.best_x, z.best_y);
}

/* TEncSearch::xTZSearchSelective (TEncSearch.cpp:4231-4383) with SEL_SEARCH_CONFIGURATION (:317-330): start points = the MVP,
 * the left / above / above-right predictors, zero and the optional 2Nx2N integer MV (window re-centred, local bounds only);
 * a step-4 grid of centres within SearchRange/4 of the best start, each followed by the diamonds at distance 1 and 2; then a
 * full raster of the window when the best moved more than 8 samples away, else the star refinement of xTZSearch. */
void hmo_tz_selective(hmo_search_t* s)
{
  int bd[4];
  hmo_clip_bounds(s->pic_w, s->pic_h, s->cu_x, s->cu_y, bd);
  const int win[4] = { s->l, s->t, s->r, s->b };          /* what the pattern helpers see (pcMvSrchRngLT/RB) */
  int rwin[4] = { s->l, s->t, s->r, s->b };               /* iSrchRngHorLeft.. : re-centred when the 2Nx2N MV was tested */
  tz_state z; memset(&z, 0, sizeof z); z.best_cost = UINT_MAX; z.selective = 1;
  s->n_cand = 0;

  int st[2] = { s->start_x, s->start_y };
  clip_with(bd, st);
  tz_try(s, &z, st[0] >> 2, st[1] >> 2, 0, 0);             /* :4267 */
  for (int i = 0; i < 3; i++)                              /* :4270-4279 bTestOtherPredictedMV */
  {
    int m[2] = { s->sel_pred[i][0], s->sel_pred[i][1] };
    clip_with(bd, m);
    tz_try(s, &z, m[0] >> 2, m[1] >> 2, 0, 0);
  }
  tz_try(s, &z, 0, 0, 0, 0);                               /* :4282-4285 */
  if (s->has_2nx2n)                                        /* :4287-4306 */
  {
    int m[2] = { s16(s->i2n_x << 2), s16(s->i2n_y << 2) };
    clip_with(bd, m);
    tz_try(s, &z, m[0] >> 2, m[1] >> 2, 0, 0);
    search_range_with(bd, s16(z.best_x << 2), s16(z.best_y << 2), s->search_range, rwin);
  }

  /* initial search (:4308-4324) */
  const int bx0 = z.best_x, by0 = z.best_y;
  const int sri = s->search_range >> 2, step = 4;
  const int fl = imax(bx0 - sri, rwin[0]), ft = imax(by0 - sri, rwin[1]);
  const int fr = imin(bx0 + sri, rwin[2]), fb = imin(by0 + sri, rwin[3]);
  for (int y = ft; y <= fb; y += step)
    for (int x = fl; x <= fr; x += step)
    {
      tz_try(s, &z, x, y, 0, 0);
      tz_diamond(s, &z, win, x, y, 1);
      tz_diamond(s, &z, win, x, y, 2);
    }

  const int far_from_pred = iabs(z.best_x - bx0) > 8 || iabs(z.best_y - by0) > 8;   /* iMVDistThresh, :4326 */
  if (far_from_pred)                                       /* :4329-4338: every position of the window */
  {
    for (int y = rwin[1]; y <= rwin[3]; y++)
      for (int x = rwin[0]; x <= rwin[2]; x++)
        tz_try(s, &z, x, y, 0, 1);
  }
  else if (z.best_dist > 0)                                /* :4340-4375 star refinement */
  {
    while (z.best_dist > 0)
    {
      const int cx = z.best_x, cy = z.best_y;
      z.best_dist = 0; z.point_nr = 0;
      for (int d = 1; d < s->search_range + 1; d *= 2) tz_diamond(s, &z, win, cx, cy, d);
      if (z.best_dist == 1)
      {
        z.best_dist = 0;
        if (z.point_nr != 0) tz_two_point(s, &z, win);
      }
    }
  }
  s->mv_x = z.best_x; s->mv_y = z.best_y;
  s->sad = z.best_cost - hmo_mv_cost(s->ui_cost, s->pred_x, s->pred_y, 2, z.best_x, z.best_y);
}

/* =====================================================================================
 * Fractional refinement
 * ===================================================================================== */

/* candidate orders of s_acMvRefineH / s_acMvRefineQ (TEncSearch.cpp:51-75) */
static const int kRefineH[9][2] = { {0,0},{0,-1},{0,1},{-1,0},{1,0},{-1,-1},{1,-1},{-1,1},{1,1} };
static const int kRefineQ[9][2] = { {0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1} };

/* distortion of the PU against the reference sampled at quarter-pel MV (qx,qy) */
static uint32_t frac_dist(const hmo_search_t* s, int qx, int qy, int16_t* scratch)
{
  const int16_t* r = s->ref + (qx >> 2) + (qy >> 2) * s->ref_stride;
  interp_block_luma(r, s->ref_stride, qx & 3, qy & 3, s->w, s->h, s->bit_depth, scratch, s->w);
  if (s->hadme && !s->lossless)
    return hmo_hads(s->org, s->org_stride, scratch, s->w, s->w, s->h, s->bit_depth);
  /* SADS by width index; 12/24/48 take the AMP SAD (TComRdCost.cpp:340-381); no sub-sampling */
  const int generic = !(s->w == 4 || s->w == 8 || s->w == 16 || s->w == 32 || s->w == 64 ||
                        s->w == 12 || s->w == 24 || s->w == 48);
  return hmo_sad(s->org, s->org_stride, scratch, s->w, s->w, s->h, 0, s->bit_depth, generic);
}

/* xPatternSearchFracDIF + xPatternRefinement (TEncSearch.cpp:4386-4422, 799-852).
 * Half-pel: 9 candidates (table H) around the integer MV, MV cost at scale 1 on
 * (2*int + h); quarter-pel: 9 candidates (table Q) around the best half, scale 0 on
 * (4*int + 2*half + q).  First strict minimum in table order. */
void hmo_frac_search(hmo_search_t* s)
{
  int16_t* scratch = (int16_t*)malloc(sizeof(int16_t) * (size_t)s->w * (size_t)s->h);
  uint32_t best = UINT_MAX; int bi = 0;
  for (int i = 0; i < 9; i++)
  {
    const int hx = kRefineH[i][0], hy = kRefineH[i][1];
    uint32_t c = frac_dist(s, 4 * s->mv_x + 2 * hx, 4 * s->mv_y + 2 * hy, scratch);
    c += hmo_mv_cost(s->ui_cost, s->pred_x, s->pred_y, 1, 2 * s->mv_x + hx, 2 * s->mv_y + hy);
    s->n_cand++;
    if (c < best) { best = c; bi = i; }
  }
  s->half_x = kRefineH[bi][0]; s->half_y = kRefineH[bi][1];
  best = UINT_MAX; bi = 0;
  for (int i = 0; i < 9; i++)
  {
    const int qx = 4 * s->mv_x + 2 * s->half_x + kRefineQ[i][0];
    const int qy = 4 * s->mv_y + 2 * s->half_y + kRefineQ[i][1];
    uint32_t c = frac_dist(s, qx, qy, scratch);
    c += hmo_mv_cost(s->ui_cost, s->pred_x, s->pred_y, 0, qx, qy);
    s->n_cand++;
    if (c < best) { best = c; bi = i; }
  }
  s->qter_x = kRefineQ[bi][0]; s->qter_y = kRefineQ[bi][1];
  s->frac_cost = best;
  free(scratch);
}

/* Tail of TEncSearch::xMotionEstimation (TEncSearch.cpp:3870-3905): integer search, then
 * fractional search, final MV = (int<<2) + (half<<1) + qter, and
 * cost = floor(w*(cost - getCost(mvBits))) + getCost(bits+mvBits) in double, w = 0.5 for bi. */
void hmo_motion_estimation(hmo_search_t* s, int full_search, int bi, uint32_t* io_bits,
                           int out_mv[2], uint32_t* out_cost)
{
  uint32_t n = 0;
  if (full_search || bi) hmo_pattern_search(s); else hmo_tz_search(s);
  n = s->n_cand;
  s->n_cand = 0;
  hmo_frac_search(s);
  s->n_cand += n;
  out_mv[0] = s16(s16(s->mv_x << 2) + s16(s->half_x << 1) + s->qter_x);
  out_mv[1] = s16(s16(s->mv_y << 2) + s16(s->half_y << 1) + s->qter_y);
  const uint32_t mv_bits = hmo_mv_bits(s->pred_x, s->pred_y, 0, out_mv[0], out_mv[1]);
  *io_bits += mv_bits;
  const double wgt = bi ? 0.5 : 1.0;
  *out_cost = (uint32_t)(floor(wgt * ((double)s->frac_cost - (double)hmo_bits_cost(s->ui_cost, mv_bits)))
                         + (double)hmo_bits_cost(s->ui_cost, *io_bits));
}

/* =====================================================================================
 * Forward transform + scalar quantiser
 * ===================================================================================== */

/* The 31 distinct magnitudes of the HEVC core transform (H.265 8.6.4.2, first column of the
 * 32x32 matrix); the reference carries the same numbers as g_aiT4/8/16/32 (TComRom.cpp:456+,
 * 6-bit set).  C[j] ~ 64*sqrt(2)*cos(j*pi/64). */
static const int kDctC[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                               61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
static const int kDst4[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };

static int dct_coef(int n, int k, int i)
{
  if (k == 0) return 64;
  int m = ((k * (32 / n)) * (2 * i + 1)) % 128;   /* angle in units of pi/64 */
  if (m > 64) m = 128 - m;
  return (m > 32) ? -kDctC[64 - m] : kDctC[m];
}

void hmo_transform_matrix(int n, int32_t* m)
{
  for (int k = 0; k < n; k++)
    for (int i = 0; i < n; i++)
      m[k * n + i] = dct_coef(n, k, i);
}

/* One 1-D stage over `line` rows of length n: dst[k*line + j] = (sum_i M[k][i]*src[j*n+i]
 * + rnd) >> shift.  The reference's partialButterfly4/8/16/32 (TComTrQuant.cpp:387-758)
 * factor exactly this integer sum (no intermediate rounding), fastForwardDst (:413-435)
 * is the same with the DST matrix. */
static void fwd_stage(const int32_t* src, int32_t* dst, int n, int line, int shift, int use_dst)
{
  const int32_t rnd = shift > 0 ? (1 << (shift - 1)) : 0;
  for (int j = 0; j < line; j++)
    for (int k = 0; k < n; k++)
    {
      int32_t acc = 0;
      for (int i = 0; i < n; i++)
        acc += (use_dst ? kDst4[k][i] : dct_coef(n, k, i)) * src[j * n + i];
      dst[k * line + j] = (acc + rnd) >> shift;
    }
}

/* xTrMxN (TComTrQuant.cpp:836-885): rows then columns; shift1 = log2(w) + bitDepth + 6 - 15,
 * shift2 = log2(h) + 6.  DST only for 4x4 with useDST. */
void hmo_fwd_transform(int bit_depth, const int32_t* block, int32_t* coeff, int w, int h, int use_dst)
{
  int lw = 0, lh = 0;
  while ((1 << lw) < w) lw++;
  while ((1 << lh) < h) lh++;
  const int shift1 = lw + bit_depth + 6 - 15, shift2 = lh + 6;
  const int dst = use_dst && w == 4 && h == 4;
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)w * (size_t)h);
  fwd_stage(block, tmp, w, h, shift1, dst);
  fwd_stage(tmp, coeff, h, w, shift2, dst);
  free(tmp);
}

/* one stage of the inverse transform: partialButterflyInverse4/8/16/32 / fastInverseDst (TComTrQuant.cpp:437-810) factor
 * dst[j][k] = clip((sum_i M[i][k] * src[i][j] + rnd) >> shift): the transposed matrix applied to column j, result stored
 * transposed, clipped to [lo, hi] */
static void inv_stage(const int32_t* src, int32_t* dst, int n, int shift, int use_dst, int32_t lo, int32_t hi)
{
  const int32_t rnd = shift > 0 ? (1 << (shift - 1)) : 0;
  for (int j = 0; j < n; j++)
    for (int k = 0; k < n; k++)
    {
      int32_t acc = 0;
      for (int i = 0; i < n; i++)
        acc += (use_dst ? kDst4[i][k] : dct_coef(n, i, k)) * src[i * n + j];
      acc = (acc + rnd) >> shift;
      dst[j * n + k] = acc < lo ? lo : (acc > hi ? hi : acc);
    }
}

/* xITrMxN (TComTrQuant.cpp:894-960): first stage shift 7 with the clip to the 16-bit transform dynamic range, second stage
 * shift 20 - bitDepth with the clip to the Pel range */
void hmo_inv_transform(int bit_depth, const int32_t* coeff, int32_t* block, int n, int use_dst)
{
  const int dst = use_dst && n == 4;
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * (size_t)n);
  inv_stage(coeff, tmp, n, 7, dst, -32768, 32767);
  inv_stage(tmp, block, n, 20 - bit_depth, dst, -32768, 32767);
  free(tmp);
}

/* Scalar (non-RDOQ) quantiser, flat scaling: TComTrQuant::xQuant else-branch
 * (TComTrQuant.cpp:1120-1199) without the sign-hiding post-pass.
 * qbits = 14 + per + transformShift; add = (I ? 171 : 85) << (qbits-9). Returns absSum. */
uint32_t hmo_quant(const int32_t* coef, int n_coef, int qp_per, int qp_rem, int transform_shift,
                   int is_intra_slice, int32_t* level, int32_t* delta_u)
{
  static const int kScale[6] = { 26214, 23302, 20560, 18396, 16384, 14564 }; /* g_quantScales, TComRom.cpp:321 */
  const int qbits = 14 + qp_per + transform_shift;
  const int64_t add = (int64_t)(is_intra_slice ? 171 : 85) << (qbits - 9);
  const int qbits8 = qbits - 8;
  uint32_t abs_sum = 0;
  for (int i = 0; i < n_coef; i++)
  {
    const int32_t c = coef[i];
    const int64_t t = (int64_t)iabs(c) * kScale[qp_rem];
    const int32_t q = (int32_t)((t + add) >> qbits);
    if (delta_u) delta_u[i] = (int32_t)((t - ((int64_t)q << qbits)) >> qbits8);
    abs_sum += (uint32_t)q;
    int32_t v = c < 0 ? -q : q;
    if (v < -32768) v = -32768;
    if (v > 32767) v = 32767;
    level[i] = v;
  }
  return abs_sum;
}

/* =====================================================================================
 * Batch driver over the job / result PODs of include/hmgpu.h (test + cpu_baseline "port")
 * ===================================================================================== */
#include <time.h>
#include "../include/hmgpu.h"

/* refs[slot] -> sample (0,0) of a padded int16 plane (stride ref_stride); org -> (0,0) of the
 * source picture.  Returns CPU seconds (thread time). */
double hmo_me_batch(const hmgpu_me_job* jobs, int n_jobs, const int16_t* const* refs, int ref_stride,
                    const int16_t* org, int org_stride, const int16_t* org_blocks, int bit_depth,
                    hmgpu_me_result* results)
{
  struct timespec t0, t1;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &t0);
  for (int i = 0; i < n_jobs; i++)
  {
    const hmgpu_me_job* j = &jobs[i];
    hmgpu_me_result* r = &results[i];
    hmo_search_t s;
    memset(&s, 0, sizeof s);
    memset(r, 0, sizeof *r);
    s.w = j->pu_w; s.h = j->pu_h;
    if (j->flags & HMGPU_F_ORG_BLOCK) { s.org = org_blocks + j->org_offset; s.org_stride = s.w; }
    else { s.org = org + (size_t)j->pu_y * org_stride + j->pu_x; s.org_stride = org_stride; }
    s.ref = refs[j->ref_slot] + (ptrdiff_t)j->pu_y * ref_stride + j->pu_x; s.ref_stride = ref_stride;
    s.l = j->win_l; s.t = j->win_t; s.r = j->win_r; s.b = j->win_b;
    s.ui_cost = j->ui_cost; s.pred_x = j->pred_x; s.pred_y = j->pred_y;
    s.fen = (j->flags & HMGPU_F_FEN) != 0; s.hadme = (j->flags & HMGPU_F_HADME) != 0;
    s.lossless = (j->flags & HMGPU_F_LOSSLESS) != 0; s.bit_depth = bit_depth;
    s.cu_x = -(j->clip_hmin / 4) - 71; s.cu_y = -(j->clip_vmin / 4) - 71;
    s.pic_w = j->clip_hmax / 4 - 7 + s.cu_x; s.pic_h = j->clip_vmax / 4 - 7 + s.cu_y;
    s.search_range = j->search_range; s.start_x = j->start_x; s.start_y = j->start_y;
    s.has_2nx2n = (j->flags & HMGPU_F_HAS_2NX2N) != 0; s.i2n_x = j->i2n_x; s.i2n_y = j->i2n_y;
    uint32_t n = 0;
    if (j->flags & HMGPU_F_INTEGER)
    {
      if (j->flags & HMGPU_F_FULL) hmo_pattern_search(&s); else hmo_tz_search(&s);
      n = s.n_cand;
      r->int_sad = s.sad;
    }
    else { s.mv_x = j->start_x; s.mv_y = j->start_y; }
    r->int_x = (int16_t)s.mv_x; r->int_y = (int16_t)s.mv_y;
    if (j->flags & HMGPU_F_FRAC)
    {
      s.n_cand = 0;
      hmo_frac_search(&s);
      n += s.n_cand;
      r->half_x = (int16_t)s.half_x; r->half_y = (int16_t)s.half_y;
      r->qter_x = (int16_t)s.qter_x; r->qter_y = (int16_t)s.qter_y;
      r->frac_cost = s.frac_cost;
    }
    r->n_cand = n;
  }
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}


/* ==== intra mode pre-selection ==========================================================
 * Restates TComPrediction::predIntraAng for a luma block without DPCM (TComPrediction.cpp:407-492):
 * xPredIntraPlanar (:746-803), predIntraGetPredValDC (:182-225), xPredIntraAng (:245-405),
 * xDCPredFiltering (:808-835), and the choice of the filtered / unfiltered reference samples,
 * TComPrediction::filteringIntraReferenceSamples (TComPattern.cpp:529-554) with m_aucIntraFilter
 * (TComPrediction.cpp:49-57).  The reference keeps the samples as the first row / first column of a
 * (2n+1) x (2n+1) buffer and copies them into refAbove / refLeft; here one line m[-2n .. 2n] holds both,
 * m[0] = top-left corner, m[k] = above sample k-1, m[-k] = left sample k-1 (k >= 1), so that
 * refAbove[k] = m[k] and refLeft[k] = m[-k]. */
static const int k_intra_ang[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };
static const int k_intra_inv_ang[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };
static const int k_intra_filter_thr[5] = { 10, 7, 1, 0, 10 };   /* 4x4 .. 64x64, luma */

static int hmo_log2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

int hmo_intra_use_filtered(int mode, int n, int no_smooth)
{
  if (no_smooth || mode == 1) return 0;                       /* DC_IDX: never smoothed */
  const int d_hor = abs(mode - 10), d_ver = abs(mode - 26);   /* HOR_IDX 10, VER_IDX 26; planar (0): min(10, 26) = 10 */
  const int diff = d_hor < d_ver ? d_hor : d_ver;
  return diff > k_intra_filter_thr[hmo_log2(n) - 2];
}

void hmo_intra_pred(const int16_t* line, int n, int mode, int bit_depth, int above, int left, int edge_filters, int16_t* dst)
{
  const int16_t* m = line + 2 * n;
  const int lg = hmo_log2(n);
  if (mode == 0)
  {
    /* planar (:746-803): ((n-1-x) L[y] + (x+1) TR + (n-1-y) T[x] + (y+1) BL + n) >> (log2 n + 1), the closed form of the
     * reference's running sums */
    const int tr = m[n + 1], bl = m[-(n + 1)];
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++)
        dst[y * n + x] = (int16_t)(((n - 1 - x) * m[-(y + 1)] + (x + 1) * tr + (n - 1 - y) * m[x + 1] + (y + 1) * bl + n) >> (lg + 1));
    return;
  }
  if (mode == 1)
  {
    int sum = 0, dc;
    if (above) for (int i = 0; i < n; i++) sum += m[i + 1];
    if (left) for (int i = 0; i < n; i++) sum += m[-(i + 1)];
    if (above && left) dc = (sum + n) / (2 * n);
    else if (above) dc = (sum + n / 2) / n;
    else if (left) dc = (sum + n / 2) / n;
    else dc = m[-1];                                          /* pSrc[-1] (:220) */
    for (int i = 0; i < n * n; i++) dst[i] = (int16_t)dc;
    if (above && left && n <= 16)                             /* xDCPredFiltering, luma, MAXIMUM_INTRA_FILTERED_* = 16 */
    {
      dst[0] = (int16_t)((m[1] + m[-1] + 2 * dc + 2) >> 2);
      for (int x = 1; x < n; x++) dst[x] = (int16_t)((m[x + 1] + 3 * dc + 2) >> 2);
      for (int y = 1; y < n; y++) dst[y * n] = (int16_t)((m[-(y + 1)] + 3 * dc + 2) >> 2);
    }
    return;
  }
  /* angular (:273-403).  s = +1: vertical modes (main reference = above), -1: horizontal (main = left; the block is
   * predicted transposed and flipped back) */
  const int ver = mode >= 18;
  const int ang_mode = ver ? mode - 26 : -(mode - 10);
  const int abs_mode = abs(ang_mode), sgn = ang_mode < 0 ? -1 : 1;
  const int angle = sgn * k_intra_ang[abs_mode], inv = k_intra_inv_ang[abs_mode];
  const int s = ver ? 1 : -1;
  const int maxv = (1 << bit_depth) - 1;
  const int edge = edge_filters && n <= 16;
  for (int y = 0; y < n; y++)          /* (x, y) in the orientation of the main reference */
    for (int x = 0; x < n; x++)
    {
      int v;
      if (angle == 0)
      {
        v = m[s * (x + 1)];
        if (edge && x == 0)
        {
          v += (m[-s * (y + 1)] - m[0]) >> 1;                 /* refSide[y+1] - refSide[0] (:355) */
          v = v < 0 ? 0 : (v > maxv ? maxv : v);
        }
      }
      else
      {
        const int pos = (y + 1) * angle, di = pos >> 5, df = pos & 31;
        int k0 = x + di + 1, k1 = k0 + 1, a, b;
        /* main reference at index k: k >= 0 the line itself, k < 0 the side reference projected with the inverse angle
         * (:300-305: refMain[k] = refSide[(128 + |k| * invAngle) >> 8]) */
        a = k0 >= 0 ? m[s * k0] : m[-s * ((128 + (-k0) * inv) >> 8)];
        if (df)
        {
          b = k1 >= 0 ? m[s * k1] : m[-s * ((128 + (-k1) * inv) >> 8)];
          v = ((32 - df) * a + df * b + 16) >> 5;
        }
        else v = a;
      }
      if (ver) dst[y * n + x] = (int16_t)v; else dst[x * n + y] = (int16_t)v;
    }
}

void hmo_intra_costs(const int16_t* line_unfiltered, const int16_t* line_filtered, const int16_t* org, int n, int bit_depth,
                     int flags, uint32_t dist[35])
{
  int16_t pred[64 * 64];
  for (int mode = 0; mode < 35; mode++)
  {
    const int16_t* line = hmo_intra_use_filtered(mode, n, (flags & 16) != 0) ? line_filtered : line_unfiltered;
    hmo_intra_pred(line, n, mode, bit_depth, flags & 1, (flags & 2) != 0, (flags & 4) != 0, pred);
    /* setDistParam(dp, bitDepth, org, stride, pred, stride, w, h, bUseHadamard) (TEncSearch.cpp:2350): DF_HADS or DF_SADS */
    dist[mode] = (flags & 8) ? hmo_hads(org, n, pred, n, n, n, bit_depth) : hmo_sad(org, n, pred, n, n, n, 0, bit_depth, 0);
  }
}


/* ==== SAO statistics ===================================================================
 * Restates TEncSampleAdaptiveOffset::getBlkStats (TEncSampleAdaptiveOffset.cpp:910-1340) with
 * isCalculatePreDeblockSamples = false.  The reference walks each line with running sign buffers; the class of a sample
 * only depends on the sample and its two neighbours along the direction of the type,
 *   edgeType = sgn(c - a) + sgn(c - b),
 * so the statistics are restated per sample, with the region of each type written out: which samples of the block are
 * visited depends on the availability of the neighbouring blocks and on the lines skipped next to the right / bottom
 * boundary (not yet deblocked when the statistics are gathered per CTU). */
static int hmo_sgn(int v) { return (v > 0) - (v < 0); }

void hmo_sao_blk_stats(const int16_t* src, int src_stride, const int16_t* org, int org_stride, int w, int h, int flags,
                       const int32_t skip_r[5], const int32_t skip_b[5], int bit_depth, int64_t diff[5][32], int64_t count[5][32])
{
  const int L = flags & 1, R = (flags >> 1) & 1, A = (flags >> 2) & 1, B = (flags >> 3) & 1, AL = (flags >> 4) & 1, AR = (flags >> 5) & 1;
  memset(diff, 0, sizeof(int64_t) * 5 * 32);
  memset(count, 0, sizeof(int64_t) * 5 * 32);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int16_t* p = src + (ptrdiff_t)y * src_stride + x;
      const int c = p[0], d = org[(ptrdiff_t)y * org_stride + x] - c;
      /* EO_0 (:947-1004): left and right neighbour */
      if (y < (B ? h - skip_b[0] : h) && x >= (L ? 0 : 1) && x < (R ? w - skip_r[0] : w - 1))
      {
        const int e = 2 + hmo_sgn(c - p[-1]) + hmo_sgn(c - p[1]);
        diff[0][e] += d; count[0][e]++;
      }
      /* EO_90 (:1005-1091): above and below */
      if (y >= (A ? 0 : 1) && y < (B ? h - skip_b[1] : h - 1) && x < (R ? w - skip_r[1] : w))
      {
        const int e = 2 + hmo_sgn(c - p[-src_stride]) + hmo_sgn(c - p[src_stride]);
        diff[1][e] += d; count[1][e]++;
      }
      /* EO_135 (:1092-1190): above-left and below-right; the first line has its own range (:1129-1141) */
      {
        const int sx = L ? 0 : 1, ex = R ? w - skip_r[2] : w - 1, ey = B ? h - skip_b[2] : h - 1;
        const int in = y == 0 ? (x >= (AL ? 0 : 1) && x < (A ? ex : 1)) : (y < ey && x >= sx && x < ex);
        if (in)
        {
          const int e = 2 + hmo_sgn(c - p[-src_stride - 1]) + hmo_sgn(c - p[src_stride + 1]);
          diff[2][e] += d; count[2][e]++;
        }
      }
      /* EO_45 (:1191-1292): above-right and below-left; first line :1226-1245 */
      {
        const int sx = L ? 0 : 1, ex = R ? w - skip_r[3] : w - 1, ey = B ? h - skip_b[3] : h - 1;
        const int in = y == 0 ? (x >= (A ? sx : ex) && x < ((!R && AR) ? w : ex)) : (y < ey && x >= sx && x < ex);
        if (in)
        {
          const int e = 2 + hmo_sgn(c - p[-src_stride + 1]) + hmo_sgn(c - p[src_stride - 1]);
          diff[3][e] += d; count[3][e]++;
        }
      }
      /* BO (:1293-1340): band of the sample */
      if (y < (B ? h - skip_b[4] : h) && x < (R ? w - skip_r[4] : w))
      {
        const int b = c >> (bit_depth - 5);
        diff[4][b] += d; count[4][b]++;
      }
    }
}


/* TComSampleAdaptiveOffset::offsetBlock (TComSampleAdaptiveOffset.cpp:309-545), restated per sample like the statistics:
 * the region of each type (first and last line of the diagonal types on their own, :407-415, :441-449, :468-476, :496-504). */
void hmo_sao_offset_block(int type, const int32_t* offset, const int16_t* src, int src_stride, int16_t* res, int res_stride,
                          int w, int h, int flags, int bit_depth)
{
  const int L = flags & 1, R = (flags >> 1) & 1, A = (flags >> 2) & 1, B = (flags >> 3) & 1;
  const int AL = (flags >> 4) & 1, AR = (flags >> 5) & 1, BL = (flags >> 6) & 1, BR = (flags >> 7) & 1;
  const int maxv = (1 << bit_depth) - 1;
  const int sx = L ? 0 : 1, ex = R ? w : w - 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int16_t* p = src + (ptrdiff_t)y * src_stride + x;
      const int c = p[0];
      int in = 0, cls = 0;
      switch (type)
      {
        case 0: in = x >= sx && x < ex; if (in) cls = 2 + hmo_sgn(c - p[-1]) + hmo_sgn(c - p[1]); break;
        case 1: in = y >= (A ? 0 : 1) && y < (B ? h : h - 1); if (in) cls = 2 + hmo_sgn(c - p[-src_stride]) + hmo_sgn(c - p[src_stride]); break;
        case 2:
          if (y == 0) in = x >= (AL ? 0 : 1) && x < (A ? ex : 1);
          else if (y == h - 1) in = x >= (B ? sx : w - 1) && x < (BR ? w : w - 1);
          else in = x >= sx && x < ex;
          if (in) cls = 2 + hmo_sgn(c - p[-src_stride - 1]) + hmo_sgn(c - p[src_stride + 1]);
          break;
        case 3:
          if (y == 0) in = x >= (A ? sx : w - 1) && x < (AR ? w : w - 1);
          else if (y == h - 1) in = x >= (BL ? 0 : 1) && x < (B ? ex : 1);
          else in = x >= sx && x < ex;
          if (in) cls = 2 + hmo_sgn(c - p[-src_stride + 1]) + hmo_sgn(c - p[src_stride - 1]);
          break;
        default: in = 1; cls = c >> (bit_depth - 5); break;
      }
      if (in)
      {
        const int v = c + offset[cls];
        res[(ptrdiff_t)y * res_stride + x] = (int16_t)(v < 0 ? 0 : (v > maxv ? maxv : v));
      }
    }
}


/* ==== deblocking filter ===============================================================
 * Restates the edge filtering of TComLoopFilter: xEdgeFilterLuma (TComLoopFilter.cpp:530-660), xEdgeFilterChroma (:663-790),
 * xPelFilterLuma (:804-860), xPelFilterChroma (:872-890), xUseStrongFiltering (:902-912), xCalcDP / xCalcDQ (:914-922), with the
 * tables sm_tcTable / sm_betaTable (:59-67) and the 4:2:0 chroma QP mapping g_aucChromaScale (TComRom.cpp:499-506).  The
 * reference walks the CU quadtree; the filtering itself only depends on per-unit data (boundary strength, QP, no-filter flag),
 * so it is restated over the picture's grid of 4x4 units.  Pinned to pictures decoded by the instrumented reference decoder
 * (oracle/dbk_dump.inc, tests/golden/make_deblock_golden.py). */
static const uint8_t k_dbk_tc[54] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };
static const uint8_t k_dbk_beta[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };
static const uint8_t k_chroma_scale_420[58] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51 };

static int dbk_clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* one 4-sample segment of a luma edge; p points at sample q0 of line 0, off = step across the edge, step = step along it */
static void dbk_luma_segment(int16_t* p, int off, int step, int bs, int qp_p, int qp_q, int nf_p, int nf_q, int beta_off2, int tc_off2, int bit_depth)
{
  const int qp = (qp_p + qp_q + 1) >> 1, scale = 1 << (bit_depth - 8), maxv = (1 << bit_depth) - 1;
  const int tc = k_dbk_tc[dbk_clip3(0, 53, qp + 2 * (bs - 1) + (tc_off2 << 1))] * scale;     /* DEFAULT_INTRA_TC_OFFSET 2 */
  const int beta = k_dbk_beta[dbk_clip3(0, 51, qp + (beta_off2 << 1))] * scale;
  const int side_thr = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
  int16_t* l0 = p; int16_t* l3 = p + 3 * step;
  const int dp0 = abs(l0[-3 * off] - 2 * l0[-2 * off] + l0[-off]), dq0 = abs(l0[0] - 2 * l0[off] + l0[2 * off]);
  const int dp3 = abs(l3[-3 * off] - 2 * l3[-2 * off] + l3[-off]), dq3 = abs(l3[0] - 2 * l3[off] + l3[2 * off]);
  const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = d0 + d3;
  if (d >= beta) return;
  const int filt_p = dp < side_thr, filt_q = dq < side_thr;
  const int s0 = (abs(l0[-4 * off] - l0[-off]) + abs(l0[3 * off] - l0[0]) < (beta >> 3)) && (2 * d0 < (beta >> 2)) && (abs(l0[-off] - l0[0]) < ((tc * 5 + 1) >> 1));
  const int s3 = (abs(l3[-4 * off] - l3[-off]) + abs(l3[3 * off] - l3[0]) < (beta >> 3)) && (2 * d3 < (beta >> 2)) && (abs(l3[-off] - l3[0]) < ((tc * 5 + 1) >> 1));
  const int sw = s0 && s3;
  for (int i = 0; i < 4; i++)
  {
    int16_t* s = p + i * step;
    const int m0 = s[-4 * off], m1 = s[-3 * off], m2 = s[-2 * off], m3 = s[-off], m4 = s[0], m5 = s[off], m6 = s[2 * off], m7 = s[3 * off];
    int n1 = m1, n2 = m2, n3 = m3, n4 = m4, n5 = m5, n6 = m6;
    if (sw)
    {
      n3 = dbk_clip3(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
      n4 = dbk_clip3(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
      n2 = dbk_clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
      n5 = dbk_clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
      n1 = dbk_clip3(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
      n6 = dbk_clip3(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
    }
    else
    {
      int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
      if (abs(delta) < thr_cut)
      {
        const int tc2 = tc >> 1;
        delta = dbk_clip3(-tc, tc, delta);
        n3 = dbk_clip3(0, maxv, m3 + delta);
        n4 = dbk_clip3(0, maxv, m4 - delta);
        if (filt_p) n2 = dbk_clip3(0, maxv, m2 + dbk_clip3(-tc2, tc2, ((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1)));
        if (filt_q) n5 = dbk_clip3(0, maxv, m5 + dbk_clip3(-tc2, tc2, ((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1)));
      }
    }
    if (!nf_p) { s[-off] = (int16_t)n3; s[-2 * off] = (int16_t)n2; s[-3 * off] = (int16_t)n1; }
    if (!nf_q) { s[0] = (int16_t)n4; s[off] = (int16_t)n5; s[2 * off] = (int16_t)n6; }
  }
}

void hmo_deblock_picture(int16_t* y, int16_t* cb, int16_t* cr, int w, int h, int bit_depth_luma, int bit_depth_chroma,
                         const uint8_t* bs_ver, const uint8_t* bs_hor, const int8_t* qp, const uint8_t* nofilter,
                         int beta_offset_div2, int tc_offset_div2, int cb_qp_offset, int cr_qp_offset)
{
  const int uw = (w + 3) >> 2, uh = (h + 3) >> 2, cw = w >> 1;
  for (int dir = 0; dir < 2; dir++)            /* all vertical edges of the picture, then all horizontal edges (:128-157) */
  {
    const uint8_t* bs = dir ? bs_hor : bs_ver;
    for (int uy = 0; uy < uh; uy++)
      for (int ux = 0; ux < uw; ux++)
      {
        /* edges lie on the 8-sample grid (xDeblockCU steps iEdge by DEBLOCK_SMALLEST_BLOCK / 4 = 2 units, :251-266); the
         * reference's own BS array holds the "internal edge" flag of 4x4 transform units at odd positions, never filtered */
        const int b = ((dir ? uy : ux) & 1) ? 0 : bs[uy * uw + ux];
        if (!b) continue;
        const int q = uy * uw + ux, pn = dir ? q - uw : q - 1;      /* the unit across the edge: left / above */
        /* luma: one 4-sample segment per unit */
        dbk_luma_segment(y + (size_t)(uy * 4) * w + ux * 4, dir ? w : 1, dir ? 1 : w, b, qp[pn], qp[q], nofilter[pn], nofilter[q],
                         beta_offset_div2, tc_offset_div2, bit_depth_luma);
        /* chroma (4:2:0): edges on the 8-sample chroma grid = 16 luma samples, intra boundaries only (bs > 1, :737);
         * a 4x4 luma unit covers 2 chroma samples of the edge */
        if (b > 1 && ((dir ? uy : ux) & 3) == 0)
        {
          for (int c = 0; c < 2; c++)
          {
            int16_t* pl = c ? cr : cb;
            int iqp = ((qp[pn] + qp[q] + 1) >> 1) + (c ? cr_qp_offset : cb_qp_offset);
            if (iqp >= 58) iqp -= 6; else if (iqp >= 0) iqp = k_chroma_scale_420[iqp];
            const int tc = k_dbk_tc[dbk_clip3(0, 53, iqp + 2 * (b - 1) + (tc_offset_div2 << 1))] * (1 << (bit_depth_chroma - 8));
            const int maxv = (1 << bit_depth_chroma) - 1;
            const int off = dir ? cw : 1, step = dir ? 1 : cw;
            int16_t* p0 = pl + (size_t)(uy * 2) * cw + ux * 2;
            for (int i = 0; i < 2; i++)
            {
              int16_t* s = p0 + i * step;
              const int m2 = s[-2 * off], m3 = s[-off], m4 = s[0], m5 = s[off];
              const int delta = dbk_clip3(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
              if (!nofilter[pn]) s[-off] = (int16_t)dbk_clip3(0, maxv, m3 + delta);
              if (!nofilter[q]) s[0] = (int16_t)dbk_clip3(0, maxv, m4 - delta);
            }
          }
        }
      }
  }
}
