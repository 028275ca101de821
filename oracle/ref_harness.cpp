// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" shim (ours) linked against the UNMODIFIED HM-16.2 reference objects
// (compiled from /root/reference by oracle/Makefile into oracle/_ref/libhmref.so) so
// that python/ctypes can call the real reference functions of the inter-search hot
// path.  It is used for two things only:
//   * pinning oracle/hm_oracle.c (our restatement) against the real reference, and
//   * generating the golden vectors committed under tests/golden/.
// Nothing in the product path (libhmgpu) links, loads or calls this file.
//
// Access to protected/private members of TEncSearch / TComRdCost / TComDataCU is
// obtained with the classic "#define private public" trick: it changes no layout
// under the Itanium ABI / GCC, only access checks at compile time.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cassert>
#include <cmath>
#include <limits>
#include <vector>
#include <list>
#include <map>
#include <string>
#include <sstream>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <iomanip>
#include <stdint.h>

#define private public
#define protected public
#include "TLibCommon/TypeDef.h"
#include "TLibCommon/CommonDef.h"
#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComMv.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibCommon/TComInterpolationFilter.h"
#include "TLibCommon/TComTrQuant.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComYuv.h"
#include "TLibCommon/TComSlice.h"
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComPrediction.h"
#include "TLibEncoder/TEncCfg.h"
#include "TLibEncoder/TEncSearch.h"
#include "TLibEncoder/TEncSampleAdaptiveOffset.h"
#undef private
#undef protected

// free functions with external linkage in TComTrQuant.cpp:387-885 (not in any header)
extern Void partialButterfly4 (TCoeff *src, TCoeff *dst, Int shift, Int line);
extern Void partialButterfly8 (TCoeff *src, TCoeff *dst, Int shift, Int line);
extern Void partialButterfly16(TCoeff *src, TCoeff *dst, Int shift, Int line);
extern Void partialButterfly32(TCoeff *src, TCoeff *dst, Int shift, Int line);
extern Void fastForwardDst(TCoeff *block, TCoeff *coeff, Int shift);
extern Void xTrMxN(Int bitDepth, TCoeff *block, TCoeff *coeff, Int iWidth, Int iHeight, Bool useDST, const Int maxTrDynamicRange);
extern Void xITrMxN(Int bitDepth, TCoeff *coeff, TCoeff *block, Int iWidth, Int iHeight, Bool useDST, const Int maxTrDynamicRange);

namespace {

struct RefState
{
  bool         inited;
  TComRdCost   rd;
  TEncCfg      cfg;
  TEncSearch*  search;
  TComSPS      sps;
  TComSlice*   slice;
  TComDataCU*  cu;
  RefState() : inited(false), search(NULL), slice(NULL), cu(NULL) {}
};
RefState g;

void ensure_init()
{
  if (g.inited) return;
  initROM();
  g.rd.init();
  g.cfg.setChromaFormatIdc(CHROMA_420);
  g.cfg.setQuadtreeTULog2MaxSize(5);
  g.cfg.setQuadtreeTULog2MinSize(2);
  g.cfg.setFastSearch(1);
  g.cfg.setUseFastEnc(true);
  g.cfg.setUseHADME(true);
  g_uiMaxCUDepth = 4;
  g.search = new TEncSearch;
  g.search->init(&g.cfg, NULL, 64, 4, 1, 0, NULL, &g.rd, NULL, NULL);
  g.slice = new TComSlice;
  g.slice->setSPS(&g.sps);
  g.cu = new TComDataCU;
  g.cu->m_pcSlice = g.slice;
  g.inited = true;
}

void set_bitdepth(int bd)
{
  g_bitDepth[CHANNEL_TYPE_LUMA] = bd;
  g_bitDepth[CHANNEL_TYPE_CHROMA] = bd;
  g_maxTrDynamicRange[CHANNEL_TYPE_LUMA] = 15;
  g_maxTrDynamicRange[CHANNEL_TYPE_CHROMA] = 15;
}

void set_cu(int picW, int picH, int cuX, int cuY)
{
  g.sps.setPicWidthInLumaSamples(picW);
  g.sps.setPicHeightInLumaSamples(picH);
  g.cu->m_uiCUPelX = cuX;
  g.cu->m_uiCUPelY = cuY;
}

void set_cost(unsigned uiCost, int predX, int predY, int scale)
{
  TComMv p(predX, predY);
  g.rd.m_uiCost = uiCost;
  g.rd.setPredictor(p);
  g.rd.setCostScale(scale);
}

} // namespace

extern "C" {

int ref_version() { return 1602; }

// ---- a1/a2/a3: integer-ME SAD through setDistParam(TComPattern*,...) (TComRdCost.cpp:305-337),
//      iSubShift applied by the caller exactly like TEncSearch.cpp:347-353 / 3950-3956.
unsigned ref_sad_me(const int16_t* org, int orgStride, const int16_t* cur, int curStride,
                    int w, int h, int subShift, int bitDepth)
{
  ensure_init(); set_bitdepth(bitDepth);
  TComPattern pat; pat.initPattern((Pel*)org, w, h, orgStride);
  DistParam dp;
  g.rd.setDistParam(&pat, (Pel*)cur, curStride, dp);
  dp.iSubShift = subShift;
  dp.bitDepth = bitDepth;
  dp.bApplyWeight = false;
  return dp.DistFunc(&dp);
}

// ---- sub-pel ME distortion through setDistParam(pattern, ref, stride, iStep=1, dp, bHADME)
//      (TComRdCost.cpp:340-381): SADS (hadamard=0) or HADS (hadamard=1).
unsigned ref_dist_subpel(const int16_t* org, int orgStride, const int16_t* cur, int curStride,
                         int w, int h, int hadamard, int bitDepth)
{
  ensure_init(); set_bitdepth(bitDepth);
  TComPattern pat; pat.initPattern((Pel*)org, w, h, orgStride);
  DistParam dp;
  g.rd.setDistParam(&pat, (Pel*)cur, curStride, 1, dp, hadamard != 0);
  dp.bitDepth = bitDepth;
  dp.bApplyWeight = false;
  return dp.DistFunc(&dp);
}

// ---- generic two-pointer setDistParam (TComRdCost.cpp:384-396): DF_SADS/DF_HADS by width index
unsigned ref_dist_generic(const int16_t* p1, int s1, const int16_t* p2, int s2,
                          int w, int h, int hadamard, int bitDepth, int subShift)
{
  ensure_init(); set_bitdepth(bitDepth);
  DistParam dp;
  g.rd.setDistParam(dp, bitDepth, (Pel*)p1, s1, (Pel*)p2, s2, w, h, hadamard != 0);
  dp.iSubShift = subShift;
  dp.bApplyWeight = false;
  return dp.DistFunc(&dp);
}

unsigned ref_calc_had(const int16_t* p0, int s0, const int16_t* p1, int s1, int w, int h, int bitDepth)
{
  ensure_init(); set_bitdepth(bitDepth);
  return g.rd.calcHAD(bitDepth, (Pel*)p0, s0, (Pel*)p1, s1, w, h);
}

// ---- a5: SSE through getDistPart (TComRdCost.cpp:433-456), luma
unsigned ref_sse(const int16_t* cur, int curStride, const int16_t* org, int orgStride, int w, int h, int bitDepth)
{
  ensure_init(); set_bitdepth(bitDepth);
  return g.rd.getDistPart(bitDepth, (Pel*)cur, curStride, (Pel*)org, orgStride, w, h, COMPONENT_Y, DF_SSE);
}

// ---- a6: MV rate cost (TComRdCost.h:163-189, .cpp:278-292)
unsigned ref_lambda_to_cost(double lambda)
{
  ensure_init();
  g.rd.setLambda(lambda);
  g.rd.getMotionCost(true, 0, false);
  return g.rd.m_uiCost;
}
unsigned ref_mv_cost(unsigned uiCost, int predX, int predY, int scale, int x, int y)
{
  ensure_init(); set_cost(uiCost, predX, predY, scale);
  return g.rd.getCost(x, y);
}
unsigned ref_mv_bits(int predX, int predY, int scale, int x, int y)
{
  ensure_init(); set_cost(0, predX, predY, scale);
  return g.rd.getBits(x, y);
}
unsigned ref_bits_cost(unsigned uiCost, unsigned bits)
{
  ensure_init(); g.rd.m_uiCost = uiCost;
  return g.rd.getCost(bits);
}
double ref_calc_rd_cost_sad(double lambda, unsigned bits, unsigned dist)
{
  ensure_init(); g.rd.setLambda(lambda);
  return g.rd.calcRdCost(bits, dist, false, DF_SAD);
}

// ---- a8: clipMv (TComDataCU.cpp:2917-2929) and xSetSearchRange (TEncSearch.cpp:3911-3927)
void ref_clip_mv(int picW, int picH, int cuX, int cuY, int* mv)
{
  ensure_init(); set_cu(picW, picH, cuX, cuY);
  TComMv m(mv[0], mv[1]); g.cu->clipMv(m); mv[0] = m.getHor(); mv[1] = m.getVer();
}
void ref_set_search_range(int picW, int picH, int cuX, int cuY, int predX, int predY, int srchRng, int* ltrb)
{
  ensure_init(); set_cu(picW, picH, cuX, cuY);
  TComMv p(predX, predY), lt, rb;
  g.search->xSetSearchRange(g.cu, p, srchRng, lt, rb);
  ltrb[0] = lt.getHor(); ltrb[1] = lt.getVer(); ltrb[2] = rb.getHor(); ltrb[3] = rb.getVer();
}

// ---- a14: interpolation primitives (TComInterpolationFilter.cpp:331-381)
void ref_filter_hor(int chroma, const int16_t* src, int srcStride, int16_t* dst, int dstStride,
                    int w, int h, int frac, int isLast, int bitDepth)
{
  ensure_init(); set_bitdepth(bitDepth);
  TComInterpolationFilter f;
  f.filterHor(chroma ? COMPONENT_Cb : COMPONENT_Y, (Pel*)src, srcStride, (Pel*)dst, dstStride, w, h, frac, isLast != 0, CHROMA_420);
}
void ref_filter_ver(int chroma, const int16_t* src, int srcStride, int16_t* dst, int dstStride,
                    int w, int h, int frac, int isFirst, int isLast, int bitDepth)
{
  ensure_init(); set_bitdepth(bitDepth);
  TComInterpolationFilter f;
  f.filterVer(chroma ? COMPONENT_Cb : COMPONENT_Y, (Pel*)src, srcStride, (Pel*)dst, dstStride, w, h, frac, isFirst != 0, isLast != 0, CHROMA_420);
}

// ---- a19: picture allocation + border extension (TComPicYuv.cpp:81-134,171-215)
// in: unpadded w x h luma; out: (w+2*margin) x (h+2*margin) padded plane, returns margin
int ref_extend_border(const int16_t* src, int w, int h, int16_t* dst, int dstCapacityElems)
{
  ensure_init();
  TComPicYuv pic;
  pic.create(w, h, CHROMA_420, g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth);
  Pel* y = pic.getAddr(COMPONENT_Y);
  const int stride = pic.getStride(COMPONENT_Y);
  for (int r = 0; r < h; r++) memcpy(y + r * stride, src + r * w, w * sizeof(Pel));
  pic.extendPicBorder();
  const int mx = pic.getMarginX(COMPONENT_Y), my = pic.getMarginY(COMPONENT_Y);
  const int W = w + 2 * mx, H = h + 2 * my;
  if (W * H <= dstCapacityElems)
  {
    for (int r = 0; r < H; r++) memcpy(dst + r * W, y + (r - my) * stride - mx, W * sizeof(Pel));
  }
  pic.destroy();
  return (mx == my) ? mx : -1;
}

// ---- a9: xPatternSearch (TEncSearch.cpp:3932-3989).  `ref` points at the PU origin inside a
// padded reference plane (so negative offsets are readable).
void ref_pattern_search(const int16_t* org, int orgStride, int w, int h,
                        const int16_t* ref, int refStride,
                        int l, int t, int r, int b,
                        unsigned uiCost, int predX, int predY, int fen, int bitDepth,
                        int* outMv, unsigned* outSad)
{
  ensure_init(); set_bitdepth(bitDepth);
  g.cfg.setUseFastEnc(fen != 0);
  set_cost(uiCost, predX, predY, 2);
  TComPattern pat; pat.initPattern((Pel*)org, w, h, orgStride);
  TComMv lt(l, t), rb(r, b), mv;
  Distortion sad = 0;
  g.search->m_cDistParam.bApplyWeight = false;
  g.search->xPatternSearch(&pat, (Pel*)ref, refStride, &lt, &rb, mv, sad);
  outMv[0] = mv.getHor(); outMv[1] = mv.getVer(); *outSad = sad;
}

// ---- a10: xTZSearch (TEncSearch.cpp:4027-4228); mvInOut = start MV in quarter-pel (the MVP)
void ref_tz_search(const int16_t* org, int orgStride, int w, int h,
                   const int16_t* ref, int refStride,
                   int l, int t, int r, int b,
                   unsigned uiCost, int predX, int predY, int fen, int bitDepth,
                   int picW, int picH, int cuX, int cuY, int searchRange,
                   int has2Nx2N, int i2NX, int i2NY,
                   int* mvInOut, unsigned* outSad)
{
  ensure_init(); set_bitdepth(bitDepth);
  g.cfg.setUseFastEnc(fen != 0);
  g.cfg.setFastSearch(1);
  set_cu(picW, picH, cuX, cuY);
  set_cost(uiCost, predX, predY, 2);
  g.search->m_iSearchRange = searchRange;
  TComPattern pat; pat.initPattern((Pel*)org, w, h, orgStride);
  TComMv lt(l, t), rb(r, b), mv(mvInOut[0], mvInOut[1]);
  TComMv i2n(i2NX, i2NY);
  Distortion sad = 0;
  g.search->m_cDistParam.bApplyWeight = false;
  g.search->xTZSearch(g.cu, &pat, (Pel*)ref, refStride, &lt, &rb, mv, sad, has2Nx2N ? &i2n : NULL);
  mvInOut[0] = mv.getHor(); mvInOut[1] = mv.getVer(); *outSad = sad;
}

// ---- a11: xTZSearchSelective (TEncSearch.cpp:4231-4383), FastSearch = 2; selPred = m_acMvPredictors (left, above,
// above-right; quarter-pel hor/ver pairs)
void ref_tz_selective(const int16_t* org, int orgStride, int w, int h,
                      const int16_t* ref, int refStride,
                      int l, int t, int r, int b,
                      unsigned uiCost, int predX, int predY, int bitDepth,
                      int picW, int picH, int cuX, int cuY, int searchRange,
                      int has2Nx2N, int i2NX, int i2NY,
                      const int* selPred, int* mvInOut, unsigned* outSad)
{
  ensure_init(); set_bitdepth(bitDepth);
  g.cfg.setFastSearch(2);
  set_cu(picW, picH, cuX, cuY);
  set_cost(uiCost, predX, predY, 2);
  g.search->m_iSearchRange = searchRange;
  for (int i = 0; i < 3; i++) g.search->m_acMvPredictors[i].set(selPred[2 * i], selPred[2 * i + 1]);
  TComPattern pat; pat.initPattern((Pel*)org, w, h, orgStride);
  TComMv lt(l, t), rb(r, b), mv(mvInOut[0], mvInOut[1]);
  TComMv i2n(i2NX, i2NY);
  Distortion sad = 0;
  g.search->m_cDistParam.bApplyWeight = false;
  g.search->xTZSearchSelective(g.cu, &pat, (Pel*)ref, refStride, &lt, &rb, mv, sad, has2Nx2N ? &i2n : NULL);
  g.cfg.setFastSearch(1);
  mvInOut[0] = mv.getHor(); mvInOut[1] = mv.getVer(); *outSad = sad;
}

// ---- a12/a13: xPatternSearchFracDIF (TEncSearch.cpp:4386-4422); cost scale follows
// xMotionEstimation (:3892): scale 1 on entry, the function itself switches to 0.
void ref_frac_search(const int16_t* org, int orgStride, int w, int h,
                     const int16_t* ref, int refStride, int mvIntX, int mvIntY,
                     unsigned uiCost, int predX, int predY, int hadme, int lossless, int bitDepth,
                     int* outHalf, int* outQter, unsigned* outCost)
{
  ensure_init(); set_bitdepth(bitDepth);
  g.cfg.setUseHADME(hadme != 0);
  set_cost(uiCost, predX, predY, 1);
  TComPattern pat; pat.initPattern((Pel*)org, w, h, orgStride);
  TComMv mvInt(mvIntX, mvIntY), half, qter;
  Distortion cost = 0;
  g.search->m_cDistParam.bApplyWeight = false;
  g.search->xPatternSearchFracDIF(lossless != 0, &pat, (Pel*)ref, refStride, &mvInt, half, qter, cost, false);
  outHalf[0] = half.getHor(); outHalf[1] = half.getVer();
  outQter[0] = qter.getHor(); outQter[1] = qter.getVer();
  *outCost = cost;
}

// ---- CPU baseline: a batch of xMotionEstimation bodies run by the real reference -------------
// jobs/results use the POD layouts of include/hmgpu.h (hmgpu_me_job / hmgpu_me_result) so the
// very same work-list can be given to the GPU library and to the reference.  refs[slot] points
// at sample (0,0) of a padded int16 plane with stride ref_stride; org at (0,0) of the source
// picture.  Returns the CPU seconds spent inside the reference calls (CLOCK_THREAD_CPUTIME_ID).
#include <time.h>
#include "../include/hmgpu.h"

double ref_me_batch(const hmgpu_me_job* jobs, int n_jobs, const int16_t* const* refs, int ref_stride,
                    const int16_t* org, int org_stride, const int16_t* org_blocks, int bitDepth,
                    hmgpu_me_result* results)
{
  ensure_init(); set_bitdepth(bitDepth);
  g.search->m_cDistParam.bApplyWeight = false;
  struct timespec t0, t1;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &t0);
  for (int i = 0; i < n_jobs; i++)
  {
    const hmgpu_me_job& j = jobs[i];
    hmgpu_me_result& r = results[i];
    memset(&r, 0, sizeof r);
    const int w = j.pu_w, h = j.pu_h;
    TComPattern pat;
    if (j.flags & HMGPU_F_ORG_BLOCK) pat.initPattern((Pel*)org_blocks + j.org_offset, w, h, w);
    else pat.initPattern((Pel*)org + (size_t)j.pu_y * org_stride + j.pu_x, w, h, org_stride);
    Pel* ref = (Pel*)refs[j.ref_slot] + (ptrdiff_t)j.pu_y * ref_stride + j.pu_x;
    g.cfg.setUseFastEnc((j.flags & HMGPU_F_FEN) != 0);
    g.cfg.setUseHADME((j.flags & HMGPU_F_HADME) != 0);
    g.cfg.setFastSearch(1);
    // clipMv reads (pic size, CU origin): recover them from the bounds the job carries
    const int cuX = -(j.clip_hmin / 4) - 71, cuY = -(j.clip_vmin / 4) - 71;
    set_cu(j.clip_hmax / 4 - 7 + cuX, j.clip_vmax / 4 - 7 + cuY, cuX, cuY);
    set_cost(j.ui_cost, j.pred_x, j.pred_y, 2);
    g.search->m_iSearchRange = j.search_range;
    TComMv lt(j.win_l, j.win_t), rb(j.win_r, j.win_b), mv(j.start_x, j.start_y);
    Distortion sad = 0;
    if (j.flags & HMGPU_F_INTEGER)
    {
      if (j.flags & HMGPU_F_FULL)
      {
        g.search->xPatternSearch(&pat, ref, ref_stride, &lt, &rb, mv, sad);
      }
      else
      {
        TComMv i2n(j.i2n_x, j.i2n_y);
        g.search->xTZSearch(g.cu, &pat, ref, ref_stride, &lt, &rb, mv, sad, (j.flags & HMGPU_F_HAS_2NX2N) ? &i2n : NULL);
      }
      r.int_sad = sad;
    }
    r.int_x = mv.getHor(); r.int_y = mv.getVer();
    if (j.flags & HMGPU_F_FRAC)
    {
      g.rd.setCostScale(1);
      TComMv half, qter;
      Distortion cost = 0;
      g.search->xPatternSearchFracDIF((j.flags & HMGPU_F_LOSSLESS) != 0, &pat, ref, ref_stride, &mv, half, qter, cost, false);
      r.half_x = half.getHor(); r.half_y = half.getVer();
      r.qter_x = qter.getHor(); r.qter_y = qter.getVer();
      r.frac_cost = cost;
    }
  }
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// ---- a17: forward transform (TComTrQuant.cpp:836-885); block is w*h TCoeff (int32) row-major
void ref_fwd_transform(int bitDepth, const int32_t* block, int32_t* coeff, int w, int h, int useDST)
{
  ensure_init();
  std::vector<TCoeff> tmp(block, block + w * h);
  xTrMxN(bitDepth, &tmp[0], (TCoeff*)coeff, w, h, useDST != 0, 15);
}
void ref_inv_transform(int bitDepth, const int32_t* coeff, int32_t* block, int n, int useDST)
{
  ensure_init();
  std::vector<TCoeff> tmp(coeff, coeff + n * n);
  xITrMxN(bitDepth, &tmp[0], (TCoeff*)block, n, n, useDST != 0, 15);
}
void ref_partial_butterfly(int n, const int32_t* src, int32_t* dst, int shift, int line)
{
  std::vector<TCoeff> tmp(src, src + n * line);
  switch (n)
  {
    case 4:  partialButterfly4 (&tmp[0], (TCoeff*)dst, shift, line); break;
    case 8:  partialButterfly8 (&tmp[0], (TCoeff*)dst, shift, line); break;
    case 16: partialButterfly16(&tmp[0], (TCoeff*)dst, shift, line); break;
    case 32: partialButterfly32(&tmp[0], (TCoeff*)dst, shift, line); break;
    default: break;
  }
}
int ref_quant_scale(int rem) { return g_quantScales[rem]; }

// ---- a15: luma/chroma uni-directional MC block (TComPrediction.cpp:660-698, xPredInterBlk).
// `ref` = pointer to the block origin inside a padded plane; mv in quarter-pel (luma) or
// eighth-pel units after the 4:2:0 scaling (chroma), exactly what xPredInterBlk derives.
void ref_pred_inter_blk(int chroma, const int16_t* ref, int refStride, int mvx, int mvy,
                        int w, int h, int bi, int bitDepth, int16_t* dst, int dstStride)
{
  ensure_init(); set_bitdepth(bitDepth);
  // Restates the pointer arithmetic of xPredInterBlk with the reference's own filters:
  TComInterpolationFilter f;
  const ComponentID comp = chroma ? COMPONENT_Cb : COMPONENT_Y;
  const int shiftHor = 2 + (chroma ? 1 : 0), shiftVer = 2 + (chroma ? 1 : 0);
  const int16_t* r = ref + (mvx >> shiftHor) + (mvy >> shiftVer) * refStride;
  const int xFrac = mvx & ((1 << shiftHor) - 1), yFrac = mvy & ((1 << shiftVer) - 1);
  const int csxy = chroma ? 1 : 0; // getComponentScaleX for 4:2:0 chroma
  const int fx = chroma ? xFrac : xFrac, fy = chroma ? yFrac : yFrac;
  (void)csxy;
  // the public filterHor/filterVer of the reference take frac in component units
  // (TComInterpolationFilter.cpp:346,379 shift chroma frac by (1-csx)); for 4:2:0, csx=1.
  if (yFrac == 0)
  {
    f.filterHor(comp, (Pel*)r, refStride, (Pel*)dst, dstStride, w, h, fx, !bi, CHROMA_420);
  }
  else if (xFrac == 0)
  {
    f.filterVer(comp, (Pel*)r, refStride, (Pel*)dst, dstStride, w, h, fy, true, !bi, CHROMA_420);
  }
  else
  {
    const int ntaps = chroma ? NTAPS_CHROMA : NTAPS_LUMA;
    const int half = ntaps >> 1;
    const int tmpStride = w;
    std::vector<Pel> tmp((h + ntaps) * tmpStride);
    f.filterHor(comp, (Pel*)r - (half - 1) * refStride, refStride, &tmp[0], tmpStride, w, h + ntaps - 1, fx, false, CHROMA_420);
    f.filterVer(comp, &tmp[0] + (half - 1) * tmpStride, tmpStride, (Pel*)dst, dstStride, w, h, fy, false, !bi, CHROMA_420);
  }
}

// ---- f4: one luma intra prediction exactly as TComPrediction::predIntraAng (TComPrediction.cpp:407-492) performs it for a
//      block without DPCM.  line = the 4n+1 reference samples from the bottom-left neighbour up to the top-left corner and on
//      to the above-right neighbour; they are laid out as the first column / first row of the (2n+1) x (2n+1) buffer the
//      reference predicts from (m_piYuvExt).
void ref_intra_pred(int bitDepth, const int16_t* line, int n, int mode, int above, int left, int edgeFilters, int16_t* dst)
{
  ensure_init(); set_bitdepth(bitDepth);
  const int sw = 2 * n + 1;
  std::vector<Pel> roi(sw * sw, 0);
  for (int y = 0; y < 2 * n; y++) roi[(y + 1) * sw] = line[2 * n - 1 - y];     // left column, top to bottom
  for (int x = 0; x <= 2 * n; x++) roi[x] = line[2 * n + x];                  // top-left corner + top row
  const Pel* src = &roi[0] + sw + 1;
  if (mode == PLANAR_IDX) g.search->xPredIntraPlanar(src, sw, (Pel*)dst, n, n, n, CHANNEL_TYPE_LUMA, CHROMA_420);
  else
  {
    g.search->xPredIntraAng(bitDepth, src, sw, (Pel*)dst, n, n, n, CHANNEL_TYPE_LUMA, CHROMA_420, mode, above != 0, left != 0, edgeFilters != 0);
    if (mode == DC_IDX && above && left) g.search->xDCPredFiltering(src, sw, (Pel*)dst, n, n, n, CHANNEL_TYPE_LUMA);
  }
}
int ref_intra_use_filtered(int mode, int n, int disableSmoothing)
{
  ensure_init();
  return TComPrediction::filteringIntraReferenceSamples(COMPONENT_Y, mode, n, n, CHROMA_420, disableSmoothing != 0) ? 1 : 0;
}

// ---- f3 (first half): the real TEncSampleAdaptiveOffset::getBlkStats on one block of one component, without the pre-deblock
//      sample mode.  flags: 1 left, 2 right, 4 above, 8 below, 16 above-left, 32 above-right available.
void ref_sao_blk_stats(const int16_t* src, int srcStride, const int16_t* org, int orgStride, int w, int h, int flags,
                       const int32_t* skipR, const int32_t* skipB, int bitDepth, int64_t* diff /* [5][32] */, int64_t* count)
{
  ensure_init(); set_bitdepth(bitDepth);
  static TEncSampleAdaptiveOffset* sao = NULL;
  if (!sao) { sao = new TEncSampleAdaptiveOffset; sao->m_maxCUWidth = 64; sao->m_maxCUHeight = 64; }
  for (int t = 0; t < NUM_SAO_NEW_TYPES; t++)
  {
    sao->m_skipLinesR[COMPONENT_Y][t] = skipR[t];
    sao->m_skipLinesB[COMPONENT_Y][t] = skipB[t];
  }
  SAOStatData st[NUM_SAO_NEW_TYPES];
  sao->getBlkStats(COMPONENT_Y, st, (Pel*)src, (Pel*)org, srcStride, orgStride, w, h,
                   (flags & 1) != 0, (flags & 2) != 0, (flags & 4) != 0, (flags & 8) != 0, (flags & 16) != 0, (flags & 32) != 0,
                   false, false
#if SAO_ENCODE_ALLOW_USE_PREDEBLOCK
                   , false
#endif
                   );
  for (int t = 0; t < NUM_SAO_NEW_TYPES; t++)
    for (int c = 0; c < 32; c++) { diff[t * 32 + c] = st[t].diff[c]; count[t * 32 + c] = st[t].count[c]; }
}

// f3: the real TComSampleAdaptiveOffset::offsetBlock.  flags as ref_sao_blk_stats plus 64 below-left, 128 below-right.
void ref_sao_offset_block(int type, const int32_t* offset, const int16_t* src, int srcStride, int16_t* res, int resStride,
                          int w, int h, int flags, int bitDepth)
{
  ensure_init(); set_bitdepth(bitDepth);
  static TEncSampleAdaptiveOffset* sao = NULL;
  if (!sao) { sao = new TEncSampleAdaptiveOffset; sao->m_maxCUWidth = 64; sao->m_maxCUHeight = 64; }
  Int off[MAX_NUM_SAO_CLASSES];
  for (int i = 0; i < MAX_NUM_SAO_CLASSES; i++) off[i] = offset[i];
  sao->offsetBlock(COMPONENT_Y, type, off, (Pel*)src, (Pel*)res, srcStride, resStride, w, h,
                   (flags & 1) != 0, (flags & 2) != 0, (flags & 4) != 0, (flags & 8) != 0, (flags & 16) != 0, (flags & 32) != 0,
                   (flags & 64) != 0, (flags & 128) != 0);
}

} // extern "C"
