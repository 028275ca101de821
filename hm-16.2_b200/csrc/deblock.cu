// deblock.cu -- the edge filtering of the in-loop deblocking filter for one picture (second half).
//
// Replaces the sample work of TComLoopFilter::loopFilterPic (TComLoopFilter.cpp:128-157): xEdgeFilterLuma (:530-660) and
// xEdgeFilterChroma (:663-790) with xPelFilterLuma / xPelFilterChroma / xUseStrongFiltering / xCalcDP / xCalcDQ (:804-922),
// given what xDeblockCU derived per 4x4 luma unit and what stays on the host (the boundary strength needs the CU modes, motion
// and coded-block flags of both sides: xGetBoundaryStrengthSingle, :398-528): bs of the unit's left / top edge, QP, no-filter flag.
//
// The reference walks the CU quadtree CTU by CTU; the edges of one direction lie 8 samples apart, touch 3 samples and read 4 on
// each side, so they are independent: one launch filters every vertical edge of the picture, one thread per 4-sample segment
// (luma, and where the edge is also a chroma edge its two Cb and two Cr samples), a second launch every horizontal edge.
#include "hmgpu_internal.cuh"

__device__ __constant__ uint8_t c_dbk_tc[54] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };
__device__ __constant__ uint8_t c_dbk_beta[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };
__device__ __constant__ uint8_t c_chroma_scale_420[58] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51 };

struct DbkParams
{
  int w, h, bd_luma, bd_chroma, beta_off2, tc_off2, cb_off, cr_off;
};

__device__ __forceinline__ int dbk_clip3(int lo, int hi, int v) { return min(hi, max(lo, v)); }

// DIR = 0: vertical edges (filter across x), 1: horizontal edges
template <int DIR>
__global__ void __launch_bounds__(256)
deblock_edges_kernel(int16_t* __restrict__ y, int16_t* __restrict__ cb, int16_t* __restrict__ cr, const uint8_t* __restrict__ bs_map,
                     const int8_t* __restrict__ qp_map, const uint8_t* __restrict__ nf_map, DbkParams P)
{
  const int uw = (P.w + 3) >> 2, uh = (P.h + 3) >> 2;
  // one thread per unit ON the 8-sample grid of this direction
  const int eu = DIR ? (uh + 1) >> 1 : (uw + 1) >> 1;           // units along the filtered axis that can carry an edge
  const int n = DIR ? eu * uw : uh * eu;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ux, uy;
  if (DIR) { uy = (i / uw) * 2; ux = i - (i / uw) * uw; }
  else { uy = i / eu; ux = (i - uy * eu) * 2; }
  if ((DIR ? uy : ux) == 0) return;                             // no edge at the picture boundary
  const int q = uy * uw + ux, pn = DIR ? q - uw : q - 1;
  const int b = bs_map[q];
  if (!b) return;
  const int qp_p = qp_map[pn], qp_q = qp_map[q];
  const bool nf_p = nf_map[pn] != 0, nf_q = nf_map[q] != 0;
  {
    const int off = DIR ? P.w : 1, step = DIR ? 1 : P.w;
    int16_t* p = y + (size_t)(uy * 4) * P.w + ux * 4;
    const int qp = (qp_p + qp_q + 1) >> 1, scale = 1 << (P.bd_luma - 8), maxv = (1 << P.bd_luma) - 1;
    const int tc = c_dbk_tc[dbk_clip3(0, 53, qp + 2 * (b - 1) + (P.tc_off2 << 1))] * scale;
    const int beta = c_dbk_beta[dbk_clip3(0, 51, qp + (P.beta_off2 << 1))] * scale;
    const int side_thr = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
    int m[4][8];                                                // the segment: 4 lines x (p3 p2 p1 p0 | q0 q1 q2 q3)
#pragma unroll
    for (int l = 0; l < 4; l++)
#pragma unroll
      for (int k = 0; k < 8; k++) m[l][k] = p[l * step + (k - 4) * off];
    const int dp0 = abs(m[0][1] - 2 * m[0][2] + m[0][3]), dq0 = abs(m[0][4] - 2 * m[0][5] + m[0][6]);
    const int dp3 = abs(m[3][1] - 2 * m[3][2] + m[3][3]), dq3 = abs(m[3][4] - 2 * m[3][5] + m[3][6]);
    const int d0 = dp0 + dq0, d3 = dp3 + dq3;
    if (d0 + d3 < beta)
    {
      const bool filt_p = dp0 + dp3 < side_thr, filt_q = dq0 + dq3 < side_thr;
      const bool s0 = abs(m[0][0] - m[0][3]) + abs(m[0][7] - m[0][4]) < (beta >> 3) && 2 * d0 < (beta >> 2) && abs(m[0][3] - m[0][4]) < ((tc * 5 + 1) >> 1);
      const bool s3 = abs(m[3][0] - m[3][3]) + abs(m[3][7] - m[3][4]) < (beta >> 3) && 2 * d3 < (beta >> 2) && abs(m[3][3] - m[3][4]) < ((tc * 5 + 1) >> 1);
      const bool sw = s0 && s3;
#pragma unroll
      for (int l = 0; l < 4; l++)
      {
        const int m0 = m[l][0], m1 = m[l][1], m2 = m[l][2], m3 = m[l][3], m4 = m[l][4], m5 = m[l][5], m6 = m[l][6], m7 = m[l][7];
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
        int16_t* s = p + l * step;
        if (!nf_p) { s[-off] = (int16_t)n3; s[-2 * off] = (int16_t)n2; s[-3 * off] = (int16_t)n1; }
        if (!nf_q) { s[0] = (int16_t)n4; s[off] = (int16_t)n5; s[2 * off] = (int16_t)n6; }
      }
    }
  }
  // chroma (4:2:0): the 8-sample chroma grid = every fourth unit, intra boundaries only
  if (b > 1 && ((DIR ? uy : ux) & 3) == 0)
  {
    const int cw = P.w >> 1;
    const int off = DIR ? cw : 1, step = DIR ? 1 : cw;
    const int maxv = (1 << P.bd_chroma) - 1;
#pragma unroll
    for (int c = 0; c < 2; c++)
    {
      int16_t* pl = c ? cr : cb;
      int iqp = ((qp_p + qp_q + 1) >> 1) + (c ? P.cr_off : P.cb_off);
      if (iqp >= 58) iqp -= 6; else if (iqp >= 0) iqp = c_chroma_scale_420[iqp];
      const int tc = c_dbk_tc[dbk_clip3(0, 53, iqp + 2 * (b - 1) + (P.tc_off2 << 1))] * (1 << (P.bd_chroma - 8));
      int16_t* p0 = pl + (size_t)(uy * 2) * cw + ux * 2;
#pragma unroll
      for (int l = 0; l < 2; l++)
      {
        int16_t* s = p0 + l * step;
        const int m2 = s[-2 * off], m3 = s[-off], m4 = s[0], m5 = s[off];
        const int delta = dbk_clip3(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
        if (!nf_p) s[-off] = (int16_t)dbk_clip3(0, maxv, m3 + delta);
        if (!nf_q) s[0] = (int16_t)dbk_clip3(0, maxv, m4 - delta);
      }
    }
  }
}

int hmgpu_launch_deblock(hmgpu_ctx* ctx, int16_t* d_y, int16_t* d_cb, int16_t* d_cr, int w, int h, int bd_luma, int bd_chroma,
                         const uint8_t* d_bs_ver, const uint8_t* d_bs_hor, const int8_t* d_qp, const uint8_t* d_nf,
                         int beta_off2, int tc_off2, int cb_off, int cr_off)
{
  DbkParams P = { w, h, bd_luma, bd_chroma, beta_off2, tc_off2, cb_off, cr_off };
  const int uw = (w + 3) >> 2, uh = (h + 3) >> 2;
  const int nv = uh * ((uw + 1) >> 1), nh = ((uh + 1) >> 1) * uw;
  HmgpuStage st(ctx, HMGPU_ST_DIST, 2);
  deblock_edges_kernel<0><<<(nv + 255) / 256, 256, 0, ctx->stream>>>(d_y, d_cb, d_cr, d_bs_ver, d_qp, d_nf, P);
  deblock_edges_kernel<1><<<(nh + 255) / 256, 256, 0, ctx->stream>>>(d_y, d_cb, d_cr, d_bs_hor, d_qp, d_nf, P);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
