/* oracle/hm_oracle.h -- TEST INFRASTRUCTURE ONLY (see hm_oracle.c). */
#ifndef HM_ORACLE_H
#define HM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- distortion family (a3, a4, a5) ---- */
uint32_t hmo_sad(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                 int w, int h, int sub_shift, int bit_depth, int generic);
uint32_t hmo_hads(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                  int w, int h, int bit_depth);
uint32_t hmo_sse(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                 int w, int h, int bit_depth);

/* ---- MV rate cost (a6) ---- */
uint32_t hmo_component_bits(int v);
uint32_t hmo_mv_bits(int pred_x, int pred_y, int scale, int x, int y);
uint32_t hmo_mv_cost(uint32_t ui_cost, int pred_x, int pred_y, int scale, int x, int y);
uint32_t hmo_bits_cost(uint32_t ui_cost, uint32_t bits);
uint32_t hmo_lambda_to_cost(double lambda);
double   hmo_calc_rd_cost_sad(double lambda, uint32_t bits, uint32_t dist);

/* ---- window derivation (a8) ---- */
void hmo_clip_bounds(int pic_w, int pic_h, int cu_x, int cu_y, int bounds[4]); /* hmin,hmax,vmin,vmax (qpel) */
void hmo_clip_mv(int pic_w, int pic_h, int cu_x, int cu_y, int mv[2]);
void hmo_set_search_range(int pic_w, int pic_h, int cu_x, int cu_y, int pred_x, int pred_y,
                          int srch_rng, int ltrb[4]);

/* ---- interpolation (a14) and picture padding (a19) ---- */
void hmo_filter_hor(int chroma, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                    int w, int h, int frac, int is_last, int bit_depth);
void hmo_filter_ver(int chroma, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                    int w, int h, int frac, int is_first, int is_last, int bit_depth);
void hmo_extend_border(const int16_t* src, int w, int h, int margin, int16_t* dst);
/* 16 quarter-pel luma planes of a padded picture; planes[(v*4+h)] each pw*ph, valid where the
 * 8-tap support is inside the padded picture (source reads are clamped elsewhere). */
void hmo_phase_planes(const int16_t* padded, int pw, int ph, int bit_depth, int16_t* planes);

/* ---- motion compensation (a15) ---- */
void hmo_pred_inter_blk(int chroma, const int16_t* ref, int ref_stride, int mvx, int mvy,
                        int w, int h, int bi, int bit_depth, int16_t* dst, int dst_stride);
void hmo_add_avg(const int16_t* s0, int st0, const int16_t* s1, int st1, int w, int h,
                 int bit_depth, int16_t* dst, int dst_stride);
void hmo_bipred_key(const int16_t* org, int org_stride, const int16_t* other_pred, int pred_stride,
                    int w, int h, int16_t* dst, int dst_stride);

/* ---- searches (a9, a10, a12/a13, a7) ---- */
typedef struct
{
  /* inputs */
  const int16_t* org; int org_stride; int w, h;      /* key pattern */
  const int16_t* ref; int ref_stride;                /* PU origin inside a padded plane */
  int l, t, r, b;                                    /* integer search window */
  uint32_t ui_cost; int pred_x, pred_y;              /* m_uiCost, predictor (qpel) */
  int fen, hadme, lossless, bit_depth;
  int pic_w, pic_h, cu_x, cu_y, search_range;        /* TZ only (clipMv / re-centre) */
  int start_x, start_y;                              /* TZ start MV (qpel, = the MVP) */
  int has_2nx2n, i2n_x, i2n_y;                       /* TZ optional integer 2Nx2N MV */
  /* outputs */
  int mv_x, mv_y; uint32_t sad;                      /* integer search result */
  int half_x, half_y, qter_x, qter_y; uint32_t frac_cost;
  uint32_t n_cand;                                   /* candidates evaluated */
  /* selective TZ only (FastSearch=2): m_acMvPredictors[MD_LEFT, MD_ABOVE, MD_ABOVE_RIGHT], quarter-pel */
  int sel_pred[3][2];
} hmo_search_t;

void hmo_pattern_search(hmo_search_t* s);
void hmo_tz_search(hmo_search_t* s);
void hmo_tz_selective(hmo_search_t* s);   /* xTZSearchSelective (FastSearch=2) */
void hmo_frac_search(hmo_search_t* s); /* uses s->mv_x/mv_y as integer MV */

/* full xMotionEstimation tail: integer (fs: 0=TZ,1=full) + frac + final cost fix-up */
void hmo_motion_estimation(hmo_search_t* s, int full_search, int bi, uint32_t* io_bits,
                           int out_mv[2], uint32_t* out_cost);

/* ---- forward transform + scalar quantiser (a17, a18) ---- */
void hmo_fwd_transform(int bit_depth, const int32_t* block, int32_t* coeff, int w, int h, int use_dst);
void hmo_transform_matrix(int n, int32_t* m /* n*n */);
/* xITrMxN (TComTrQuant.cpp:894-960), square blocks: coefficients -> residual (TCoeff, clipped to the Pel range) */
void hmo_inv_transform(int bit_depth, const int32_t* coeff, int32_t* block, int n, int use_dst);
uint32_t hmo_quant(const int32_t* coef, int n_coef, int qp_per, int qp_rem, int transform_shift,
                   int is_intra_slice, int32_t* level, int32_t* delta_u);

/* xDeQuant (TComTrQuant.cpp:1203-1313), flat quantiser (no scaling lists), no transform skip / extended precision:
   levels -> transform coefficients of a square TU of 2^log2_size samples a side, clipped to the 16-bit transform range */
void hmo_dequant(const int32_t* level, int n_coef, int log2_size, int qp_per, int qp_rem, int bit_depth, int32_t* coef);

/* ---- intra mode pre-selection: prediction of one luma mode + the 35 distortions of estIntraPredQT's first pass ----
 * line: the 4n+1 reference samples of the PU in the order initAdiPatternChType walks them (TComPattern.cpp:225-330):
 * line[0] = bottom-left neighbour ... line[2n-1] = left neighbour of row 0, line[2n] = top-left corner,
 * line[2n+1] = above sample of column 0 ... line[4n] = above-right.  dst: n x n, stride n. */
void hmo_intra_pred(const int16_t* line, int n, int mode, int bit_depth, int above, int left, int edge_filters, int16_t* dst);
int  hmo_intra_use_filtered(int mode, int n, int no_smooth);
/* flags: HMGPU_IF_* of include/hmgpu.h (1 above, 2 left, 4 edge filters, 8 SATD (else SAD), 16 reference smoothing disabled) */
void hmo_intra_costs(const int16_t* line_unfiltered, const int16_t* line_filtered, const int16_t* org, int n, int bit_depth,
                     int flags, uint32_t dist[35]);

/* ---- SAO statistics (first half): one block of one component, TEncSampleAdaptiveOffset::getBlkStats without the
 * pre-deblock sample mode.  flags: 1 left, 2 right, 4 above, 8 below, 16 above-left, 32 above-right available.
 * diff / count: [type 0..4 = EO_0, EO_90, EO_135, EO_45, BO][class]; EO classes are edgeType + 2 (0..4), BO classes the 32 bands. */
void hmo_sao_blk_stats(const int16_t* src, int src_stride, const int16_t* org, int org_stride, int w, int h, int flags,
                       const int32_t skip_r[5], const int32_t skip_b[5], int bit_depth, int64_t diff[5][32], int64_t count[5][32]);

/* SAO applied to one block: TComSampleAdaptiveOffset::offsetBlock (TComSampleAdaptiveOffset.cpp:309-545).
 * flags: 1 left, 2 right, 4 above, 8 below, 16 above-left, 32 above-right, 64 below-left, 128 below-right available.
 * type 0..4 as above; offset[32]: EO offsets of the classes edgeType + 2 (0..4), BO offsets of the 32 bands.
 * res receives the visited samples only (the reference leaves the others as they are). */
void hmo_sao_offset_block(int type, const int32_t* offset, const int16_t* src, int src_stride, int16_t* res, int res_stride,
                          int w, int h, int flags, int bit_depth);

/* ---- deblocking filter of one picture (second half): the edge filtering of TComLoopFilter::loopFilterPic
 * (TComLoopFilter.cpp:128-157: every vertical edge of the picture, then every horizontal edge) given what xDeblockCU derived
 * per 4x4 luma unit: bs_ver / bs_hor = boundary strength of the unit's LEFT / TOP edge (0 off the 8-sample grid), qp,
 * nofilter (IPCM with pcm_loop_filter_disabled or lossless).  Planes are filtered in place; 4:2:0. */
void hmo_deblock_picture(int16_t* y, int16_t* cb, int16_t* cr, int w, int h, int bit_depth_luma, int bit_depth_chroma,
                         const uint8_t* bs_ver, const uint8_t* bs_hor, const int8_t* qp, const uint8_t* nofilter,
                         int beta_offset_div2, int tc_offset_div2, int cb_qp_offset, int cr_qp_offset);

/* ---- rate-distortion optimised quantisation (f1; oracle/hm_rdoq.c) ---- */
/* the CABAC bit estimates RDOQ reads: estBitsSbacStruct (TComTrQuant.h:59-73) without its cbf tables (the two cbf values that
 * apply to a TU travel in hmo_rdoq_tu); 15-bit fixed point */
typedef struct hmo_rdoq_bits {
  int32_t sig_group[2][2];       /* significantCoeffGroupBits */
  int32_t sig[44][2];            /* significantBits: 28 luma contexts, then 16 chroma */
  int32_t last_x[2][10];         /* lastXBits[channel][group] */
  int32_t last_y[2][10];
  int32_t greater_one[24][2];    /* m_greaterOneBits */
  int32_t level_abs[6][2];       /* m_levelAbsBits */
} hmo_rdoq_bits;
typedef struct hmo_rdoq_tu {
  int32_t log2_size;             /* 2..5 */
  int32_t channel;               /* 0 luma, 1 chroma */
  int32_t scan;                  /* 0 diagonal, 1 horizontal, 2 vertical */
  int32_t sign_hide;
  int32_t qbits;                 /* QUANT_SHIFT + per + transform shift */
  int32_t qp_per, qp_rem;
  int32_t go_rice_init;
  int32_t cbf_bits[2];           /* bits of cbf = 0 / cbf = 1 in the context the TU's (root) cbf is coded in */
  int32_t bit_depth;
  double  err_scale;             /* getErrScaleCoeffNoScalingList */
  double  lambda;                /* m_dLambda */
} hmo_rdoq_tu;
/* coef, level: raster, pitch = width; returns uiAbsSum */
int hmo_rdoq(const hmo_rdoq_tu* tu, const hmo_rdoq_bits* bits, const int32_t* coef, int32_t* level);
long long hmo_rdoq_batch(const hmo_rdoq_tu* tus, const int32_t* bits_index, const uint32_t* coef_offset, int n_tus,
                         const hmo_rdoq_bits* bits, const int32_t* coef, int32_t* level);
/* scan[k] = raster position of the k-th coefficient (SCAN_GROUPED_4x4), scan_cg[i] = raster index of the i-th coefficient group */
void hmo_scan_order(int log2_size, int scan_type, uint16_t* scan, uint16_t* scan_cg);

#ifdef __cplusplus
}
#endif
#endif
