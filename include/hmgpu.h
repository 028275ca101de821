/* include/hmgpu.h -- C ABI of libhmgpu.so, the B200 (sm_100a) implementation of the HM-16.2
 * inter-search hot path.
 *
 * The reference (liron88/HM-16.2) has no plugin/FFI interface: the boundary is cut along its
 * C++ seams and every entry point below names the reference interface it
 * replaces (file:line relative to the reference root).  The HM-side binding that calls these
 * (TEncSearch / TComRdCost shims selected by the GPUME cfg switch) is in
 * hm-16.2_b200/host/ and described in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative HMGPU_E_* code,
 *     hmgpu_last_error() gives the text.  Nothing throws or aborts across the ABI.
 *   - the caller owns every host buffer; the library has finished reading / writing it when
 *     the call returns (blocking semantics -- HM needs the MV before its next line).
 *   - one context per encoder instance, used from that encoder's single thread; several
 *     contexts / processes may share one GPU.
 *   - there is no CPU fallback: without a CUDA device hmgpu_create() fails.
 *   - several encoder PROCESSES on one GPU: start one broker daemon per GPU (hm-16.2_b200/hmgpud, csrc/hmgpud.cu) and set
 *     HMGPU_BROKER=<its socket path> in the encoders' environment.  hmgpu_create() then attaches to the daemon instead of
 *     creating a CUDA context (the process makes no CUDA call at all); searches, picture uploads and prediction costs work
 *     as documented below, the test / measurement entry points that hand out device pointers, streams or profiles return
 *     HMGPU_E_STATE.  Without a reachable daemon hmgpu_create() fails (no private context is created behind the caller's back).
 *   - Pel = int16_t, TCoeff = int32_t, Distortion = uint32_t, MV = 2 x int16_t quarter-pel
 *     (TypeDef.h:692-703, TComMv.h:53-55).
 */
#ifndef HMGPU_H
#define HMGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMGPU_ABI_VERSION 3   /* 2: hmgpu_me_submit / hmgpu_me_wait, hmgpu_predict / hmgpu_pred_error, job.kind
                                 3: hmgpu_pu_submit / hmgpu_pu_wait, hmgpu_set_option, hmgpu_clip_bounds_ctu, broker client mode */

enum
{
  HMGPU_OK = 0,
  HMGPU_E_INVALID = -1,   /* bad argument */
  HMGPU_E_CUDA = -2,      /* CUDA runtime error (text in hmgpu_last_error) */
  HMGPU_E_NOMEM = -3,
  HMGPU_E_STATE = -4      /* e.g. reference slot not uploaded */
};

typedef struct hmgpu_ctx hmgpu_ctx;

/* ------------------------------------------------------------------------------------------
 * Context.  Replaces nothing in the reference (it has no device); created once per encoder
 * next to TEncTop::create (TEncTop.cpp:206-224 wires m_cSearch/m_cRdCost/m_cTrQuant).
 * pic_w/pic_h: luma size as coded (multiple of the min CU size); bit_depth 8..12;
 * max_refs: number of reference slots (DPB size).
 * ------------------------------------------------------------------------------------------ */
int  hmgpu_create(int device, int pic_w, int pic_h, int bit_depth, int max_refs, hmgpu_ctx** out);
void hmgpu_destroy(hmgpu_ctx* ctx);
const char* hmgpu_last_error(const hmgpu_ctx* ctx);   /* ctx may be NULL (creation errors) */
int  hmgpu_abi_version(void);
/* number of kernels launched by this context so far (bench.py "gpu_launches") */
uint64_t hmgpu_launch_count(const hmgpu_ctx* ctx);
/* the CUDA stream all work of this context is issued on (cudaStream_t as void*) */
void* hmgpu_stream(const hmgpu_ctx* ctx);
int  hmgpu_synchronize(hmgpu_ctx* ctx);
/* Tuning knobs of a context (kernel mapping switches, the resident server, tracing).  Their defaults come from the environment
 * (HMGPU_* variables, read once by hmgpu_create); nothing on a launch path reads the environment.  Names: tz_thread,
 * tz_thread_min, tz_merge, tz_carve, tz_p2, frac_v1, frac_overlap, pipe_chunk, pipeline, fastpath, server, server_idle_us,
 * trace, server_stats.  Every setting produces the same results (tests/test_gpu_parity.py switches them). */
int  hmgpu_set_option(hmgpu_ctx* ctx, const char* name, int value);
/* Page-locked host memory.  Buffers passed to hmgpu_me_search / hmgpu_ref_upload / hmgpu_org_upload that
 * come from here (or from the caller's own cudaHostAlloc) are copied directly, without the library's
 * internal staging copy. */
int  hmgpu_host_alloc(hmgpu_ctx* ctx, size_t bytes, void** out);
int  hmgpu_host_free(hmgpu_ctx* ctx, void* p);
/* sizeof of the ABI structs: me_job, me_result, dist_item, mc_job, pred_job (binding self-check) */
void hmgpu_struct_sizes(int out[5]);

/* Per-stage device timing (CUDA events on hmgpu_stream()).  The reference's only timer is
 * clock() (encmain.cpp:95-101, TEncGOP.cpp:646); these are what bench.py's roofline reads.
 * hmgpu_profile_read fills ms[] / launches[] (hmgpu_profile_stage_count() entries each). */
int  hmgpu_profile_enable(hmgpu_ctx* ctx, int on);
int  hmgpu_profile_stage_count(void);
const char* hmgpu_profile_stage_name(int stage);
int  hmgpu_profile_read(hmgpu_ctx* ctx, double* ms, uint64_t* launches, int reset);
/* integer-pipe microbenchmark: which = 0 -> 32-bit add/logic stream, 1 -> VABSDIFF4.U8.ACC;
 * result in giga lane-operations per second (the INT32 roofline denominators) */
int  hmgpu_microbench(hmgpu_ctx* ctx, int which, double* gops);

/* ------------------------------------------------------------------------------------------
 * Reference pictures.
 * hmgpu_ref_upload replaces TComPicYuv::extendPicBorder (TComPicYuv.cpp:171-215, called from
 * TComSlice::setRefPicList, TComSlice.cpp:346,359,372) plus the per-call interpolation of
 * xExtDIFUpSamplingH/Q (TEncSearch.cpp:5565-5766): the reconstructed luma plane is uploaded
 * once, replicate-padded by 80 samples on the device, and all 16 quarter-pel phase planes
 * P[v][h] = clip(V_v(H_h(ref))) are built (IF_INTERNAL_OFFS / shift semantics of
 * TComInterpolationFilter.cpp:166-251).  `luma` points at sample (0,0); stride in samples.
 * Chroma (4:2:0, may be NULL) is padded by 40 and kept for motion compensation.
 * ------------------------------------------------------------------------------------------ */
int hmgpu_ref_upload(hmgpu_ctx* ctx, int slot, const int16_t* luma, int luma_stride,
                     const int16_t* cb, const int16_t* cr, int chroma_stride);
int hmgpu_ref_release(hmgpu_ctx* ctx, int slot);
/* read back one padded phase plane (tests): dst is (pic_w+160) x (pic_h+160) int16 */
int hmgpu_ref_download_plane(hmgpu_ctx* ctx, int slot, int frac_x, int frac_y, int16_t* dst);
/* device-resident variant (no host copy): src is a device pointer to int16 luma */
int hmgpu_ref_upload_device(hmgpu_ctx* ctx, int slot, const void* d_luma, int luma_stride);

/* Original (source) picture of the frame being coded: the key pattern of uni-directional
 * searches (TEncSearch.cpp:3852, pcYuvOrg = the CTU's copy of this picture). */
int hmgpu_org_upload(hmgpu_ctx* ctx, const int16_t* luma, int luma_stride);
int hmgpu_org_upload_device(hmgpu_ctx* ctx, const void* d_luma, int luma_stride);

/* ------------------------------------------------------------------------------------------
 * Motion search.  One job = one TEncSearch::xMotionEstimation call (TEncSearch.cpp:3816-3906)
 * minus its host-side prologue/epilogue: integer search (xPatternSearch :3932-3989 or
 * xTZSearch :4027-4228) followed by xPatternSearchFracDIF (:4386-4422).
 * ------------------------------------------------------------------------------------------ */
enum
{
  HMGPU_F_FEN        = 1 << 0,  /* m_pcEncCfg->getUseFastEnc(): sub-sampled SAD when rows > 8 */
  HMGPU_F_HADME      = 1 << 1,  /* m_pcEncCfg->getUseHADME(): SATD in the fractional search */
  HMGPU_F_LOSSLESS   = 1 << 2,  /* CU transquant bypass: SAD in the fractional search */
  HMGPU_F_HAS_2NX2N  = 1 << 3,  /* pIntegerMv2Nx2NPred != NULL (TEncSearch.cpp:3879-3883) */
  HMGPU_F_FULL       = 1 << 4,  /* xPatternSearch (FastSearch=0 or bi-pred) instead of xTZSearch */
  HMGPU_F_INTEGER    = 1 << 5,  /* run the integer search (else start_x/y IS the integer MV) */
  HMGPU_F_FRAC       = 1 << 6,  /* run xPatternSearchFracDIF after it */
  HMGPU_F_ORG_BLOCK  = 1 << 7   /* key pattern = explicit int16 block (bi-pred 2*org-pred,
                                   TComYuv.cpp:393-424) at org_offset, stride pu_w */
};

/* job.kind.  HMGPU_KIND_SELECTIVE: the integer search is xTZSearchSelective (FastSearch = 2, TEncSearch.cpp:4231-4383); needs
 * HMGPU_F_INTEGER without HMGPU_F_FULL / HMGPU_F_ORG_BLOCK; org_offset then addresses six int16 in org_blocks:
 * m_acMvPredictors[MD_LEFT], [MD_ABOVE], [MD_ABOVE_RIGHT] (hor, ver; quarter-pel) as xPatternSearchFast collects them (:4004-4008). */
enum { HMGPU_KIND_DEFAULT = 0, HMGPU_KIND_SELECTIVE = 1 };

typedef struct hmgpu_me_job
{
  int16_t  pu_x, pu_y;            /* PU origin, luma samples, picture coordinates */
  uint8_t  pu_w, pu_h;            /* multiples of 4 in 4..64 with at most 64 SATD tiles (8x8 tiles iff both are multiples of 8,
                                     else 4x4): every HEVC PU shape; e.g. 36x32 is rejected */
  uint8_t  ref_slot;
  uint8_t  flags;                 /* HMGPU_F_* */
  int16_t  pred_x, pred_y;        /* m_mvPredictor (quarter-pel), TComRdCost::setPredictor */
  int16_t  start_x, start_y;      /* TZ: rcMv on entry (quarter-pel MVP); no INTEGER flag: integer MV */
  int16_t  win_l, win_t, win_r, win_b;  /* cMvSrchRngLT/RB after xSetSearchRange (integer pel) */
  int16_t  i2n_x, i2n_y;          /* m_integerMv2Nx2N[list][ref] (integer pel) */
  int16_t  clip_hmin, clip_hmax, clip_vmin, clip_vmax; /* TComDataCU::clipMv bounds (quarter-pel) */
  int16_t  search_range;          /* m_iSearchRange (adaptive SR), TZ only */
  int16_t  kind;                  /* HMGPU_KIND_*: 0 = the search the flags select */
  uint32_t ui_cost;               /* TComRdCost::m_uiCost after getMotionCost(true,0,..) */
  uint32_t org_offset;            /* HMGPU_F_ORG_BLOCK: element offset into org_blocks */
} hmgpu_me_job;                   /* 48 bytes */

typedef struct hmgpu_me_result
{
  int16_t  int_x, int_y;          /* rcMv after the integer search (integer pel) */
  uint32_t int_sad;               /* ruiSAD of the integer search (MV cost removed) */
  int16_t  half_x, half_y;        /* cMvHalf */
  int16_t  qter_x, qter_y;        /* cMvQter */
  uint32_t frac_cost;             /* ruiCost after xPatternSearchFracDIF */
  uint32_t n_cand;                /* candidates evaluated (integer + 18 sub-pel) */
} hmgpu_me_result;                /* 24 bytes */

/* host buffers in, host buffers out; blocking.  Up to 32 jobs: the resident mailbox server (no kernel launch; one fused launch
 * when the server is switched off); larger batches: the batch kernels, from 65 536 jobs on as a two-lane pipeline whose range
 * checks run on the device.
 * On an error return the contents of `results` are unspecified (a pipelined batch may have been searched in part). */
int hmgpu_me_search(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs,
                    const int16_t* org_blocks, int n_org_elems, hmgpu_me_result* results);
/* Asynchronous pair for small batches (1..32 jobs): hmgpu_me_submit returns once the jobs are visible to the device (the
 * resident mailbox server), the caller does host work that does not need the vectors -- in HM the merge estimation
 * of the PU (TEncSearch::xMergeEstimation, TEncSearch.cpp:2987), which predInterSearch otherwise runs after the searches -- and
 * hmgpu_me_wait blocks until the results are complete.  The buffers passed to hmgpu_me_submit may be reused when it returns.
 * One submit may be outstanding per context; no search or upload call between a submit and its wait. */
int hmgpu_me_submit(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs, const int16_t* org_blocks, int n_org_elems);
int hmgpu_me_wait(hmgpu_ctx* ctx, hmgpu_me_result* results);
/* The searches of a PU together with prediction-error costs that do not depend on them, in ONE mailbox round trip: up to 32
 * searches plus prediction-error jobs (luma MC + SAD / SATD against the source picture, see hmgpu_pred_error), at most 32 job
 * lines in all.  In HM these are the merge candidates of TEncSearch::xMergeEstimation (TEncSearch.cpp:2987-3040, func
 * HMGPU_DF_HADS) and the AMVP candidates of xEstimateMvPredAMVP / xGetTemplateCost (:3571-3637, :3771-3811, HMGPU_DF_SAD) of the
 * same PU.  pred_funcs[i] is the HMGPU_DF_* of pred_jobs[i]; hmgpu_pu_wait fills results[n_jobs] and pred_out[n_pred].
 * hmgpu_me_submit / hmgpu_me_wait are the n_pred = 0 case. */
struct hmgpu_pred_job;   /* defined with hmgpu_predict below */
int hmgpu_pu_submit(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs, const int16_t* org_blocks, int n_org_elems,
                    const struct hmgpu_pred_job* pred_jobs, const uint8_t* pred_funcs, int n_pred);
int hmgpu_pu_wait(hmgpu_ctx* ctx, hmgpu_me_result* results, uint32_t* pred_out);
/* device-resident variant used for kernel-only timing: d_jobs/d_results are device pointers,
 * asynchronous on hmgpu_stream(); call hmgpu_synchronize() to wait.  The host cannot inspect
 * device-resident jobs: flags_any is the OR of the flags of all jobs (selects the kernels to
 * launch) and the jobs must already satisfy the range checks hmgpu_me_search performs. */
int hmgpu_me_search_device(hmgpu_ctx* ctx, const void* d_jobs, int n_jobs,
                           const void* d_org_blocks, void* d_results, int flags_any);

/* Host-side helpers with the reference's exact arithmetic (cheap, scalar):
 * TComDataCU::clipMv bounds (TComDataCU.cpp:2917-2929) and xSetSearchRange
 * (TEncSearch.cpp:3911-3927). */
void hmgpu_clip_bounds(int pic_w, int pic_h, int cu_x, int cu_y, int16_t bounds[4]);   /* g_uiMaxCUWidth = Height = 64 */
void hmgpu_clip_bounds_ctu(int pic_w, int pic_h, int cu_x, int cu_y, int max_cu_w, int max_cu_h, int16_t bounds[4]);
void hmgpu_search_range(const int16_t bounds[4], int pred_x, int pred_y, int srch_rng, int16_t ltrb[4]);

/* ------------------------------------------------------------------------------------------
 * Distortion family, batched.  Replaces DistParam::DistFunc / TComRdCost::m_afpDistortFunc
 * (TComRdCost.h:60-109, TComRdCost.cpp:223-276) evaluated on caller-supplied Pel blocks:
 * one item = one DistFunc(&DistParam) call.
 * ------------------------------------------------------------------------------------------ */
enum
{
  HMGPU_DF_SAD = 0,      /* xGetSAD4..64 / 12/24/48 (TComRdCost.cpp:493-964): honours sub_shift */
  HMGPU_DF_SAD_GENERIC,  /* xGetSAD (:465-491): ignores sub_shift */
  HMGPU_DF_HADS,         /* xGetHADs (:1537-1604) */
  HMGPU_DF_SSE           /* xGetSSE* (:970-1315) */
};

typedef struct hmgpu_dist_item
{
  uint32_t org_offset, cur_offset; /* element offsets into the two Pel arrays */
  int32_t  org_stride, cur_stride;
  uint8_t  w, h;
  uint8_t  func;                   /* HMGPU_DF_* */
  uint8_t  sub_shift;              /* DistParam::iSubShift */
} hmgpu_dist_item;

int hmgpu_dist_batch(hmgpu_ctx* ctx, const int16_t* org, int n_org, const int16_t* cur, int n_cur,
                     const hmgpu_dist_item* items, int n_items, uint32_t* out);

/* ---------------------------------------------------------------------------------------------
 * SAO statistics of one picture component (first half).  Replaces
 * TEncSampleAdaptiveOffset::getStatistics -> getBlkStats (TEncSampleAdaptiveOffset.cpp:312-363, 910-1340)
 * without the pre-deblock sample mode (SAOLcuBoundary 0): per CTU and SAO type (EO_0, EO_90, EO_135, EO_45, BO)
 * the sum of (source - reconstruction) and the sample count per class.  deriveModeNewRDO / deriveModeMergeRDO
 * (rate-distortion with CABAC estimates) stay on the host.
 *   rec / org       the deblocked reconstruction and the source of the component, width x height samples
 *   ctu_w / ctu_h   CTU size in samples of this component (64 for luma, 32 for 4:2:0 chroma)
 *   ctu_flags       per CTU, or NULL for a picture of one slice and one tile: bit 0 left, bit 2 above, bit 4 above-left
 *                   neighbour available (TComPicSym::deriveLoopFilterBoundaryAvailibility); right / below / above-right
 *                   follow from the picture geometry, as getStatistics sets them
 *   skip_r / skip_b m_skipLinesR / m_skipLinesB of the component, per SAO type
 *   stats           [n_ctus][5 types][2: diff, count][32 classes]; EO classes are edgeType + 2 (0..4)
 * ------------------------------------------------------------------------------------------ */
int hmgpu_sao_stats(hmgpu_ctx* ctx, const int16_t* rec, int rec_stride, const int16_t* org, int org_stride, int width, int height,
                    int ctu_w, int ctu_h, const uint8_t* ctu_flags, const int32_t skip_r[5], const int32_t skip_b[5], int64_t* stats);

/* SAO applied to one picture component: TComSampleAdaptiveOffset::offsetCTU -> offsetBlock
 * (TComSampleAdaptiveOffset.cpp:309-620) for every CTU.
 *   rec        the deblocked reconstruction (the reference's m_tempPicYuv copy), width x height samples
 *   types      per CTU: SAO type 0..4 (EO_0, EO_90, EO_135, EO_45, BO) of the component's SAOOffset after
 *              reconstructBlkSAOParams, or -1 for SAO_MODE_OFF
 *   offsets    per CTU 32 values, SAOOffset::offset: the EO offsets of the classes edgeType + 2 (0..4), the BO offsets of
 *              the 32 bands
 *   ctu_flags  per CTU all eight availabilities of deriveLoopFilterBoundaryAvailibility -- bit 0 left, 1 right, 2 above,
 *              3 below, 4 above-left, 5 above-right, 6 below-left, 7 below-right -- or NULL for one slice and one tile
 *   out        width x height samples, stride = width */
int hmgpu_sao_apply(hmgpu_ctx* ctx, const int16_t* rec, int rec_stride, int width, int height, int ctu_w, int ctu_h,
                    const uint8_t* ctu_flags, const int8_t* types, const int32_t* offsets, int16_t* out);

/* Deblocking of one picture: the edge filtering of TComLoopFilter::loopFilterPic (TComLoopFilter.cpp:128-157,
 * xEdgeFilterLuma :530-660, xEdgeFilterChroma :663-790; 4:2:0) -- every vertical edge, then every horizontal edge -- given per
 * 4x4 luma unit of the picture (raster, ceil(w/4) per row) what xDeblockCU derives and what stays on the host:
 *   bs_ver / bs_hor  boundary strength of the unit's left / top edge as in m_aapucBS (only units on the 8-sample grid are
 *                    looked at; xGetBoundaryStrengthSingle, :398-528, needs the CU data of both sides)
 *   qp               TComDataCU::getQP of the unit;   nofilter: IPCM with pcm_loop_filter_disabled, or lossless (:636-645)
 *   beta / tc offsets: slice_beta_offset_div2 / slice_tc_offset_div2; cb / cr: pps_cb_qp_offset / pps_cr_qp_offset
 * y / cb / cr: the reconstruction before deblocking, tight rows (w, w/2, w/2 samples); filtered in place.
 * Bit depths: the context's for luma and chroma. */
int hmgpu_deblock(hmgpu_ctx* ctx, int16_t* y, int16_t* cb, int16_t* cr, int width, int height,
                  const uint8_t* bs_ver, const uint8_t* bs_hor, const int8_t* qp, const uint8_t* nofilter,
                  int beta_offset_div2, int tc_offset_div2, int cb_qp_offset, int cr_qp_offset);

/* ---------------------------------------------------------------------------------------------
 * Intra mode pre-selection of luma PUs.  Replaces the first-pass loop of
 * TEncSearch::estIntraPredQT (TEncSearch.cpp:2352-2395): for every one of the 35 modes
 * TComPrediction::predIntraAng (TComPrediction.cpp:407-492, blocks without DPCM) followed by
 * distParam.DistFunc (HADS, or SADS for transquant-bypass CUs).  One job = one PU; the mode bits
 * (xModeBitsIntra, CABAC state) and the candidate list stay on the host.
 * The reference samples are what initAdiPatternChType (TComPattern.cpp:225-330) leaves in
 * m_piYuvExt[COMPONENT_Y][PRED_BUF_UNFILTERED / PRED_BUF_FILTERED], as ONE LINE of 4N+1 samples
 * each: from the bottom-left neighbour up the left column to the top-left corner (element 2N)
 * and along the row above to the above-right neighbour -- the order that function walks.
 * ------------------------------------------------------------------------------------------ */
enum
{
  HMGPU_IF_ABOVE = 1, HMGPU_IF_LEFT = 2,   /* bAbove / bLeft of predIntraAng (DC value and DC edge filter) */
  HMGPU_IF_EDGE_FILTERS = 4,               /* enableEdgeFilters (:474): off only for RDPCM transquant-bypass CUs */
  HMGPU_IF_SATD = 8,                       /* bUseHadamard (TEncSearch.cpp:2349): DF_HADS, else DF_SADS */
  HMGPU_IF_NO_SMOOTH = 16                  /* sps.getDisableIntraReferenceSmoothing(): always the unfiltered samples */
};

typedef struct hmgpu_intra_job
{
  uint32_t org_offset;   /* element offset of the N x N source block (stride N) in org_blocks */
  uint32_t ref_offset;   /* element offset in ref_lines: 4N+1 unfiltered samples, then 4N+1 filtered samples */
  uint8_t  size;         /* N = 4, 8, 16, 32 or 64 */
  uint8_t  flags;        /* HMGPU_IF_* */
  uint16_t reserved;
} hmgpu_intra_job;       /* 12 bytes */

/* dist[job * 35 + mode] = the distortion TEncSearch.cpp:2373 adds to uiSad for that mode */
int hmgpu_intra_costs(hmgpu_ctx* ctx, const hmgpu_intra_job* jobs, int n_jobs, const int16_t* org_blocks, int n_org_elems,
                      const int16_t* ref_lines, int n_ref_elems, uint32_t* dist);

/* MV rate cost, TComRdCost::getCost(x,y)/getBits (TComRdCost.h:171-188); host-side scalar */
uint32_t hmgpu_mv_bits(int pred_x, int pred_y, int scale, int x, int y);
uint32_t hmgpu_mv_cost(uint32_t ui_cost, int pred_x, int pred_y, int scale, int x, int y);

/* ------------------------------------------------------------------------------------------
 * Motion compensation from the phase planes: TComPrediction::xPredInterBlk
 * (TComPrediction.cpp:660-698) for luma, uni-directional (clipped output).
 * ------------------------------------------------------------------------------------------ */
typedef struct hmgpu_mc_job
{
  int16_t pu_x, pu_y;
  uint8_t pu_w, pu_h;
  uint8_t ref_slot;
  uint8_t reserved;
  int16_t mv_x, mv_y;             /* quarter-pel, already clipped by the caller (clipMv) */
  uint32_t dst_offset;            /* element offset of the w x h output block (stride w) */
} hmgpu_mc_job;

int hmgpu_mc_luma(hmgpu_ctx* ctx, const hmgpu_mc_job* jobs, int n_jobs, int16_t* dst, int n_dst);

/* Motion compensation of a whole PU: TComPrediction::motionCompensation -> xPredInterUni / xPredInterBi ->
 * xPredInterBlk (TComPrediction.cpp:514-698) for luma (quarter-pel, 8 taps) and 4:2:0 chroma (eighth-pel,
 * 4 taps; the chroma planes given to hmgpu_ref_upload), and TComYuv::addAvg (TComYuv.cpp:336-392) when both
 * lists are used (14-bit intermediates).  A list is unused when its ref_slot is negative.  The caller
 * applies xCheckIdenticalMotion (TComPrediction.cpp:497) itself: identical motion is submitted as list 0 only.
 * Output of a job at dst_offset: the w x h luma block, then (with_chroma) the w/2 x h/2 Cb and Cr blocks. */
typedef struct hmgpu_pred_job
{
  int16_t pu_x, pu_y;             /* luma samples, picture coordinates */
  uint8_t pu_w, pu_h;             /* luma size, multiples of 4, 4..64 */
  int8_t  ref_slot[2];            /* per reference list; < 0: list unused */
  int16_t mv_x[2], mv_y[2];       /* quarter-pel luma MVs, already clipped by the caller (clipMv) */
  uint32_t dst_offset;            /* element offset of this job's output */
} hmgpu_pred_job;                 /* 20 bytes */

int hmgpu_predict(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n_jobs, int with_chroma, int16_t* dst, int n_dst);

/* Prediction-error cost of luma predictions against the source picture, batched:
 *   func = HMGPU_DF_HADS: TEncSearch::xGetInterPredictionError (TEncSearch.cpp:2952-2972), the merge-candidate
 *          cost of xMergeEstimation (:2987-3040) -- luma MC (uni or bi) + xGetHADs;
 *   func = HMGPU_DF_SAD_GENERIC / HMGPU_DF_SAD: the distortion of xGetTemplateCost (:3771-3811), the AMVP
 *          candidate cost -- luma MC (uni) + SAD without sub-sampling.
 * The double-precision calcRdCost that follows (TComRdCost.cpp:106) stays on the host.  dst_offset is ignored. */
int hmgpu_pred_error(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n_jobs, int func, uint32_t* out);

/* Merge / skip candidate evaluation, the part that needs no CABAC state: for every merge candidate of a CU
 * (TEncCu::xCheckRDCostMerge2Nx2N, TEncCu.cpp:1406-1528) the motion compensation of the whole CU -- Y, Cb, Cr, uni or bi,
 * TComPrediction::motionCompensation (TComPrediction.cpp:514-714) -- and the distortion of the SKIP reconstruction
 * (reco = pred, TEncSearch::encodeResAndCalcRdInterCU, TEncSearch.cpp:4452-4483) against the source CU: getDistPart = xGetSSE*
 * per component (TComRdCost.cpp:970-1315), BEFORE the chroma weight m_distortionWeight (a double, applied by the host together
 * with the skip-flag / merge-index bits and calcRdCost).  One call costs all candidates of any number of CUs.
 *   jobs[i]        candidate i (its dst_offset: where its prediction lands in pred -- w x h luma, then Cb, Cr of w/2 x h/2)
 *   org_offset[i]  element offset in org_blocks of the source CU of candidate i, in the same layout (Y, Cb, Cr)
 *   sse[i*3 + c]   distortion of component c;  pred may be NULL when only the distortions are wanted. */
int hmgpu_merge_skip_dist(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n_jobs, const uint32_t* org_offset,
                          const int16_t* org_blocks, int n_org_elems, int16_t* pred, int n_pred, uint32_t* sse);

/* ------------------------------------------------------------------------------------------
 * Residual costing: forward core transform and scalar quantiser, batched over TUs.
 * hmgpu_fwd_transform replaces TComTrQuant::xT -> xTrMxN -> partialButterfly4/8/16/32 /
 * fastForwardDst (TComTrQuant.cpp:1805-1827, 836-885, 387-758).  Input: n_tus residual
 * blocks of n x n Pel, each contiguous row-major; output n x n TCoeff each.
 * hmgpu_quant replaces the scalar branch of TComTrQuant::xQuant (TComTrQuant.cpp:1120-1199,
 * flat scaling list, no sign hiding); the RDOQ branch is hmgpu_rdoq below.
 * ------------------------------------------------------------------------------------------ */
int hmgpu_fwd_transform(hmgpu_ctx* ctx, const int16_t* resi, int n_tus, int n, int use_dst,
                        int32_t* coeff);
/* hmgpu_inv_transform replaces TComTrQuant::xIT -> xITrMxN -> partialButterflyInverse4/8/16/32 / fastInverseDst
 * (TComTrQuant.cpp:1830-1850, 894-960, 437-810): n_tus blocks of n x n TCoeff in, n x n residual Pel out (the reconstruction
 * side of the residual-costing loop; hmgpu_dequant and hmgpu_residual_tus below complete it). */
int hmgpu_inv_transform(hmgpu_ctx* ctx, const int32_t* coeff, int n_tus, int n, int use_dst, int16_t* resi);
int hmgpu_quant(hmgpu_ctx* ctx, const int32_t* coeff, int n_tus, int n, int qp_per, int qp_rem,
                int is_intra_slice, int32_t* level, int32_t* delta_u, uint32_t* abs_sum);

/* ------------------------------------------------------------------------------------------
 * Rate-distortion optimised quantisation, batched over TUs.
 * hmgpu_rdoq replaces TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1974-2520; called from TComTrQuant::xQuant :1074-1118 when
 * RDOQ = 1) with xGetCodedLevel (:2660), xGetICRate (:2725), xGetRateLast (:2815), getSigCtxInc (:2548): the level decision per
 * coefficient, the zero-out of coefficient groups, the choice of the last significant position and sign-bit hiding, for square
 * TUs of 4:2:0 pictures without scaling lists / extended precision / Golomb-Rice adaptation.  The CABAC bit estimates stay the
 * host's (TEncSbac::estBit fills estBitsSbacStruct, TComTrQuant.h:59-73, from the coder's state of the moment): the caller
 * passes the tables RDOQ reads, one set per distinct coder state, and every TU names its set.  Costs are IEEE doubles summed in
 * the reference's order; levels and uiAbsSum are bit-identical.
 *   jobs[i]          TU i (coef_offset: where its n x n coefficients / levels start in coef / level, raster, pitch n)
 *   bits[k]          set k of bit estimates, 15-bit fixed point
 *   abs_sum[i]       uiAbsSum of TU i (sum of the magnitudes before sign-bit hiding, as the reference returns it)
 * ------------------------------------------------------------------------------------------ */
typedef struct hmgpu_rdoq_bits {
  int32_t sig_group[2][2];       /* significantCoeffGroupBits */
  int32_t sig[44][2];            /* significantBits: 28 luma contexts, then 16 chroma (getSignificanceMapContextOffset) */
  int32_t last_x[2][10];         /* lastXBits[channel type][group] */
  int32_t last_y[2][10];
  int32_t greater_one[24][2];    /* m_greaterOneBits */
  int32_t level_abs[6][2];       /* m_levelAbsBits */
} hmgpu_rdoq_bits;
#define HMGPU_RDOQ_SIGN_HIDE 1u  /* flags: cu.getSlice()->getPPS()->getSignHideFlag() (and no transquant bypass) */
typedef struct hmgpu_rdoq_job {
  int32_t  log2_size;            /* 2..5 */
  int32_t  channel;              /* 0 luma, 1 chroma */
  int32_t  scan;                 /* 0 diagonal, 1 horizontal, 2 vertical (getTUEntropyCodingParameters) */
  uint32_t flags;
  int32_t  qbits;                /* QUANT_SHIFT + per + transform shift */
  int32_t  qp_per, qp_rem;
  int32_t  go_rice_init;         /* 0 without persistent Rice adaptation */
  int32_t  cbf_bits[2];          /* blockCbpBits / blockRootCbpBits of cbf = 0 / 1 in the context this TU's flag is coded in */
  int32_t  bit_depth;
  int32_t  bits_index;           /* which set of `bits` */
  uint32_t coef_offset;
  uint32_t reserved;
  double   err_scale;            /* getErrScaleCoeffNoScalingList (TComTrQuant.cpp:2043, set at :2954) */
  double   lambda;               /* m_dLambda of the component */
} hmgpu_rdoq_job;
int hmgpu_rdoq(hmgpu_ctx* ctx, const hmgpu_rdoq_job* jobs, int n_jobs, const hmgpu_rdoq_bits* bits, int n_bits,
               const int32_t* coef, int n_coef, int32_t* level, int32_t* abs_sum);

/* hmgpu_dequant replaces TComTrQuant::xDeQuant (TComTrQuant.cpp:1203-1313; called from invTransformNxN :1495) for square TUs with
 * the flat quantiser (no scaling lists), no transform skip, no extended precision: n_tus blocks of n x n levels in, n x n
 * transform coefficients out, one QP (per, rem) for the call.  Bit depth: the context's.  (Transform-skipped TUs are dequantised
 * by the same arithmetic while extended precision is off.) */
int hmgpu_dequant(hmgpu_ctx* ctx, const int32_t* level, int n_tus, int n, int qp_per, int qp_rem, int32_t* coef);

/* ------------------------------------------------------------------------------------------
 * The residual-costing loop for a batch of TUs of one size, without leaving the device.
 * hmgpu_residual_tus does for every TU what TEncSearch::xEstimateResidualQT (TEncSearch.cpp:4680-5310) asks of TComTrQuant per
 * component: transformNxN (TComTrQuant.cpp:1341-1432: xT :1805, then the RDOQ branch of xQuant :1074-1118), invTransformNxN
 * (:1435-1530: xDeQuant :1203, xIT :1830) and the two distortions it then compares (getDistPart with DF_SSE, TComRdCost.cpp:
 * 970-1315): the residual against its reconstruction and against nothing coded.  The bits of the coefficients (CABAC estimate),
 * the chroma distortion weight and calcRdCost stay the host's.
 *   resi        n_tus blocks of n x n residual Pel, contiguous
 *   jobs[i]     the quantiser of TU i as for hmgpu_rdoq; log2_size must say n, coef_offset must be i * n * n, bit_depth the context's
 *   level       n_tus x n x n levels;  abs_sum[i] = uiAbsSum of TU i
 *   rec_resi    n_tus blocks of the reconstructed residual (zeros where nothing was coded)
 *   dist[2i]    SSE(resi, rec_resi) of TU i;  dist[2i + 1] = SSE(resi, 0)
 * ------------------------------------------------------------------------------------------ */
int hmgpu_residual_tus(hmgpu_ctx* ctx, const int16_t* resi, int n_tus, int n, int use_dst, const hmgpu_rdoq_job* jobs,
                       const hmgpu_rdoq_bits* bits, int n_bits, int32_t* level, int32_t* abs_sum, int16_t* rec_resi, uint32_t* dist);

#ifdef __cplusplus
}
#endif
#endif /* HMGPU_H */
