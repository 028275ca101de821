// HmGpuHost.h -- HM-side binding of libhmgpu (the GPUME cfg switch).
//
// This is the host half of the drop-in: C++ that lives inside the HM encoder, keeps the
// TEncSearch / TComRdCost interfaces and turns one TEncSearch::xMotionEstimation call
// (TEncSearch.cpp:3816-3906) into one hmgpu_me_job for the extern "C" library.  It is compiled
// together with the reference sources by hm-16.2_b200/host/Makefile after hm_gpume.patch has
// added the hook lines; nothing here is needed when GPUME=0.
#ifndef HM_GPU_HOST_H
#define HM_GPU_HOST_H

#include "TLibCommon/CommonDef.h"
#include "TLibCommon/TComMv.h"

class TComDataCU;
class TComPic;
class TComPattern;
class TComMvField;
class TComPicYuv;
struct hmgpu_ctx;

struct HmGpuSearchOut
{
  TComMv     mvInt;    ///< rcMv after the integer search (integer pel)
  TComMv     mvHalf;   ///< cMvHalf
  TComMv     mvQter;   ///< cMvQter
  Distortion sadInt;   ///< ruiSAD of the integer search
  Distortion cost;     ///< ruiCost after xPatternSearchFracDIF
};

class HmGpuHost
{
public:
  static HmGpuHost& instance();

  /// Starts the CUDA context / library set-up on a helper thread as soon as the encoder is configured, so that it runs
  /// under the first (intra) picture instead of in front of the first inter search (0.3 - 4 s, box and load dependent).
  Void prewarm      ( Int iPicWidth, Int iPicHeight, Int iBitDepth );

  /// Replaces xPatternSearch / xPatternSearchFast + xPatternSearchFracDIF of one xMotionEstimation call.
  /// piRefY points at the PU origin inside the reference reconstruction (TEncSearch.cpp:3857).
  Void motionSearch( TComDataCU* pcCU, TComPic* pcRefPic, TComPattern* pcPatternKey, Pel* piRefY, Int iRefStride,
                     const TComMv& rcMvSrchRngLT, const TComMv& rcMvSrchRngRB, const TComMv& rcMvPred, const TComMv& rcMvIn,
                     Bool bBi, Bool bFullSearch, Int iSearchRange, Bool bFastEnc, Bool bHADME, Bool bLossless,
                     UInt uiMotionCost, const TComMv* pIntegerMv2Nx2NPred, HmGpuSearchOut& rcOut,
                     const TComMv* pacSelectivePred = 0 );   ///< FastSearch=2: m_acMvPredictors[3] (xPatternSearchFast, TEncSearch.cpp:4004-4008)

  /// Batching of the uni-directional searches of one PU (TEncSearch::predInterSearch, TEncSearch.cpp:3177-3257): the
  /// searches of the different lists / reference pictures of a PU do not depend on each other, so the patched loop
  /// first queues them all (beginQueue, motionSearch x N), sends them to the GPU as ONE hmgpu_me_search call
  /// (flushQueue) and then runs its original body, where motionSearch hands out the queued results.
  Void beginQueue   ();
  Void submitQueue  ();   ///< hands the queued searches to the GPU and returns (hmgpu_me_submit)
  Void waitQueue    ();   ///< collects their results (hmgpu_me_wait); host work in between overlaps the device
  Void flushQueue   () { submitQueue(); waitQueue(); }
  Bool queueing     () const { return m_queueing; }

  /// Merge estimation of a PU (TEncSearch::xMergeEstimation, TEncSearch.cpp:2987-3040) on the device: the prediction error of
  /// every merge candidate -- motionCompensation + luma SATD / SAD, xGetInterPredictionError (:2952-2972) -- travels as a
  /// prediction-error job in the same mailbox call as the PU's motion searches (hmgpu_pu_submit).  mergeOnGpu() is the policy:
  /// small PUs stay on the host, whose work then overlaps the device call (HMGPU_MERGE_MIN_AREA, default 256 luma samples).
  Bool mergeOnGpu   ( TComDataCU* pcCU, Int iWidth, Int iHeight );
  Void queueMergeCand( TComDataCU* pcCU, UInt uiAbsPartIdx, Int iWidth, Int iHeight, const TComMvField& rcMvField0, const TComMvField& rcMvField1, Bool bSatd );
  Distortion mergeCandCost( UInt uiMergeCand ) const { return m_predOut[uiMergeCand]; }

  /// Capture mode (environment HMGPU_CAPTURE=<file>, with GPUME=2): no device is used at all.  Every xMotionEstimation call is
  /// marshalled into its hmgpu_me_job exactly as for the GPU, the CPU search runs, and job + CPU result are appended to the
  /// file together with the pictures the jobs refer to (records: 'H' header, 'R' reference upload, 'O' source picture, 'J'
  /// job).  bench.py / tests replay such a stream through libhmgpu: the encoder's REAL call stream, with the reference's answers.
  Bool capturing    () const { return m_capFile != NULL; }

  /// GPUME=2: compare with what the CPU search just produced; abort on the first mismatch
  Void checkInteger   ( const HmGpuSearchOut& rcOut, const TComMv& rcMvCpu );
  Void checkFractional( const HmGpuSearchOut& rcOut, const TComMv& rcHalfCpu, const TComMv& rcQterCpu, Distortion uiCostCpu );

private:
  HmGpuHost();
  ~HmGpuHost();
  Void xInit        ( TComDataCU* pcCU );
  Int  xRefSlot     ( TComPic* pcRefPic );
  Void xUploadOrg   ( TComDataCU* pcCU );
  Void xFail        ( const char* what );

  static const Int NUM_SLOTS = 16;
  static const Int MAX_QUEUE = 32;   ///< 2 lists x 16 reference pictures
  hmgpu_ctx*  m_ctx;
  void*       m_warmThread;     ///< std::thread of prewarm(), joined by xInit
  hmgpu_ctx*  m_warmCtx;        ///< what it created (NULL on failure: xInit then creates and reports)
  Int         m_warmW, m_warmH, m_warmBitDepth;
  Int         m_picW, m_picH;
  const void* m_slotPic[NUM_SLOTS];
  Int         m_slotPoc[NUM_SLOTS];
  UInt64      m_slotUse[NUM_SLOTS];
  UInt64      m_tick;
  const void* m_orgPic;
  Int         m_orgPoc;
  Pel*        m_keyBlock;
  // queue of the current PU (see beginQueue): jobs as submitted, results once flushed
  Bool        m_queueing;
  Int         m_queueLen;
  Bool        m_queueDone;
  struct hmgpu_me_job*    m_queueJobs;
  struct hmgpu_me_result* m_queueRes;
  Bool        m_queueUsed[MAX_QUEUE];
  Short       m_queueSide[MAX_QUEUE * 6];   ///< MV predictors of queued selective searches (6 per job)
  // merge candidates of the current PU whose prediction errors travel with the queue
  static const Int MAX_PRED = 8;
  struct hmgpu_pred_job*  m_predJobs;
  UChar       m_predFuncs[MAX_PRED];
  UInt        m_predOut[MAX_PRED];
  Int         m_predLen;
  Int         m_mergeMinArea;   ///< HMGPU_MERGE_MIN_AREA
  UInt64      m_mergeCands;     ///< merge candidates costed on the device
  // capture mode
  void*       m_capFile;        ///< FILE*
  struct hmgpu_me_job*    m_capJob;
  struct hmgpu_me_result* m_capRes;
  Int         m_capKeyElems;
  Void        xCapturePicture( char tag, Int slot, Int poc, TComPicYuv* pic );
  // statistics
  UInt64      m_calls, m_cands, m_checked;
  UInt64      m_gpuCalls;       ///< hmgpu_me_search invocations (<= m_calls: queued searches share one)
  Double      m_seconds;        ///< inside hmgpu_me_search
  Double      m_totalSeconds;   ///< inside motionSearch (set-up, uploads, job marshalling included)
  Double      m_initSeconds;    ///< CUDA context + library set-up (once)
  Double      m_uploadSeconds;  ///< reference / source picture uploads
  UInt64      m_uploads;
};

#endif
