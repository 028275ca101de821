// HmGpuHost.cpp -- HM-side binding of libhmgpu (see HmGpuHost.h).
#include "HmGpuHost.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <thread>

#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComSlice.h"
#include "TLibCommon/TComMotionInfo.h"

#include "hmgpu.h"

static Double xNow()
{
  struct timespec t;
  clock_gettime( CLOCK_MONOTONIC, &t );
  return (Double)t.tv_sec + 1e-9 * (Double)t.tv_nsec;
}

HmGpuHost& HmGpuHost::instance()
{
  static HmGpuHost s_host;
  return s_host;
}

HmGpuHost::HmGpuHost()
: m_ctx( NULL ), m_warmThread( NULL ), m_warmCtx( NULL ), m_warmW( 0 ), m_warmH( 0 ), m_warmBitDepth( 0 ), m_picW( 0 ), m_picH( 0 ), m_tick( 0 ), m_orgPic( NULL ), m_orgPoc( -1 ), m_keyBlock( NULL )
, m_queueing( false ), m_queueLen( 0 ), m_queueDone( false ), m_queueJobs( NULL ), m_queueRes( NULL )
, m_predJobs( NULL ), m_predLen( 0 ), m_mergeMinArea( 256 ), m_mergeCands( 0 )
, m_capFile( NULL ), m_capJob( NULL ), m_capRes( NULL ), m_capKeyElems( 0 )
, m_gpuCalls( 0 ), m_calls( 0 ), m_cands( 0 ), m_checked( 0 ), m_seconds( 0.0 ), m_totalSeconds( 0.0 ), m_initSeconds( 0.0 ), m_uploadSeconds( 0.0 ), m_uploads( 0 )
{
  for ( Int i = 0; i < NUM_SLOTS; i++ )
  {
    m_slotPic[i] = NULL; m_slotPoc[i] = -1; m_slotUse[i] = 0;
  }
  const char* cap = getenv( "HMGPU_CAPTURE" );
  if ( cap && *cap )
  {
    m_capFile = fopen( cap, "wb" );
    if ( !m_capFile ) { fprintf( stderr, "[GPUME] cannot open the capture file %s\n", cap ); exit( 1 ); }
    m_capJob = new hmgpu_me_job; m_capRes = new hmgpu_me_result;
  }
}

HmGpuHost::~HmGpuHost()
{
  if ( m_capFile )
  {
    fclose( (FILE*)m_capFile );
    fprintf( stderr, "[GPUME] capture: %llu xMotionEstimation calls written with the CPU search's results, %llu pictures\n", (unsigned long long)m_calls, (unsigned long long)m_uploads );
    delete m_capJob; delete m_capRes;
  }
  if ( m_ctx )
  {
    fprintf( stderr, "[GPUME] %llu xMotionEstimation calls on libhmgpu in %llu GPU calls, %llu candidates, %.3f s in hmgpu_me_search (%.1f us/call, %.3f Mcand/s), "
                     "%.3f s in motionSearch overall, %.3f s waiting for the one-time %s, %llu picture uploads in %.3f s, %llu kernel launches, %llu calls cross-checked\n",
             (unsigned long long)m_calls, (unsigned long long)m_gpuCalls, (unsigned long long)m_cands, m_seconds, m_calls ? m_seconds / (Double)m_calls * 1e6 : 0.0,
             m_seconds > 0 ? (Double)m_cands / m_seconds / 1e6 : 0.0, m_totalSeconds, m_initSeconds,
             ( getenv( "HMGPU_BROKER" ) && *getenv( "HMGPU_BROKER" ) ) ? "attach (through the broker daemon)" : "CUDA set-up",
             (unsigned long long)m_uploads, m_uploadSeconds,
             (unsigned long long)hmgpu_launch_count( m_ctx ), (unsigned long long)m_checked );
    if ( m_mergeCands ) fprintf( stderr, "[GPUME] %llu merge candidates costed on the device (PUs of at least %d luma samples)\n", (unsigned long long)m_mergeCands, m_mergeMinArea );
    hmgpu_destroy( m_ctx );
  }
  delete [] m_predJobs;
  delete [] m_keyBlock;
  delete [] m_queueJobs;
  delete [] m_queueRes;
}

Void HmGpuHost::xFail( const char* what )
{
  // no CPU fallback by design: a GPU error aborts the encode
  fprintf( stderr, "[GPUME] %s: %s\n", what, hmgpu_last_error( m_ctx ) );
  exit( 1 );
}

Void HmGpuHost::xCapturePicture( char tag, Int slot, Int poc, TComPicYuv* pic )
{
  FILE* f = (FILE*)m_capFile;
  const int32_t hdr[2] = { slot, poc };
  fputc( tag, f );
  fwrite( hdr, sizeof( hdr ), 1, f );
  const Pel* p = pic->getAddr( COMPONENT_Y );
  for ( Int y = 0; y < m_picH; y++ ) fwrite( p + (ptrdiff_t)y * pic->getStride( COMPONENT_Y ), sizeof( Pel ), m_picW, f );
}

Void HmGpuHost::prewarm( Int iPicWidth, Int iPicHeight, Int iBitDepth )
{
  if ( m_capFile ) return;
  if ( m_ctx || m_warmThread ) return;
  m_warmW = iPicWidth; m_warmH = iPicHeight; m_warmBitDepth = iBitDepth;
  const char* dev = getenv( "HMGPU_DEVICE" );
  const Int iDev = dev ? atoi( dev ) : 0;
  m_warmThread = new std::thread( [this, iDev]() {
    hmgpu_ctx* c = NULL;
    if ( hmgpu_create( iDev, m_warmW, m_warmH, m_warmBitDepth, NUM_SLOTS, &c ) == HMGPU_OK ) m_warmCtx = c;
  } );
}

Void HmGpuHost::xInit( TComDataCU* pcCU )
{
  if ( m_ctx ) return;
  if ( m_capFile )
  {
    if ( m_keyBlock ) return;
    m_picW = pcCU->getSlice()->getSPS()->getPicWidthInLumaSamples();
    m_picH = pcCU->getSlice()->getSPS()->getPicHeightInLumaSamples();
    m_keyBlock = new Pel[MAX_CU_SIZE * MAX_CU_SIZE];
    const int32_t hdr[4] = { 0x50414348 /* "HCAP" */, m_picW, m_picH, g_bitDepth[CHANNEL_TYPE_LUMA] };
    fputc( 'H', (FILE*)m_capFile );
    fwrite( hdr, sizeof( hdr ), 1, (FILE*)m_capFile );
    return;
  }
  const Double t0 = xNow();
  m_picW = pcCU->getSlice()->getSPS()->getPicWidthInLumaSamples();
  m_picH = pcCU->getSlice()->getSPS()->getPicHeightInLumaSamples();
  // configurations the device path does not model are refused, not silently coded differently (no CPU fallback either):
  // weighted prediction swaps in xGetSADw / xGetHADsw (setWpScalingDistParam, TEncSearch.cpp:5790), and the 80-sample
  // reference margin the kernels rely on is g_uiMaxCUWidth + 16 with CTUs of at most 64
  if ( pcCU->getSlice()->getPPS()->getUseWP() || pcCU->getSlice()->getPPS()->getWPBiPred() )
  {
    fprintf( stderr, "[GPUME] weighted prediction (WeightedPredP / WeightedPredB) is not supported by libhmgpu\n" );
    exit( 1 );
  }
  if ( g_uiMaxCUWidth > 64 || g_uiMaxCUHeight > 64 )
  {
    fprintf( stderr, "[GPUME] MaxCUWidth / MaxCUHeight above 64 is not supported by libhmgpu\n" );
    exit( 1 );
  }
  if ( m_warmThread )
  {
    std::thread* t = (std::thread*)m_warmThread;
    t->join();
    delete t;
    m_warmThread = NULL;
    if ( m_warmCtx && ( m_warmW != m_picW || m_warmH != m_picH || m_warmBitDepth != g_bitDepth[CHANNEL_TYPE_LUMA] ) )
    {
      hmgpu_destroy( m_warmCtx );                          // configured for another picture format: start over
      m_warmCtx = NULL;
    }
    m_ctx = m_warmCtx;
    m_warmCtx = NULL;
  }
  const char* dev = getenv( "HMGPU_DEVICE" );
  if ( !m_ctx && hmgpu_create( dev ? atoi( dev ) : 0, m_picW, m_picH, g_bitDepth[CHANNEL_TYPE_LUMA], NUM_SLOTS, &m_ctx ) != HMGPU_OK )
  {
    xFail( "hmgpu_create" );
  }
  m_keyBlock = new Pel[MAX_CU_SIZE * MAX_CU_SIZE];
  m_queueJobs = new hmgpu_me_job[MAX_QUEUE];
  m_queueRes  = new hmgpu_me_result[MAX_QUEUE];
  m_predJobs  = new hmgpu_pred_job[MAX_PRED];
  const char* mma = getenv( "HMGPU_MERGE_MIN_AREA" );
  if ( mma && *mma ) m_mergeMinArea = atoi( mma );
  m_initSeconds = xNow() - t0;
}

/// reference reconstruction -> device slot (uploaded once: TComSlice.cpp:346 pads it once, too)
Int HmGpuHost::xRefSlot( TComPic* pcRefPic )
{
  const Int poc = pcRefPic->getPOC();
  Int lru = 0;
  for ( Int i = 0; i < NUM_SLOTS; i++ )
  {
    if ( m_slotPic[i] == pcRefPic && m_slotPoc[i] == poc )
    {
      m_slotUse[i] = ++m_tick;
      return i;
    }
    if ( m_slotUse[i] < m_slotUse[lru] ) lru = i;
  }
  const Double t0 = xNow();
  TComPicYuv* rec = pcRefPic->getPicYuvRec();
  if ( m_capFile ) xCapturePicture( 'R', lru, poc, rec );
  else
  if ( hmgpu_ref_upload( m_ctx, lru, rec->getAddr( COMPONENT_Y ), rec->getStride( COMPONENT_Y ), NULL, NULL, 0 ) != HMGPU_OK )
  {
    xFail( "hmgpu_ref_upload" );
  }
  m_slotPic[lru] = pcRefPic; m_slotPoc[lru] = poc; m_slotUse[lru] = ++m_tick;
  m_uploads++; m_uploadSeconds += xNow() - t0;
  return lru;
}

Void HmGpuHost::xUploadOrg( TComDataCU* pcCU )
{
  TComPic* pic = pcCU->getPic();
  if ( m_orgPic == pic && m_orgPoc == pic->getPOC() ) return;
  const Double t0 = xNow();
  TComPicYuv* org = pic->getPicYuvOrg();
  if ( m_capFile ) xCapturePicture( 'O', 0, pic->getPOC(), org );
  else
  if ( hmgpu_org_upload( m_ctx, org->getAddr( COMPONENT_Y ), org->getStride( COMPONENT_Y ) ) != HMGPU_OK )
  {
    xFail( "hmgpu_org_upload" );
  }
  m_orgPic = pic; m_orgPoc = pic->getPOC();
  m_uploads++; m_uploadSeconds += xNow() - t0;
  // a new picture: references of the previous one may have been replaced in the DPB objects
}

Void HmGpuHost::motionSearch( TComDataCU* pcCU, TComPic* pcRefPic, TComPattern* pcPatternKey, Pel* piRefY, Int iRefStride,
                              const TComMv& rcMvSrchRngLT, const TComMv& rcMvSrchRngRB, const TComMv& rcMvPred, const TComMv& rcMvIn,
                              Bool bBi, Bool bFullSearch, Int iSearchRange, Bool bFastEnc, Bool bHADME, Bool bLossless,
                              UInt uiMotionCost, const TComMv* pIntegerMv2Nx2NPred, HmGpuSearchOut& rcOut,
                              const TComMv* pacSelectivePred )
{
  const Double tEnter = xNow();
  xInit( pcCU );
  xUploadOrg( pcCU );

  hmgpu_me_job j;
  memset( &j, 0, sizeof( j ) );
  // PU origin in picture coordinates, recovered from the reference pointer (TEncSearch.cpp:3857)
  TComPicYuv* rec = pcRefPic->getPicYuvRec();
  const ptrdiff_t off = piRefY - rec->getAddr( COMPONENT_Y );
  j.pu_y = (int16_t)( off / iRefStride );
  j.pu_x = (int16_t)( off - (ptrdiff_t)j.pu_y * iRefStride );
  j.pu_w = (uint8_t)pcPatternKey->getROIYWidth();
  j.pu_h = (uint8_t)pcPatternKey->getROIYHeight();
  j.ref_slot = (uint8_t)xRefSlot( pcRefPic );
  const Double t0 = xNow();
  j.pred_x = rcMvPred.getHor(); j.pred_y = rcMvPred.getVer();
  j.win_l = rcMvSrchRngLT.getHor(); j.win_t = rcMvSrchRngLT.getVer();
  j.win_r = rcMvSrchRngRB.getHor(); j.win_b = rcMvSrchRngRB.getVer();
  hmgpu_clip_bounds_ctu( m_picW, m_picH, pcCU->getCUPelX(), pcCU->getCUPelY(), g_uiMaxCUWidth, g_uiMaxCUHeight, &j.clip_hmin );  // TComDataCU::clipMv
  j.search_range = (int16_t)iSearchRange;
  // TComRdCost::m_uiCost exactly as getMotionCost() selected it (m_uiLambdaMotionSAD[0], or [1] for transquant-bypass CUs
  // under COST_MIXED_LOSSLESS_LOSSY_CODING; TComRdCost.h:165), recovered by the caller through getCost(UInt)
  j.ui_cost = uiMotionCost;
  UInt flags = HMGPU_F_INTEGER | HMGPU_F_FRAC;
  if ( bFastEnc )    flags |= HMGPU_F_FEN;
  if ( bHADME )      flags |= HMGPU_F_HADME;
  if ( bLossless )   flags |= HMGPU_F_LOSSLESS;
  if ( bFullSearch ) flags |= HMGPU_F_FULL;
  Int keyElems = 0;
  if ( bBi )
  {
    // bi-prediction: the key pattern is 2*org - otherPred (TComYuv.cpp:393-424), not the source picture
    flags |= HMGPU_F_ORG_BLOCK;
    const Pel* src = pcPatternKey->getROIY();
    const Int  st  = pcPatternKey->getPatternLStride();
    for ( Int y = 0; y < j.pu_h; y++ )
    {
      memcpy( m_keyBlock + y * j.pu_w, src + y * st, sizeof( Pel ) * j.pu_w );
    }
    keyElems = j.pu_w * j.pu_h;
    j.start_x = rcMvIn.getHor(); j.start_y = rcMvIn.getVer();
  }
  else
  {
    j.start_x = rcMvPred.getHor(); j.start_y = rcMvPred.getVer();   // rcMv = *pcMvPred (TEncSearch.cpp:3878)
    if ( pIntegerMv2Nx2NPred )
    {
      flags |= HMGPU_F_HAS_2NX2N;
      j.i2n_x = pIntegerMv2Nx2NPred->getHor(); j.i2n_y = pIntegerMv2Nx2NPred->getVer();
    }
  }
  j.flags = (uint8_t)flags;
  Short side[6] = { 0, 0, 0, 0, 0, 0 };
  if ( pacSelectivePred && !bBi && !bFullSearch )
  {
    // FastSearch=2: xTZSearchSelective; its three spatial MV predictors travel in the side array
    j.kind = HMGPU_KIND_SELECTIVE;
    for ( Int k = 0; k < 3; k++ ) { side[2 * k] = pacSelectivePred[k].getHor(); side[2 * k + 1] = pacSelectivePred[k].getVer(); }
    memcpy( m_keyBlock, side, sizeof( side ) );
    keyElems = 6;
  }

  if ( m_capFile )
  {
    // the CPU search runs next; checkInteger / checkFractional collect its answers and write the record
    *m_capJob = j;
    m_capKeyElems = keyElems;
    memset( m_capRes, 0, sizeof( *m_capRes ) );
    m_calls++;
    return;
  }
  hmgpu_me_result r;
  if ( m_queueing && !bBi )
  {
    // pass 0 of the patched predInterSearch loop: remember the job, the result is handed out in pass 1
    if ( m_queueLen < MAX_QUEUE )
    {
      memcpy( m_queueSide + 6 * m_queueLen, side, sizeof( side ) );
      m_queueJobs[m_queueLen] = j;
      m_queueJobs[m_queueLen].org_offset = 6 * m_queueLen;
      m_queueLen++;
    }
    m_totalSeconds += xNow() - tEnter;
    return;
  }
  Int hit = -1;
  if ( m_queueDone && !bBi )
  {
    for ( Int k = 0; k < m_queueLen; k++ )
    {
      hmgpu_me_job q = m_queueJobs[k];
      q.org_offset = j.org_offset;
      if ( !m_queueUsed[k] && memcmp( &q, &j, sizeof( j ) ) == 0 && memcmp( m_queueSide + 6 * k, side, sizeof( side ) ) == 0 ) { hit = k; break; }
    }
  }
  if ( hit >= 0 )
  {
    r = m_queueRes[hit];
    m_queueUsed[hit] = true;
  }
  else
  {
    if ( hmgpu_me_search( m_ctx, &j, 1, keyElems ? m_keyBlock : NULL, keyElems, &r ) != HMGPU_OK )
    {
      xFail( "hmgpu_me_search" );
    }
    m_gpuCalls++;
  }
  rcOut.mvInt.set ( r.int_x,  r.int_y  );
  rcOut.mvHalf.set( r.half_x, r.half_y );
  rcOut.mvQter.set( r.qter_x, r.qter_y );
  rcOut.sadInt = r.int_sad;
  rcOut.cost   = r.frac_cost;
  m_calls++;
  m_cands   += r.n_cand;
  const Double tExit = xNow();
  m_seconds += tExit - t0;
  m_totalSeconds += tExit - tEnter;
}

Void HmGpuHost::beginQueue()
{
  m_queueing  = true;
  m_queueDone = false;
  m_queueLen  = 0;
  m_predLen   = 0;
}

Bool HmGpuHost::mergeOnGpu( TComDataCU* pcCU, Int iWidth, Int iHeight )
{
  xInit( pcCU );
  return m_mergeMinArea >= 0 && iWidth * iHeight >= m_mergeMinArea;
}

Void HmGpuHost::queueMergeCand( TComDataCU* pcCU, UInt uiAbsPartIdx, Int iWidth, Int iHeight, const TComMvField& rcMvField0, const TComMvField& rcMvField1, Bool bSatd )
{
  xInit( pcCU );
  xUploadOrg( pcCU );
  if ( m_predLen >= MAX_PRED ) xFail( "more merge candidates than MRG_MAX_NUM_CANDS" );
  hmgpu_pred_job& j = m_predJobs[m_predLen];
  memset( &j, 0, sizeof( j ) );
  // PU origin in picture coordinates (as xPredInterBlk addresses it, TComPrediction.cpp:662)
  TComPicYuv* rec = pcCU->getPic()->getPicYuvRec();
  const ptrdiff_t off = rec->getAddr( COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + uiAbsPartIdx ) - rec->getAddr( COMPONENT_Y );
  const Int stride = rec->getStride( COMPONENT_Y );
  j.pu_y = (int16_t)( off / stride );
  j.pu_x = (int16_t)( off - (ptrdiff_t)j.pu_y * stride );
  j.pu_w = (uint8_t)iWidth; j.pu_h = (uint8_t)iHeight;
  const TComMvField* f[2] = { &rcMvField0, &rcMvField1 };
  Int refIdx[2] = { f[0]->getRefIdx(), f[1]->getRefIdx() };
  // TComPrediction::motionCompensation( REF_PIC_LIST_X ) (TComPrediction.cpp:541-549): identical motion in both lists is
  // predicted from list 0 alone (xCheckIdenticalMotion, :497-512)
  if ( pcCU->getSlice()->isInterB() && !pcCU->getSlice()->getPPS()->getWPBiPred() && refIdx[0] >= 0 && refIdx[1] >= 0 &&
       pcCU->getSlice()->getRefPic( REF_PIC_LIST_0, refIdx[0] )->getPOC() == pcCU->getSlice()->getRefPic( REF_PIC_LIST_1, refIdx[1] )->getPOC() &&
       f[0]->getMv() == f[1]->getMv() )
  {
    refIdx[1] = -1;
  }
  for ( Int l = 0; l < 2; l++ )
  {
    j.ref_slot[l] = -1;
    if ( refIdx[l] < 0 ) continue;
    TComMv cMv = f[l]->getMv();
    pcCU->clipMv( cMv );                                   // xPredInterUni (TComPrediction.cpp:592)
    j.ref_slot[l] = (int8_t)xRefSlot( pcCU->getSlice()->getRefPic( l ? REF_PIC_LIST_1 : REF_PIC_LIST_0, refIdx[l] ) );
    j.mv_x[l] = cMv.getHor(); j.mv_y[l] = cMv.getVer();
  }
  m_predFuncs[m_predLen] = (UChar)( bSatd ? HMGPU_DF_HADS : HMGPU_DF_SAD );
  m_predLen++;
  m_mergeCands++;
}

Void HmGpuHost::submitQueue()
{
  m_queueing = false;
  if ( m_queueLen > 0 || m_predLen > 0 )
  {
    const Double t0 = xNow();
    if ( hmgpu_pu_submit( m_ctx, m_queueJobs, m_queueLen, m_queueSide, 6 * m_queueLen, m_predJobs, m_predFuncs, m_predLen ) != HMGPU_OK )
    {
      xFail( "hmgpu_pu_submit" );
    }
    m_gpuCalls++;
    const Double dt = xNow() - t0;
    m_seconds += dt; m_totalSeconds += dt;
  }
}

Void HmGpuHost::waitQueue()
{
  if ( m_queueLen > 0 || m_predLen > 0 )
  {
    const Double t0 = xNow();
    if ( hmgpu_pu_wait( m_ctx, m_queueRes, m_predOut ) != HMGPU_OK )
    {
      xFail( "hmgpu_pu_wait" );
    }
    const Double dt = xNow() - t0;
    m_seconds += dt; m_totalSeconds += dt;
  }
  for ( Int k = 0; k < MAX_QUEUE; k++ ) m_queueUsed[k] = false;
  m_queueDone = true;
}

Void HmGpuHost::checkInteger( const HmGpuSearchOut& rcOut, const TComMv& rcMvCpu )
{
  if ( m_capFile )
  {
    m_capRes->int_x = rcMvCpu.getHor(); m_capRes->int_y = rcMvCpu.getVer();
    return;
  }
  if ( rcOut.mvInt.getHor() != rcMvCpu.getHor() || rcOut.mvInt.getVer() != rcMvCpu.getVer() )
  {
    fprintf( stderr, "[GPUME] MISMATCH at call %llu: integer MV gpu (%d,%d) cpu (%d,%d)\n", (unsigned long long)m_calls,
             rcOut.mvInt.getHor(), rcOut.mvInt.getVer(), rcMvCpu.getHor(), rcMvCpu.getVer() );
    exit( 2 );
  }
}

Void HmGpuHost::checkFractional( const HmGpuSearchOut& rcOut, const TComMv& rcHalfCpu, const TComMv& rcQterCpu, Distortion uiCostCpu )
{
  if ( m_capFile )
  {
    m_capRes->half_x = rcHalfCpu.getHor(); m_capRes->half_y = rcHalfCpu.getVer();
    m_capRes->qter_x = rcQterCpu.getHor(); m_capRes->qter_y = rcQterCpu.getVer();
    m_capRes->frac_cost = uiCostCpu;
    FILE* f = (FILE*)m_capFile;
    const int32_t nKey = m_capKeyElems;
    fputc( 'J', f );
    fwrite( m_capJob, sizeof( *m_capJob ), 1, f );
    fwrite( m_capRes, sizeof( *m_capRes ), 1, f );
    fwrite( &nKey, sizeof( nKey ), 1, f );
    if ( nKey ) fwrite( m_keyBlock, sizeof( Pel ), nKey, f );
    return;
  }
  if ( rcOut.mvHalf.getHor() != rcHalfCpu.getHor() || rcOut.mvHalf.getVer() != rcHalfCpu.getVer() ||
       rcOut.mvQter.getHor() != rcQterCpu.getHor() || rcOut.mvQter.getVer() != rcQterCpu.getVer() || rcOut.cost != uiCostCpu )
  {
    fprintf( stderr, "[GPUME] MISMATCH at call %llu: half gpu (%d,%d) cpu (%d,%d), quarter gpu (%d,%d) cpu (%d,%d), cost gpu %u cpu %u\n",
             (unsigned long long)m_calls, rcOut.mvHalf.getHor(), rcOut.mvHalf.getVer(), rcHalfCpu.getHor(), rcHalfCpu.getVer(),
             rcOut.mvQter.getHor(), rcOut.mvQter.getVer(), rcQterCpu.getHor(), rcQterCpu.getVer(), rcOut.cost, uiCostCpu );
    exit( 2 );
  }
  m_checked++;
}
