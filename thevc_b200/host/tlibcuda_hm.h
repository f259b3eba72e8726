/*
 * tlibcuda_hm.h -- hook points that bind HM 7.2 (fr34k8/thevc) to libthevc_cuda.so.
 *
 * This header and tlibcuda_hm.cpp are the "source/Lib/TLibCuda" shim of INTEGRATION.md.  They are
 * compiled together with the reference's own sources (taken where they lie, patched on the fly by
 * patch_hm.py; see Makefile) into build/hm/TAppEncoderCuda.  Each hook replaces the body of one
 * reference member and returns true when the CUDA path produced the result; it returns false only
 * for cases this round does not port (bi-prediction search, scaling lists), in which case the
 * reference's own code below the hook runs -- that is the reference's implementation, not a CPU
 * fallback of ours.  A CUDA error is fatal (exit(EXIT_FAILURE)), like the reference's own errors.
 *
 * Which hooks are active is chosen with the environment variable TVC_HM: a comma list of whole tokens.  Unset = the per-call
 * hooks me,frac,tq,rdoq,mc,tables; opt-in: frame, candgrid / cand, dbk, sao, intra[N], hash, batch (decoder), verify, verbose,
 * nobipred (bi-prediction refinement back on the reference's code); "none" runs the unmodified path; an unknown token is fatal.
 */
#ifndef TLIBCUDA_HM_H
#define TLIBCUDA_HM_H

class TComDataCU;
class TComPattern;
class TComMv;
class TComRdCost;
class TEncCfg;
class TComPic;
class TComSlice;
class TComYuv;

/* TEncGOP::compressGOP, before compressSlice (TEncGOP.cpp:576): upload the current original and every
 * reference reconstruction (final + border-extended at this point), run the SAD-table pre-pass */
void tlibcuda_picture_start(TComPic* pic, TComSlice* slice);

/* TEncSearch::xTZSearch (TEncSearch.cpp:4302) */
bool tlibcuda_tz_search(TComDataCU* cu, TComPattern* key, short* refY, int refStride, TComMv* lt, TComMv* rb, TComMv& rcMv,
                        unsigned& ruiSAD, TComRdCost* rd, TEncCfg* cfg, int searchRange);
/* TEncSearch::xPatternSearch (TEncSearch.cpp:4227): the exhaustive search of the bi-prediction refinement (xMotionEstimation with bBi:
 * the pattern is 2 * org - pred of the other list, the window +-bipredSearchRange).  TVC_HM=...,bipred: integer and fractional stage
 * of that refinement in one device call (tvc_me_bipred); the xPatternSearchFracDIF call that follows takes its result from it. */
bool tlibcuda_full_search(TComPattern* key, short* refY, int refStride, TComMv* lt, TComMv* rb, TComMv& rcMv, unsigned& ruiSAD,
                          TComRdCost* rd, TEncCfg* cfg);
/* TEncSearch::xPatternSearchFracDIF (TEncSearch.cpp:4476) */
bool tlibcuda_frac_search(TComPattern* key, short* refY, int refStride, TComMv* mvInt, TComMv& half, TComMv& qter,
                          unsigned& ruiCost, TComRdCost* rd, TEncCfg* cfg, bool biPred);
/* TComPrediction::xPredInterUni (TComPrediction.cpp:483-490), after clipMv: luma + chroma prediction of one PU
 * from one reference picture into the TComYuv (bi: 14-bit intermediates for addAvg) */
bool tlibcuda_pred_inter_uni(TComDataCU* cu, TComPic* refPic, unsigned partAddr, int mvx, int mvy, int w, int h,
                             TComYuv* dst, bool bi);
/* TComTrQuant::xT / xIT / xDeQuant (TComTrQuant.cpp:1542, 1583, 1272) */
bool tlibcuda_xT(unsigned mode, short* resi, unsigned stride, int* coef, int w, int h);
bool tlibcuda_xIT(unsigned mode, int* coef, short* resi, unsigned stride, int w, int h);
bool tlibcuda_xDeQuant(const int* src, int* dst, int w, int h, int per, int rem);
/* ---- decoder, picture-level batch (TVC_HM=...,batch): inter CUs are not reconstructed one by one.  TDecCu::xReconInter
 * (TDecCu.cpp:448-466) announces the CU; the reference's own motionCompensation / xDecodeInterTexture then run with their
 * leaves recording instead of computing (xPredInterUni -> one tvc_pu, invtransformNxN -> one tvc_tu + its levels); the
 * batch is executed -- one tvc_mc_batch + one tvc_inv_tq_batch for everything pending -- when an intra CU needs its
 * neighbours (TDecCu::xReconIntraQT, :689) and before the in-loop filters (TDecGop::filterPicture, TDecGop.cpp:201), and
 * the reconstructed CUs are copied into the picture. */
class TComPic;
bool tlibcuda_dec_begin_inter(TComDataCU* cu, TComYuv* resi);
void tlibcuda_dec_flush(TComPic* pic);
/* TComTrQuant::invtransformNxN (TComTrQuant.cpp:1428): true = recorded for the picture batch, nothing to do */
bool tlibcuda_defer_itransform(bool bypass, int ttype, short* resi, unsigned stride, int* coeff, unsigned w, unsigned h, int per, int rem,
                               bool transformSkip, bool scalingList);

/* ---- deblocking (TVC_HM=...,dbk): TComLoopFilter::loopFilterPic (TComLoopFilter.cpp:153) keeps deriving the boundary
 * strengths on the host; xEdgeFilterLuma (:571) hands every edge unit with bs != 0 to tlibcuda_dbk_unit instead of
 * filtering it, xEdgeFilterChroma (:680) is skipped (the device derives the chroma edges from the same records), and
 * tlibcuda_dbk_end runs tvc_deblock_pic on the picture.  TVC_HM=dbkdump leaves the filtering to the reference and writes
 * the picture before / after together with the records (the golden vectors of tests/golden/deblock_golden.npz). */
void tlibcuda_dbk_begin(TComPic* pic);
bool tlibcuda_dbk_unit(TComDataCU* cu, unsigned absZorderIdx, int dir, int edge, unsigned idx, unsigned bs, int qp, TComDataCU* cuP,
                       unsigned partP, TComDataCU* cuQ, unsigned partQ);
bool tlibcuda_dbk_skip_chroma();
void tlibcuda_dbk_end(TComPic* pic, int betaOffsetDiv2, int tcOffsetDiv2);

/* ---- SAO apply (TVC_HM=...,sao): TComSampleAdaptiveOffset::processSaoUnitAll (TComSampleAdaptiveOffset.cpp:1072) keeps
 * walking the CTUs and resolving merges; instead of processSaoCu every CTU's type and offset tables are recorded, and
 * tlibcuda_sao_end runs tvc_sao_plane for the component.  Single-slice pictures only (m_bUseNIF == false), otherwise the
 * reference code runs.  TVC_HM=saodump: the reference filters, planes before / after + records are written. */
void tlibcuda_sao_begin(TComPic* pic, int yCbCr, bool useNIF);
bool tlibcuda_sao_unit(int addr, int typeIdx, const int* offsetEo, const int* offsetBands);
void tlibcuda_sao_end(TComPic* pic, int yCbCr);

/* ---- intra rough search (TVC_HM=...,intra[N]): TEncSearch::estIntraPredQT (TEncSearch.cpp:2530-2543).  Before the 35-mode
 * loop the hook hands the reference samples initAdiPattern left in m_piYuvExt (unfiltered; column 0 and row 0 of the
 * (2N+1)^2 array) and the original block to tvc_intra_rough and keeps the 35 SATDs; inside the loop predIntraLumaAng +
 * calcHAD are skipped and uiSad is read from that array, the mode bits, the cost and the candidate list stay the
 * reference's.  PUs narrower than N (default 16: below that one launch costs more than the host's 35 predictions) are
 * left to the reference's code.  TVC_HM=intradump: nothing runs on the device, the reference's own uiSad values are
 * written together with the line and the block (the golden vectors of tests/golden/intra_rough.npz). */
bool tlibcuda_intra_rough(const int* adiBuf, unsigned width, const short* org, unsigned orgStride, bool above, bool left, unsigned* sad35);
void tlibcuda_intra_note(unsigned mode, unsigned sad);

/* ---- merge / AMVP candidate evaluation (TVC_HM=...,cand).  TEncSearch::xMergeEstimation (TEncSearch.cpp:3096-3149):
 * before the candidate loop all candidates' motions go to one tvc_pred_cost_batch (luma motion compensation + HAD / SAD against
 * the original, = xGetInterPredictionError per candidate); inside the loop uiCostCand is read from the result, the merge-index
 * bits and the strict "<" stay.  TEncSearch::xGetTemplateCost (:4057-4118): the prediction + SAD of one AMVP candidate, the
 * m_auiMVPIdxCost term stays.  Weighted prediction: reference path. */
class TComMvField;
bool tlibcuda_merge_costs(TComDataCU* cu, int puIdx, TComMvField* cands, const unsigned char* interDir, int numCand, bool hadamard,
                          unsigned* dist);
bool tlibcuda_template_sad(TComDataCU* cu, TComPic* refPic, unsigned partAddr, int mvx, int mvy, int w, int h, unsigned& sad);

/* ---- picture hash (TVC_HM=...,hash): calcMD5 / calcCRC / calcChecksum (TComPicYuvMD5.cpp:119-200) of the digest SEI
 * (TEncGOP.cpp:1150-1172) and of the decoder's check (TDecGop.cpp:340-370): the picture goes into a device slot and tvc_pic_hash
 * returns the reference's digest[3][16]. */
class TComPicYuv;
bool tlibcuda_pic_hash(TComPicYuv& pic, int method, unsigned char digest[3][16]);

/* TEncGOP::xCalculateAddPSNR (TEncGOP.cpp:1582-1641; TVC_HM=...,psnr): the three UInt64 sums of squared differences between the original
 * and the final reconstruction from one device call (tvc_pic_ssd); the PSNR arithmetic in doubles stays the reference's.  Pictures
 * with conformance padding are left to the reference's loops. */
bool tlibcuda_pic_ssd(TComPicYuv* org, TComPicYuv* rec, int padx, int pady, unsigned long long ssd[3]);

/* ---- frame sharding of all-intra sequences (thevc_b200/host/shard_encode.py).  TVC_POC_OFFSET=k: this process
 * encodes the frames from input frame k on (-fs k) as POC k, k+1, ... (TEncTop::m_iPOCLast starts at k-1, TEncTop.cpp:54; the
 * frame limit of compressGOP, TEncGOP.cpp:211, moves with it) and, for k > 0, writes no VPS/SPS/PPS (m_bSeqFirst,
 * TEncGOP.cpp:86,680-712): with IntraPeriod 1 every picture is coded on its own, so the shards' NAL units are the single
 * run's and the host only concatenates them.  No CUDA involved. */
unsigned tlibcuda_poc_offset();

/* TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1719): est is m_pcEstBitsSbac (estBitsSbacStruct == tvc_est_bits) */
bool tlibcuda_rdoq(TComDataCU* cu, int* src, int* dst, int* arl, unsigned w, unsigned h, unsigned& absSum, int ttype,
                   unsigned absPartIdx, int per, int rem, double lambda, const void* est, bool useArl);

#endif
