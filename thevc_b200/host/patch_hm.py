"""Write patched copies of the few reference translation units that get a TLibCuda hook into a build
directory.  The reference sources are read where they lie (never copied into this repository); every
patch is a one-line insertion at a uniquely matching anchor, asserted below.

usage: patch_hm.py <reference root> <output dir>
"""
import os
import re
import sys


def rd(path):
    # bytes preserved one to one (the reference mixes line endings and has a few non-ASCII comment bytes)
    return open(path, encoding="latin-1", newline="").read()


def wr(path, text):
    open(path, "w", encoding="latin-1", newline="").write(text)


def sub_once(text, pattern, repl, name, flags=0):
    new, n = re.subn(pattern, repl, text, count=0, flags=flags)
    assert n == 1, "%s: anchor matched %d times" % (name, n)
    return new


def main():
    ref, out = sys.argv[1], sys.argv[2]
    lib = os.path.join(ref, "source", "Lib")
    os.makedirs(os.path.join(out, "TLibCommon"), exist_ok=True)
    os.makedirs(os.path.join(out, "TLibEncoder"), exist_ok=True)

    # ---- TEncSearch.cpp: xTZSearch and xPatternSearchFracDIF
    s = rd(os.path.join(lib, "TLibEncoder", "TEncSearch.cpp"))
    s = sub_once(s, r'(#include "TEncSearch.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TEncSearch include")
    s = sub_once(s, r'(Void TEncSearch::xTZSearch\( TComDataCU\* pcCU,[^\n]*\r?\n\{\r?\n)',
                 r'\1  if ( tlibcuda_tz_search( pcCU, pcPatternKey, piRefY, iRefStride, pcMvSrchRngLT, pcMvSrchRngRB, rcMv, ruiSAD, m_pcRdCost, m_pcEncCfg, m_iSearchRange ) ) return;\n',
                 "xTZSearch")
    s = sub_once(s, r'(Void TEncSearch::xPatternSearch\( TComPattern\* pcPatternKey, Pel\* piRefY, Int iRefStride, TComMv\* pcMvSrchRngLT, TComMv\* pcMvSrchRngRB, TComMv& rcMv, UInt& ruiSAD \)\r?\n\{\r?\n)',
                 r'\1  if ( tlibcuda_full_search( pcPatternKey, piRefY, iRefStride, pcMvSrchRngLT, pcMvSrchRngRB, rcMv, ruiSAD, m_pcRdCost, m_pcEncCfg ) ) return;\n',
                 "xPatternSearch")
    s = sub_once(s, r'(Void TEncSearch::xPatternSearchFracDIF\(TComDataCU\* pcCU,.*?\n\s*\)\r?\n\{\r?\n)',
                 r'\1  if ( tlibcuda_frac_search( pcPatternKey, piRefY, iRefStride, pcMvInt, rcMvHalf, rcMvQter, ruiCost, m_pcRdCost, m_pcEncCfg, biPred ) ) return;\n',
                 "xPatternSearchFracDIF", flags=re.S)
    # xMergeEstimation / xGetTemplateCost: candidate evaluation (motion compensation + distortion) in one device call
    s = sub_once(s, r'(\n)([ \t]*ruiCost = MAX_UINT;\r?\n[ \t]*for\( UInt uiMergeCand = 0; uiMergeCand < numValidMergeCand; \+\+uiMergeCand \)\r?\n)',
                 r'\1  UInt tvcDist[MRG_MAX_NUM_CANDS];\n  const Bool tvcHave = tlibcuda_merge_costs( pcCU, iPUIdx, cMvFieldNeighbours, uhInterDirNeighbours, numValidMergeCand, m_pcEncCfg->getUseHADME(), tvcDist );\n\2',
                 "xMergeEstimation loop")
    s = sub_once(s, r'(\n[ \t]*)(xGetInterPredictionError\( pcCU, pcYuvOrg, iPUIdx, uiCostCand, m_pcEncCfg->getUseHADME\(\) \);)',
                 r'\1if ( tvcHave ) uiCostCand = tvcDist[uiMergeCand]; else \2', "xMergeEstimation candidate")
    s = sub_once(s, r'(  pcCU->clipMv\( cMvCand \);\r?\n)(\r?\n  // prediction pattern\r?\n)',
                 r'\1  { UInt tvcSad = 0; if ( tlibcuda_template_sad( pcCU, pcCU->getSlice()->getRefPic( eRefPicList, iRefIdx ), uiPartAddr, cMvCand.getHor(), cMvCand.getVer(), iSizeX, iSizeY, tvcSad ) ) return (UInt) m_pcRdCost->calcRdCost( m_auiMVPIdxCost[iMVPIdx][iMVPNum], tvcSad, false, DF_SAD ); }\n\2',
                 "xGetTemplateCost")
    # estIntraPredQT: the 35-mode rough search (predIntraLumaAng + calcHAD per mode) served from one device call per PU
    s = sub_once(s, r'(\n)([ \t]*for\( Int modeIdx = 0; modeIdx < numModesAvailable; modeIdx\+\+ \)\r?\n)',
                 r'\1      UInt tvcSad[35];\n      const Bool tvcHave = tlibcuda_intra_rough( m_piYuvExt, uiWidth, piOrg, uiStride, bAboveAvail, bLeftAvail, tvcSad );\n\2',
                 "estIntraPredQT rough loop")
    s = sub_once(s, r'(\n[ \t]*)(predIntraLumaAng\( pcCU->getPattern\(\), uiMode, piPred, uiStride, uiWidth, uiHeight, pcCU, bAboveAvail, bLeftAvail \);\r?\n\s*\r?\n\s*// use hadamard transform here\r?\n)',
                 r'\1if ( !tvcHave ) \2', "estIntraPredQT predIntraLumaAng")
    s = sub_once(s, r'UInt uiSad = m_pcRdCost->calcHAD\( piOrg, uiStride, piPred, uiStride, uiWidth, uiHeight \);',
                 'UInt uiSad = tvcHave ? tvcSad[uiMode] : m_pcRdCost->calcHAD( piOrg, uiStride, piPred, uiStride, uiWidth, uiHeight ); tlibcuda_intra_note( uiMode, uiSad );',
                 "estIntraPredQT calcHAD")
    wr(os.path.join(out, "TLibEncoder", "TEncSearch.cpp"), s)

    # ---- TEncGOP.cpp: picture-start hook before the slice is compressed
    s = rd(os.path.join(lib, "TLibEncoder", "TEncGOP.cpp"))
    s = sub_once(s, r'(#include "TEncGOP.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TEncGOP include")
    s = sub_once(s, r'(\n)([ \t]*m_pcSliceEncoder->precompressSlice\( pcPic \);\r?\n)',
                 r'\1        tlibcuda_picture_start( pcPic, pcSlice );\n\2', "compressGOP")
    # frame sharding of all-intra sequences: a shard that starts at frame k numbers its pictures from POC k and leaves
    # the parameter sets to shard 0, so that the shards' streams concatenate to the single-run stream
    s = sub_once(s, r'm_bSeqFirst           = true;', 'm_bSeqFirst           = tlibcuda_poc_offset() == 0;', "TEncGOP m_bSeqFirst")
    s = sub_once(s, r'if\(uiPOCCurr>=m_pcCfg->getFrameToBeEncoded\(\)\)', 'if(uiPOCCurr>=m_pcCfg->getFrameToBeEncoded() + tlibcuda_poc_offset())',
                 "compressGOP frame limit")
    # xCalculateAddPSNR: the three sums of squared differences of a picture from one device call
    m = re.search(r'Void TEncGOP::xCalculateAddPSNR\(.*?\n\}\r?\n', s, flags=re.S)
    assert m, "xCalculateAddPSNR not found"
    body = m.group(0)
    n_loops = body.count("for( y = 0; y < iHeight; y++ )")
    assert n_loops == 3, "xCalculateAddPSNR: %d row loops" % n_loops
    first = body.index("for( y = 0; y < iHeight; y++ )")
    body = (body[:first] + "UInt64 tvcSsd[3];\n  const Bool tvcHaveSsd = tlibcuda_pic_ssd( pcPic->getPicYuvOrg(), pcPicD, m_pcEncTop->getPad(0), "
            "m_pcEncTop->getPad(1), tvcSsd );\n  " + body[first:])
    body = body.replace("for( y = 0; y < iHeight; y++ )", "for( y = 0; !tvcHaveSsd && y < iHeight; y++ )")
    body = sub_once(body, r'(\n)([ \t]*unsigned int maxval = 255 \* \(1<<\(g_uiBitDepth \+ g_uiBitIncrement -8\)\);\r?\n)',
                    r'\1  if ( tvcHaveSsd ) { uiSSDY = tvcSsd[0]; uiSSDU = tvcSsd[1]; uiSSDV = tvcSsd[2]; }\n\2', "xCalculateAddPSNR sums")
    s = s[:m.start()] + body + s[m.end():]
    wr(os.path.join(out, "TLibEncoder", "TEncGOP.cpp"), s)

    s = rd(os.path.join(lib, "TLibEncoder", "TEncTop.cpp"))
    s = sub_once(s, r'(#include "TEncTop.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TEncTop include")
    s = sub_once(s, r'm_iPOCLast          = -1;', 'm_iPOCLast          = -1 + (Int)tlibcuda_poc_offset();', "TEncTop m_iPOCLast")
    wr(os.path.join(out, "TLibEncoder", "TEncTop.cpp"), s)
    # the per-picture dQP array is allocated for FramesToBeEncoded entries and indexed by POC (TAppEncCfg.cpp:426, TEncSlice.cpp:246)
    s = rd(os.path.join(lib, "TLibEncoder", "TEncSlice.cpp"))
    s = sub_once(s, r'(#include "TEncSlice.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TEncSlice include")
    s = sub_once(s, r'dQP \+= pdQPs\[ rpcSlice->getPOC\(\) \];', 'dQP += pdQPs[ rpcSlice->getPOC() - (Int)tlibcuda_poc_offset() ];', "initEncSlice dQP")
    wr(os.path.join(out, "TLibEncoder", "TEncSlice.cpp"), s)

    # ---- AnnexBwrite.h: gcc portability (non-const reference to an rvalue)
    s = rd(os.path.join(lib, "TLibEncoder", "AnnexBwrite.h"))
    s = sub_once(s, r'string &P = nalu\.m_nalUnitData\.str\(\);', 'const string &P = nalu.m_nalUnitData.str();', "AnnexBwrite")
    wr(os.path.join(out, "TLibEncoder", "AnnexBwrite.h"), s)

    # ---- TComTrQuant.cpp: xT, xIT, xDeQuant hooks + two MSVC for-scope uses gcc rejects
    s = rd(os.path.join(lib, "TLibCommon", "TComTrQuant.cpp"))
    s = sub_once(s, r'(#include "TComTrQuant.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TComTrQuant include")
    s = sub_once(s, r'for \(Int iCGScanPos = uiCGNum-1;', 'Int iCGScanPos; for (iCGScanPos = uiCGNum-1;', "for-scope 1")
    s = sub_once(s, r'for \( Int scanPos = 0; scanPos < iBestLastIdxP1;', 'Int scanPos; for ( scanPos = 0; scanPos < iBestLastIdxP1;', "for-scope 2")
    s = sub_once(s, r'(Void TComTrQuant::xT\( UInt uiMode, Pel\* piBlkResi, UInt uiStride, Int\* psCoeff, Int iWidth, Int iHeight \)\r?\n\{\r?\n)',
                 r'\1  if ( tlibcuda_xT( uiMode, piBlkResi, uiStride, psCoeff, iWidth, iHeight ) ) return;\n', "xT")
    s = sub_once(s, r'(Void TComTrQuant::xIT\( UInt uiMode, Int\* plCoef, Pel\* pResidual, UInt uiStride, Int iWidth, Int iHeight \)\r?\n\{\r?\n)',
                 r'\1  if ( tlibcuda_xIT( uiMode, plCoef, pResidual, uiStride, iWidth, iHeight ) ) return;\n', "xIT")
    s = sub_once(s, r'(Void TComTrQuant::xDeQuant\( const TCoeff\* pSrc, Int\* pDes, Int iWidth, Int iHeight, Int scalingListType \)\r?\n\{\r?\n)',
                 r'\1  if ( !getUseScalingList() && tlibcuda_xDeQuant( pSrc, pDes, iWidth, iHeight, m_cQP.m_iPer, m_cQP.m_iRem ) ) return;\n', "xDeQuant")
    s = sub_once(s, r'(  Int    iQBits      = m_cQP\.m_iBits;\r?\n  Double dTemp       = 0;)',
                 r'  if ( !getUseScalingList() && tlibcuda_rdoq( pcCU, plSrcCoeff, piDstCoeff, piArlDstCoeff, uiWidth, uiHeight, uiAbsSum, (int)eTType, uiAbsPartIdx, m_cQP.m_iPer, m_cQP.m_iRem, m_dLambda, m_pcEstBitsSbac, m_bUseAdaptQpSelect ) ) return;\n\1',
                 "xRateDistOptQuant")
    s = sub_once(s, r'(Void TComTrQuant::invtransformNxN\( Bool transQuantBypass, TextType eText, UInt uiMode,Pel\* rpcResidual, UInt uiStride, TCoeff\*   pcCoeff, UInt uiWidth, UInt uiHeight,  Int scalingListType, Bool useTransformSkip \)\r?\n\{\r?\n)',
                 r'\1  if ( tlibcuda_defer_itransform( transQuantBypass, (int)eText, rpcResidual, uiStride, pcCoeff, uiWidth, uiHeight, m_cQP.m_iPer, m_cQP.m_iRem, useTransformSkip, getUseScalingList() ) ) return;\n',
                 "invtransformNxN")
    wr(os.path.join(out, "TLibCommon", "TComTrQuant.cpp"), s)
    # ---- TDecCu.cpp / TDecGop.cpp: picture-level batch of the inter reconstruction
    os.makedirs(os.path.join(out, "TLibDecoder"), exist_ok=True)
    s = rd(os.path.join(lib, "TLibDecoder", "TDecCu.cpp"))
    s = sub_once(s, r'(#include "TDecCu.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TDecCu include")
    s = sub_once(s, r'(Void TDecCu::xReconInter\( TComDataCU\* pcCU, UInt uiAbsPartIdx, UInt uiDepth \)\r?\n\{\r?\n)',
                 r'\1  tlibcuda_dec_begin_inter( pcCU, m_ppcYuvResi[uiDepth] );\n', "xReconInter")
    s = sub_once(s, r'(TDecCu::xReconIntraQT\( TComDataCU\* pcCU, UInt uiAbsPartIdx, UInt uiDepth \)\r?\n\{\r?\n)',
                 r'\1  tlibcuda_dec_flush( pcCU->getPic() );\n', "xReconIntraQT")
    wr(os.path.join(out, "TLibDecoder", "TDecCu.cpp"), s)
    s = rd(os.path.join(lib, "TLibDecoder", "TDecGop.cpp"))
    s = sub_once(s, r'(#include "TDecGop.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TDecGop include")
    s = sub_once(s, r'(Void TDecGop::filterPicture\(TComPic\*& rpcPic\)\r?\n\{\r?\n)', r'\1  tlibcuda_dec_flush( rpcPic );\n', "filterPicture")
    wr(os.path.join(out, "TLibDecoder", "TDecGop.cpp"), s)
    # ---- TComLoopFilter.cpp: the deblocking sample work (encoder and decoder)
    s = rd(os.path.join(lib, "TLibCommon", "TComLoopFilter.cpp"))
    s = sub_once(s, r'(#include "TComLoopFilter.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TComLoopFilter include")
    s = sub_once(s, r'(\n  // Horizontal filtering\r?\n  UInt uiCUAddr;\r?\n)', r'\n  tlibcuda_dbk_begin( pcPic );\1', "loopFilterPic begin")
    s = sub_once(s, r'(    xDeblockCU\( pcCU, 0, 0, EDGE_HOR \);\r?\n  \}\r?\n)(\}\r?\n)',
                 r'\1  tlibcuda_dbk_end( pcPic, m_betaOffsetDiv2, m_tcOffsetDiv2 );\n\2', "loopFilterPic end")
    s = sub_once(s, r'(      iQP = \(iQP_P \+ iQP_Q \+ 1\) >> 1;\r?\n)',
                 r'\1      if ( tlibcuda_dbk_unit( pcCU, uiAbsZorderIdx, iDir, iEdge, iIdx, uiBs, iQP, pcCUP, uiPartPIdx, pcCUQ, uiPartQIdx ) ) continue;\n',
                 "xEdgeFilterLuma")
    s = sub_once(s, r'(Void TComLoopFilter::xEdgeFilterChroma\( TComDataCU\* pcCU, UInt uiAbsZorderIdx, UInt uiDepth, Int iDir, Int iEdge \)\r?\n\{\r?\n)',
                 r'\1  if ( tlibcuda_dbk_skip_chroma() ) return;\n', "xEdgeFilterChroma")
    wr(os.path.join(out, "TLibCommon", "TComLoopFilter.cpp"), s)
    # ---- TComSampleAdaptiveOffset.cpp: the SAO sample work (encoder and decoder)
    s = rd(os.path.join(lib, "TLibCommon", "TComSampleAdaptiveOffset.cpp"))
    s = sub_once(s, r'(#include "TComSampleAdaptiveOffset.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TComSampleAdaptiveOffset include")
    s = sub_once(s, r'(  memcpy\(m_pTmpU1, pRec, sizeof\(Pel\)\*picWidthTmp\);\r?\n)',
                 r'  tlibcuda_sao_begin( m_pcPic, yCbCr, m_bUseNIF );\n\1', "processSaoUnitAll begin")
    s = sub_once(s, r'(        )(processSaoCu\(addr, typeIdx, yCbCr\);\r?\n)',
                 r'\1if ( !tlibcuda_sao_unit( addr, typeIdx, m_iOffsetEo, offset ) ) \2', "processSaoUnitAll unit")
    s = sub_once(s, r'(    tmpUSwap = m_pTmpU1;\r?\n    m_pTmpU1 = m_pTmpU2;\r?\n    m_pTmpU2 = tmpUSwap;\r?\n  \}\r?\n)',
                 r'\1  tlibcuda_sao_end( m_pcPic, yCbCr );\n', "processSaoUnitAll end")
    wr(os.path.join(out, "TLibCommon", "TComSampleAdaptiveOffset.cpp"), s)
    # ---- TComPicYuvMD5.cpp: the picture hashes of the digest SEI (encoder) and of its check (decoder)
    s = rd(os.path.join(lib, "TLibCommon", "TComPicYuvMD5.cpp"))
    s = sub_once(s, r'(#include "TComPicYuv.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TComPicYuvMD5 include")
    for fn, method in (("calcCRC", 2), ("calcChecksum", 3), ("calcMD5", 1)):
        s = sub_once(s, r'(void %s\(TComPicYuv& pic, unsigned char digest\[3\]\[16\]\)\r?\n\{\r?\n)' % fn,
                     r'\1  if ( tlibcuda_pic_hash( pic, %d, digest ) ) return;\n' % method, fn)
    wr(os.path.join(out, "TLibCommon", "TComPicYuvMD5.cpp"), s)
    # ---- TComPrediction.cpp: xPredInterUni (shared by encoder and decoder)
    s = rd(os.path.join(lib, "TLibCommon", "TComPrediction.cpp"))
    s = sub_once(s, r'(#include "TComPrediction.h"\r?\n)', r'\1#include "tlibcuda_hm.h"\n', "TComPrediction include")
    s = sub_once(s, r'(Void TComPrediction::xPredInterUni \(.*?pcCU->clipMv\(cMv\);\r?\n)',
                 r'\1  if ( tlibcuda_pred_inter_uni( pcCU, pcCU->getSlice()->getRefPic( eRefPicList, iRefIdx ), uiPartAddr, cMv.getHor(), cMv.getVer(), iWidth, iHeight, rpcYuvPred, bi ) ) return;\n',
                 "xPredInterUni", flags=re.S)
    wr(os.path.join(out, "TLibCommon", "TComPrediction.cpp"), s)
    print("patched 12 files into", out)


if __name__ == "__main__":
    main()
