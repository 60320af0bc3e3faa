// mp_stages.cu -- stages S2 (single-end DP on both mates of every pair the deep DP left unaligned) and
// S3 ("default DP": mate rescue next to a single-end hit) of soap3_dp_pair_align (alignment.cpp:165-275).
//
//   S2  DPForUnalignSingle2 -> SingleDPWrapper::transferSeed      DV-DPForSingleReads.cpp:121-215
//       SingleEndSeedingEngine::singleMerge                        DV-DPfunctions.cpp:295-342
//       SingleEndAlgnBatch::pack                                   DV-DPfunctions.cpp:401-445
//       SingleDP_Space::algnmtCPUThread (result assembly)          DV-DPfunctions.cpp:678-750
//   S3  semiGlobalDPForSingleEndDpAlignment                        DV-SemiDP.cpp:197-283, 165-181
//       HalfEndOccStream::fetchNextSingleAlgnResult                DV-DPfunctions.cpp:1048-1079
//       HalfEndAlgnBatch::pack                                     DV-DPfunctions.cpp:1151-1231
//       DP_Space::algnmtCPUThread + DPOutputThread                 DV-DPfunctions.cpp:1476-1747
//
// These stages see only the pairs stage S1 could not place (a few percent of a batch).  The DP itself
// (ref-window extraction from the HBM text, fill, traceback) runs in the same kernels as S1; the list
// plumbing around it (seed thinning, the reference's sort orders, per-pair grouping) is host code that
// follows the reference's own containers and std::sort calls so that tie orders match.
#include "mp_context.h"
#include "mp_cigar.h"
#include <algorithm>
#include <tuple>
#include <string.h>

namespace {

struct SCand { uint32_t readID; uint32_t strand; uint64_t pos; uint32_t seedLen; };

// RadixTraitsCandidateInfo (DV-DPForSingleReads.cpp:109-119)
inline bool scand_less(const SCand &x, const SCand &y)
{
    return std::make_tuple(x.readID, x.strand, x.pos, x.seedLen) < std::make_tuple(y.readID, y.strand, y.pos, y.seedLen);
}

// singleMerge (DV-DPfunctions.cpp:295-342); `c` ends with the 0x7FFFFFFF sentinel
void single_merge(const std::vector<SCand> &c, std::vector<SCand> &out)
{
    const SCand *p = c.data();
    while (p->readID != 0x7FFFFFFFu) {
        uint32_t readID = p->readID;
        size_t oldSize = out.size();
        for (; p->readID == readID; p++) {
            if (p->seedLen < 17) continue;
            out.push_back(*p);
            while ((p + 1)->readID == readID) {
                if ((p + 1)->pos < out.back().pos + 5 && out.back().strand == (p + 1)->strand) {   // DPS_DIVIDE_GAP
                    if ((p + 1)->seedLen > out.back().seedLen) out.back() = *(p + 1);
                } else break;
                ++p;
            }
        }
        std::sort(out.begin() + oldSize, out.end(), [](const SCand &a, const SCand &b) { return a.seedLen > b.seedLen; });
        if (oldSize < out.size())
            while (out.back().seedLen < out[oldSize].seedLen * 0.6) out.pop_back();
    }
}

// encode one pattern -> cigar text appended to the arena; returns offset; fills stats
// (a failed growth of the pinned arena returns 0xFFFFFFFF and leaves the error message set)
uint32_t append_cigar(PinnedBuf<char> &arena, const uint8_t *pat, int open, int ext, CigStats &st)
{
    st = cigar_encode(pat, open, ext, nullptr, 0);
    size_t off = arena.size();
    if (arena.resize(off + st.textLen + 1)) return 0xFFFFFFFFu;
    cigar_encode(pat, open, ext, arena.data() + off, st.textLen);
    arena[off + st.textLen] = 0;
    return (uint32_t)off;
}

}  // namespace

// seeds (SeedPos entries, DV-DPfunctions.cpp:2555-2594 / SeedPool.cpp:191-207) of every read whose pair the deep DP
// did not place, appended in arbitrary order (the host sorts them as transferSeed does)
struct GatherRec { uint64_t pos; uint32_t readID; uint32_t strand_len; };
__global__ void k_gather_unplaced(const uint8_t *__restrict__ alignedPair, uint32_t nReads, const uint32_t *__restrict__ hitStart,
                                  const uint32_t *__restrict__ nPos, const uint32_t *__restrict__ nNeg, const mp_seed_pos *__restrict__ sp,
                                  GatherRec *__restrict__ out, uint32_t cap, unsigned int *__restrict__ cursor, unsigned int *__restrict__ nUnplacedPairs)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nReads || alignedPair[r >> 1]) return;
    if ((r & 1) == 0) atomicAdd(nUnplacedPairs, 1u);
    const uint32_t c = nPos[r] + nNeg[r];
    if (c == 0) return;
    const uint32_t base = atomicAdd(cursor, c);
    const mp_seed_pos *src = sp + hitStart[r];
    for (uint32_t a = 0; a < c; ++a)
        if (base + a < cap) {
            GatherRec g; g.pos = src[a].pos; g.readID = r; g.strand_len = (src[a].strand_readID & 0x80000000u) | (src[a].paired_seedLength & 0x7FFFFFFFu);
            out[base + a] = g;
        }
}

// host task list -> DP on the device -> host outputs (chunked)
int mpd_run_host_tasks(mp_context *ctx, const std::vector<MpDpTask> &tasks, uint32_t maxRefLen, uint32_t maxReadLen, const MpDpParams &P,
                       std::vector<MpDpOut> &outs, std::vector<uint8_t> &pats, uint32_t patStride)
{
    const size_t n = tasks.size();
    outs.resize(n); pats.assign(n * (size_t)patStride, 0);
    const size_t CH = 1u << 17;
    for (size_t base = 0; base < n; base += CH) {
        uint32_t m = (uint32_t)std::min(CH, n - base);
        if (ctx->dTasks.reserve((size_t)m * sizeof(MpDpTask)) || ctx->dDpOut.reserve((size_t)m * sizeof(MpDpOut)) ||
            ctx->dPattern.reserve((size_t)m * patStride)) return MP_ERR_CUDA;
        MP_CUDA(cudaMemcpyAsync(ctx->dTasks.p, tasks.data() + base, (size_t)m * sizeof(MpDpTask), cudaMemcpyHostToDevice, ctx->stream));
        if (int rc = mpd_run_tasks(ctx, ctx->dTasks.as<MpDpTask>(), m, maxRefLen, maxReadLen, P, ctx->dDpOut.as<MpDpOut>(),
                                   ctx->dPattern.as<uint8_t>(), patStride)) return rc;
        MP_CUDA(cudaMemcpyAsync(outs.data() + base, ctx->dDpOut.p, (size_t)m * sizeof(MpDpOut), cudaMemcpyDeviceToHost, ctx->stream));
        MP_CUDA(cudaMemcpyAsync(pats.data() + base * patStride, ctx->dPattern.p, (size_t)m * patStride, cudaMemcpyDeviceToHost, ctx->stream));
        MP_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int mps_single_and_rescue(mp_context *ctx, const mp_align_params *P, mp_results *out, uint64_t &cells, uint64_t &tasksRun)
{
    const uint32_t nReads = ctx->nReads, nPairs = nReads / 2;
    const uint64_t fullLen = ctx->ix.n;
    // ---- reads of pairs without a deep-DP result: their seeds are gathered on the device ----
    cudaStream_t st = ctx->stream;
    if (P->softClipLeft != P->softClipRight) {
        mp_set_error("single-end / default DP need MaxFrontLenClipped == MaxEndLenClipped (the reference applies task 0's clip sizes to a whole batch, CPU_DPfunctions.cpp:300)");
        return MP_ERR_ARG;
    }
    const std::vector<uint32_t> &lens = ctx->hLens;
    MpTrace tr;
    std::vector<SCand> cand;
    {
        unsigned int *dCur = (unsigned int *)(ctx->dCounters.as<unsigned long long>() + 13);       // [13]: cursor, unplaced pairs
        size_t cap = std::max<size_t>(ctx->dGather.cap / sizeof(GatherRec), (size_t)1 << 16);
        unsigned int hcur[2] = { 0, 0 };
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (ctx->dGather.reserve(cap * sizeof(GatherRec))) return MP_ERR_CUDA;
            MP_CUDA(cudaMemsetAsync(dCur, 0, 8, st));
            (++g_mp_launches), k_gather_unplaced<<<(nReads + 255) / 256, 256, 0, st>>>(ctx->dAligned.as<uint8_t>(), nReads, ctx->dHitStart.as<uint32_t>(),
                ctx->dNPos.as<uint32_t>(), ctx->dNNeg.as<uint32_t>(), ctx->dSeedPos.as<mp_seed_pos>(), ctx->dGather.as<GatherRec>(), (uint32_t)cap, dCur, dCur + 1);
            MP_CUDA(cudaGetLastError());
            MP_CUDA(cudaMemcpyAsync(hcur, dCur, 8, cudaMemcpyDeviceToHost, st));
            MP_CUDA(cudaStreamSynchronize(st));
            if (hcur[0] <= cap) break;
            cap = (size_t)hcur[0] + 1024;
        }
        if (hcur[1] == 0) return 0;                       // every pair was placed by the deep DP
        std::vector<GatherRec> g(hcur[0]);
        if (hcur[0]) MP_CUDA(cudaMemcpy(g.data(), ctx->dGather.p, (size_t)hcur[0] * sizeof(GatherRec), cudaMemcpyDeviceToHost));
        cand.resize(g.size());
        for (size_t i = 0; i < g.size(); ++i) {
            SCand c; c.readID = g[i].readID; c.strand = (g[i].strand_len >> 31) + 1; c.pos = g[i].pos; c.seedLen = g[i].strand_len & 0x7FFFFFFFu;
            cand[i] = c;
        }
    }
    tr.mark("  s2 gather seeds");
    SCand sentinel; sentinel.readID = 0x7FFFFFFFu; sentinel.strand = 2; sentinel.pos = 0xFFFFFFFFull; sentinel.seedLen = 0xFFFFFFFFu;
    cand.push_back(sentinel);
    std::sort(cand.begin(), cand.end(), scand_less);
    std::vector<SCand> merged, canStream;
    single_merge(cand, merged);
    for (size_t i = 0, j; i < merged.size(); i = j) {                       // at most 200 per read (:186-199)
        j = i + 1;
        while (j < merged.size() && merged[j].readID == merged[i].readID) ++j;
        for (size_t k = i; k < j && k < i + 200; ++k) canStream.push_back(merged[k]);
    }
    // ---- S2 DP ----
    const uint32_t inputMax = (uint32_t)P->maxReadLength;
    const uint32_t maxReadLength = (inputMax / 4 + 1) * 4;
    const uint32_t maxDNALengthS = maxReadLength + 2 * MP_MARGIN(inputMax) + 8;
    MpDpParams dp; dp.mismatch = P->mismatchScore; dp.open = P->openGapScore; dp.clipLt = P->softClipLeft; dp.clipRt = P->softClipRight;
    std::vector<MpDpTask> tasks(canStream.size());
    for (size_t i = 0; i < canStream.size(); ++i) {
        const SCand &c = canStream[i];
        uint32_t readLength = lens[c.readID];
        uint32_t margin = MP_MARGIN(readLength);
        uint64_t start = c.pos - margin;
        if (start >= fullLen) start = 0;
        uint32_t dnaLen = readLength + margin * 2;
        if (start + dnaLen > fullLen) dnaLen = (uint32_t)(fullLen - start);
        MpDpTask t; memset(&t, 0, sizeof t);
        t.refStart = start; t.refLen = dnaLen; t.readID = c.readID; t.readLen = (uint16_t)readLength; t.strand = (uint8_t)c.strand;
        t.valid = 1; t.cutoff = dp_cutoff(readLength);
        t.diag = (int16_t)std::min<uint64_t>(c.pos - start, 0x7fff);          // the seed's diagonal inside the window (hint only)
        tasks[i] = t;
        cells += (uint64_t)dnaLen * readLength; ++tasksRun;
    }
    tr.mark("  s2 merge+tasks");
    std::vector<MpDpOut> outs; std::vector<uint8_t> pats;
    uint32_t patStride = (maxDNALengthS + maxReadLength + 3) & ~3u;
    if (int rc = mpd_run_host_tasks(ctx, tasks, maxDNALengthS, maxReadLength, dp, outs, pats, patStride)) return rc;
    tr.mark("  s2 dp");
    std::vector<mp_single_result> &S = ctx->hSingles;
    PinnedBuf<char> &HC = ctx->hCigars;
    for (size_t i = 0; i < tasks.size(); ++i) {
        if (outs[i].score < tasks[i].cutoff) continue;
        CigStats st;
        mp_single_result r; memset(&r, 0, sizeof r);
        r.cigar = append_cigar(HC, pats.data() + i * patStride, P->openGapScore, P->extendGapScore, st);
        if (r.cigar == 0xFFFFFFFFu) return MP_ERR_CUDA;
        r.readID = tasks[i].readID; r.strand = tasks[i].strand; r.seedAlignmentLength = canStream[i].seedLen;
        r.algnmt = tasks[i].refStart + outs[i].hitLoc; r.score = outs[i].score;
        r.startPos = tasks[i].refStart; r.refDpLength = tasks[i].refLen; r.peLeftAnchor = maxDNALengthS;
        int L = (int)tasks[i].readLen - st.nI - st.nS;
        int numMis = (L * P->matchScore + st.gapPenalty - outs[i].score) / (P->matchScore - P->mismatchScore);
        r.editdist = st.nI + st.nD + numMis;
        r.num_sameScore = (int32_t)outs[i].count;
        S.push_back(r);
    }
    // counters as DPSOutputThread keeps them: reads with >= 1 result; results after per-read de-duplication
    for (size_t i = 0, j; i < S.size(); i = j) {
        j = i + 1;
        while (j < S.size() && S[j].readID == S[i].readID) ++j;
        std::vector<std::pair<uint64_t, int32_t>> k;
        for (size_t a = i; a < j; ++a) k.push_back(std::make_pair(S[a].algnmt, S[a].score));
        std::sort(k.begin(), k.end());
        out->numSingleDPAligned += 1;
        out->numSingleDPAlignment += (uint64_t)(std::unique(k.begin(), k.end()) - k.begin());
    }
    tr.mark("  s2 assemble");
    if (P->skipDefaultDP || S.empty()) return 0;

    // ---- S3: sort by (readID, score desc) then (readID, score desc, startPos) ----
    std::vector<uint32_t> order(S.size());
    for (size_t i = 0; i < S.size(); ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const mp_single_result &x = S[a], &y = S[b];
        if (x.readID != y.readID) return x.readID < y.readID;
        if (x.score != y.score) return x.score > y.score;
        return x.startPos < y.startPos;
    });
    struct RTask { uint32_t refer; uint32_t leftOrRight; };
    std::vector<RTask> rinfo;
    std::vector<MpDpTask> rtasks;
    const int insert_high = P->insert_high, insert_low = P->insert_low;
    const uint32_t maxDNALengthR = (uint32_t)(insert_high - insert_low) + inputMax + 1;
    size_t it = 0;
    while (it < order.size()) {
        while (it + 4 < order.size() && S[order[it]].readID == S[order[it + 4]].readID) ++it;     // keep the last 4 of a read (:1053-1055)
        const uint32_t ref = order[it++];
        const mp_single_result &sr = S[ref];
        const uint32_t alignedReadID = sr.readID, unalignedReadID = alignedReadID ^ 1u;
        const uint64_t alignedPos = sr.algnmt;
        const uint32_t alignedLen = lens[alignedReadID], unalignedLen = lens[unalignedReadID];
        MpDpTask t; memset(&t, 0, sizeof t);
        t.readID = unalignedReadID; t.readLen = (uint16_t)unalignedLen; t.valid = 1; t.cutoff = dp_cutoff(unalignedLen);
        t.diag = -1;                                              // a rescue window has no seed
        if ((int)sr.strand == P->peStrandLeftLeg) {               // aligned read on the left, mate on the right
            uint64_t rightEnd = alignedPos + (uint64_t)(int64_t)insert_high;
            uint64_t rightStart = alignedPos + (uint64_t)(int64_t)insert_low - unalignedLen;
            if (rightStart < alignedPos) rightStart = alignedPos;
            if (rightStart < fullLen && rightEnd <= fullLen) {
                t.refStart = rightStart; t.refLen = (uint32_t)(rightEnd - rightStart); t.strand = (uint8_t)P->peStrandRightLeg;
                rtasks.push_back(t); RTask ri = { ref, 1u }; rinfo.push_back(ri);
            }
        }
        if ((int)sr.strand == P->peStrandRightLeg) {              // aligned read on the right, mate on the left
            uint64_t leftStart = alignedPos + alignedLen - (uint64_t)(int64_t)insert_high;
            uint64_t leftEnd = alignedPos + alignedLen - (uint64_t)(int64_t)insert_low + unalignedLen;
            if (leftEnd >= alignedPos + alignedLen) leftEnd = alignedPos + alignedLen - 1;
            if (leftStart < fullLen && leftEnd <= fullLen) {
                t.refStart = leftStart; t.refLen = (uint32_t)(leftEnd - leftStart); t.strand = (uint8_t)P->peStrandLeftLeg;
                rtasks.push_back(t); RTask ri = { ref, 0u }; rinfo.push_back(ri);
            }
        }
    }
    for (const MpDpTask &t : rtasks) {
        if (t.refLen > maxDNALengthR) { mp_set_error("default DP window %u exceeds maxDNALength %u", t.refLen, maxDNALengthR); return MP_ERR_CAPACITY; }
        cells += (uint64_t)t.refLen * t.readLen; ++tasksRun;
    }
    tr.mark("  s3 tasks");
    std::vector<MpDpOut> routs; std::vector<uint8_t> rpats;
    const uint32_t rStride = maxDNALengthR + maxReadLength;
    if (int rc = mpd_run_host_tasks(ctx, rtasks, maxDNALengthR, maxReadLength, dp, routs, rpats, rStride)) return rc;
    tr.mark("  s3 dp");
    // ---- AlgnmtDPResult records, grouped per pair (DV-DPfunctions.cpp:1476-1747) ----
    struct ADP { uint64_t a1, a2; int s1, s2; int which; mp_pair_result full; };
    std::vector<ADP> group;
    std::vector<mp_pair_result> &R = ctx->hRescued;
    auto flush = [&]() {
        if (group.empty()) return;
        // OutputBuffer::ready(1): sort + drop duplicates (ResultCompare), then drop half-aligned entries
        std::sort(group.begin(), group.end(), [](const ADP &a, const ADP &b) {
            return std::make_tuple(a.a1, a.a2, a.s1, a.s2) < std::make_tuple(b.a1, b.a2, b.s1, b.s2); });
        size_t w = 0;
        for (size_t i = 1; i < group.size(); ++i)
            if (std::make_tuple(group[w].a1, group[w].a2, group[w].s1, group[w].s2) < std::make_tuple(group[i].a1, group[i].a2, group[i].s1, group[i].s2))
                group[++w] = group[i];
        size_t n = w + 1, valid = 0;
        for (size_t i = 0; i < n; ++i) if (group[i].which < 2) { R.push_back(group[i].full); ++valid; }
        if (valid) { out->numRescuedPair += 1; out->numRescuedAlignment += valid; }
        group.clear();
    };
    uint32_t lastPair = 0xFFFFFFFFu;
    for (size_t id = 0; id < rtasks.size(); ++id) {
        const mp_single_result &sr = S[rinfo[id].refer];
        const uint32_t alignedID = sr.readID, alignedIsMate = alignedID & 1u;
        const uint32_t pairID = alignedID - alignedIsMate;
        if (pairID != lastPair) { flush(); lastPair = pairID; }
        const uint32_t canPos32 = (uint32_t)sr.algnmt;           // `uint canInfoAmbPosition` (DV-DPfunctions.cpp:1516)
        const int legStrand = rinfo[id].leftOrRight == 0 ? P->peStrandLeftLeg : P->peStrandRightLeg;
        ADP a; memset(&a, 0, sizeof a);
        mp_pair_result &f = a.full;
        f.readID = pairID;
        uint64_t dpPos = ~0ull; int dpScore = routs[id].score;
        // the DP side
        uint32_t dpCigar = 0; int dpEdit = 0; int32_t dpSame = 0;
        if (routs[id].score >= rtasks[id].cutoff) {
            CigStats st;
            dpCigar = append_cigar(HC, rpats.data() + id * rStride, P->openGapScore, P->extendGapScore, st);
            if (dpCigar == 0xFFFFFFFFu) return MP_ERR_CUDA;
            int L = (int)rtasks[id].readLen - st.nI - st.nS;
            int numMis = (L * P->matchScore + st.gapPenalty - routs[id].score) / (P->matchScore - P->mismatchScore);
            dpEdit = st.nI + st.nD + numMis;
            dpPos = rtasks[id].refStart + routs[id].hitLoc;
            a.which = 1 - (int)alignedIsMate;
            if (dpPos < (uint64_t)canPos32) f.insertSize = (int32_t)((uint64_t)canPos32 - dpPos + lens[alignedID]);
            else f.insertSize = (int32_t)(dpPos - (uint64_t)canPos32 + rtasks[id].readLen + st.nD - st.nI - st.nS);
            dpSame = (int32_t)routs[id].count;
        } else a.which = 2;
        const uint32_t lA = rinfo[id].leftOrRight == 1 ? maxDNALengthR : (uint32_t)(insert_high - insert_low + 1);
        const uint32_t rA = rinfo[id].leftOrRight == 1 ? (uint32_t)rtasks[id].readLen : 0u;
        if (alignedIsMate == 0) {          // aligned is read (mate 1), DP result is mate 2
            a.a1 = canPos32; a.a2 = dpPos; a.s1 = sr.score; a.s2 = dpScore;
            f.algnmt_1 = sr.algnmt; f.strand_1 = sr.strand; f.score_1 = sr.score; f.editdist_1 = sr.editdist; f.cigar_1 = sr.cigar;
            f.num_sameScore_1 = sr.num_sameScore; f.startPos_1 = (uint32_t)sr.startPos; f.refDpLength_1 = sr.refDpLength;
            f.peLeftAnchor_1 = sr.peLeftAnchor; f.peRightAnchor_1 = 0;
            f.algnmt_2 = dpPos; f.strand_2 = (uint8_t)legStrand; f.score_2 = dpScore; f.editdist_2 = dpEdit; f.cigar_2 = dpCigar;
            f.num_sameScore_2 = dpSame; f.startPos_2 = rtasks[id].refStart; f.refDpLength_2 = rtasks[id].refLen;
            f.peLeftAnchor_2 = lA; f.peRightAnchor_2 = rA;
        } else {                           // aligned is mate 2, DP result is mate 1
            a.a1 = dpPos; a.a2 = canPos32; a.s1 = dpScore; a.s2 = sr.score;
            f.algnmt_1 = dpPos; f.strand_1 = (uint8_t)legStrand; f.score_1 = dpScore; f.editdist_1 = dpEdit; f.cigar_1 = dpCigar;
            f.num_sameScore_1 = dpSame; f.startPos_1 = (uint32_t)rtasks[id].refStart; f.refDpLength_1 = rtasks[id].refLen;
            f.peLeftAnchor_1 = lA; f.peRightAnchor_1 = rA;
            f.algnmt_2 = sr.algnmt; f.strand_2 = sr.strand; f.score_2 = sr.score; f.editdist_2 = sr.editdist; f.cigar_2 = sr.cigar;
            f.num_sameScore_2 = sr.num_sameScore; f.startPos_2 = sr.startPos; f.refDpLength_2 = sr.refDpLength;
            f.peLeftAnchor_2 = sr.peLeftAnchor; f.peRightAnchor_2 = 0;
        }
        group.push_back(a);
    }
    flush();
    tr.mark("  s3 assemble");
    return 0;
}
