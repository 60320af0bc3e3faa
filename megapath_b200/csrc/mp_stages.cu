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
// These stages see only the pairs stage S1 could not place.  Stage S2 runs on the device from end to end: the seeds of every
// unplaced read are merged, ordered (with libstdc++'s own std::sort algorithm, mp_stdsort.h, because the reference's unstable sort
// decides which equally long seeds survive the cut) and capped by k_single_merge, turned into DP tasks, aligned by the same kernels
// as S1, and assembled into SingleAlgnmtResult records + CIGAR text by k_single_measure / k_single_write; the host receives the
// compact result list only.  Stage S3 (mate rescue) works on that list -- a few records per unplaced pair -- and keeps its list
// plumbing (the reference's sort orders, the "last four hits" rule, per-pair grouping) on the host.
#include "mp_context.h"
#include "mp_cigar.h"
#include "mp_stdsort.h"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <tuple>
#include <string.h>

namespace {

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

// ---- stage S2 on the device ----
// SingleDPWrapper::transferSeed + SingleEndSeedingEngine::singleMerge (DV-DPForSingleReads.cpp:121-215, DV-DPfunctions.cpp:295-342) for
// one read per thread.  The read's SeedPos entries (SeedPool.cpp:191-207) are already in the order the reference sorts them into
// (strand, position; k_merge writes them that way, and no two entries of a read share strand and position), so the radix sort of
// RadixTraitsCandidateInfo is the identity here.  Kept seeds go to the read's own segment of a scratch array, in their final order.
struct SSeed { uint64_t pos; uint32_t seedLen; uint32_t strand; };
struct SSeedLonger { __host__ __device__ bool operator()(const SSeed &a, const SSeed &b) const { return a.seedLen > b.seedLen; } };
__global__ void k_single_merge(const uint8_t *__restrict__ alignedPair, uint32_t nReads, const uint32_t *__restrict__ hitStart,
                               const uint32_t *__restrict__ nPos, const uint32_t *__restrict__ nNeg, const mp_seed_pos *__restrict__ sp,
                               SSeed *__restrict__ scratch, uint32_t *__restrict__ counts, unsigned int *__restrict__ nUnplacedPairs)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nReads || alignedPair[r >> 1]) return;
    if ((r & 1) == 0) atomicAdd(nUnplacedPairs, 1u);
    const uint32_t cnt = nPos[r] + nNeg[r];
    if (cnt == 0) return;
    const mp_seed_pos *src = sp + hitStart[r];
    SSeed *out = scratch + hitStart[r];
    uint32_t nOut = 0;
    for (uint32_t p = 0; p < cnt; ++p) {
        SSeed c; c.pos = src[p].pos; c.seedLen = src[p].paired_seedLength & 0x7FFFFFFFu; c.strand = (src[p].strand_readID >> 31) + 1;
        if (c.seedLen < 17) continue;
        while (p + 1 < cnt) {                                       // seeds closer than DPS_DIVIDE_GAP collapse to the longest
            SSeed d; d.pos = src[p + 1].pos; d.seedLen = src[p + 1].paired_seedLength & 0x7FFFFFFFu; d.strand = (src[p + 1].strand_readID >> 31) + 1;
            if (d.pos < c.pos + 5 && c.strand == d.strand) { if (d.seedLen > c.seedLen) c = d; }
            else break;
            ++p;
        }
        out[nOut++] = c;
    }
    mp_stdsort::sort(out, out + nOut, SSeedLonger());
    if (nOut) while ((double)out[nOut - 1].seedLen < out[0].seedLen * 0.6) --nOut;
    if (nOut > 200) nOut = 200;                                     // DV-DPForSingleReads.cpp:186-199
    counts[r] = nOut;
}
// SingleEndAlgnBatch::pack (DV-DPfunctions.cpp:401-445): one DP task per kept seed
__global__ void k_single_tasks(uint32_t nReads, const uint32_t *__restrict__ hitStart, const SSeed *__restrict__ scratch,
                               const uint32_t *__restrict__ counts, const uint32_t *__restrict__ taskStart, const uint32_t *__restrict__ lens,
                               uint64_t fullLen, MpDpTask *__restrict__ tasks, unsigned long long *__restrict__ work)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0, nt = 0;
    if (r < nReads) {
        const uint32_t n = counts[r];
        const SSeed *in = scratch + hitStart[r];
        const uint32_t readLength = lens[r], margin = MP_MARGIN(readLength);
        for (uint32_t k = 0; k < n; ++k) {
            uint64_t start = in[k].pos - margin;
            if (start >= fullLen) start = 0;
            uint32_t dnaLen = readLength + margin * 2;
            if (start + dnaLen > fullLen) dnaLen = (uint32_t)(fullLen - start);
            MpDpTask t; t.refStart = start; t.refLen = dnaLen; t.readID = r; t.readLen = (uint16_t)readLength; t.strand = (uint8_t)in[k].strand;
            t.valid = 1; t.cutoff = dp_cutoff(readLength);
            t.diag = (int16_t)min(in[k].pos - start, (uint64_t)0x7fff);      // the seed's diagonal inside the window (hint only)
            t.pad_ = (uint16_t)min(in[k].seedLen, 0xFFFFu);                  // seedAlignmentLength of the result
            tasks[taskStart[r] + k] = t;
            cells += (unsigned long long)dnaLen * readLength; ++nt;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { cells += __shfl_xor_sync(0xffffffffu, cells, d); nt += __shfl_xor_sync(0xffffffffu, nt, d); }
    if ((threadIdx.x & 31) == 0 && nt) { atomicAdd(&work[0], cells); atomicAdd(&work[1], nt); }
}
// SingleDP_Space::algnmtCPUThread (DV-DPfunctions.cpp:678-750): tasks that reached their cutoff become SingleAlgnmtResult records
__global__ void k_single_measure(uint32_t n, const MpDpTask *__restrict__ tasks, const MpDpOut *__restrict__ outs, const uint8_t *__restrict__ pats,
                                 uint32_t patStride, int open, int ext, uint32_t *__restrict__ okFlag, uint32_t *__restrict__ cigBytes)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const bool ok = outs[c].score >= tasks[c].cutoff;
    uint32_t bytes = 0;
    if (ok) bytes = (uint32_t)cigar_encode(pats + (size_t)c * patStride, open, ext, nullptr, 0).textLen + 1;
    okFlag[c] = ok; cigBytes[c] = bytes;
}
__global__ void k_single_write(uint32_t n, const MpDpTask *__restrict__ tasks, const MpDpOut *__restrict__ outs, const uint8_t *__restrict__ pats,
                               uint32_t patStride, int match, int mm, int open, int ext, uint32_t leftAnchor, const uint32_t *__restrict__ okFlag,
                               const uint32_t *__restrict__ outIdx, const uint32_t *__restrict__ cigOff, uint32_t *__restrict__ totals, uint32_t cigCap,
                               uint32_t cigArenaBase, mp_single_result *__restrict__ res, char *__restrict__ cig)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n || !okFlag[c]) return;
    const uint32_t off = totals[1] + cigOff[c];
    if ((uint64_t)totals[1] + cigOff[c + 1] > cigCap) { totals[4] = 1; return; }
    const int textLen = (int)(cigOff[c + 1] - cigOff[c]) - 1;
    const CigStats st = cigar_encode(pats + (size_t)c * patStride, open, ext, cig + off, textLen);
    cig[off + textLen] = 0;
    const MpDpTask t = tasks[c]; const MpDpOut o = outs[c];
    mp_single_result r; memset(&r, 0, sizeof r);
    r.cigar = cigArenaBase + off;
    r.readID = t.readID; r.strand = t.strand; r.seedAlignmentLength = t.pad_;
    r.algnmt = t.refStart + o.hitLoc; r.score = o.score;
    r.startPos = t.refStart; r.refDpLength = t.refLen; r.peLeftAnchor = leftAnchor;
    const int L = (int)t.readLen - st.nI - st.nS;
    const int numMis = (L * match + st.gapPenalty - o.score) / (match - mm);
    r.editdist = st.nI + st.nD + numMis;
    r.num_sameScore = (int32_t)o.count;
    res[totals[0] + outIdx[c]] = r;
}
__global__ void k_add_totals2(const uint32_t *__restrict__ idxTotal, const uint32_t *__restrict__ offTotal, uint32_t *__restrict__ totals)
{
    totals[0] += *idxTotal; totals[1] += *offTotal;
}
static int scan_u32_s(mp_context *ctx, const uint32_t *in, uint32_t *out, uint64_t n)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int64_t)n, ctx->stream);
    if (ctx->dScanTmp.reserve(tb)) return MP_ERR_CUDA;
    cub::DeviceScan::ExclusiveSum(ctx->dScanTmp.p, tb, in, out, (int64_t)n, ctx->stream);
    return 0;
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
    const uint32_t inputMax = (uint32_t)P->maxReadLength;
    const uint32_t maxReadLength = (inputMax / 4 + 1) * 4;
    const uint32_t maxDNALengthS = maxReadLength + 2 * MP_MARGIN(inputMax) + 8;
    MpDpParams dp; dp.mismatch = P->mismatchScore; dp.open = P->openGapScore; dp.clipLt = P->softClipLeft; dp.clipRt = P->softClipRight;
    std::vector<mp_single_result> &S = ctx->hSingles;
    PinnedBuf<char> &HC = ctx->hCigars;
    // ---- S2: merge / order / cap the seeds of every unplaced read, one DP task per kept seed (all on the device) ----
    if (ctx->dS2Counts.reserve(((size_t)nReads + 1) * 4) || ctx->dS2Start.reserve(((size_t)nReads + 1) * 4) || ctx->dTotals.reserve(16 * 4) ||
        ctx->dCounters.reserve(16 * 8)) return MP_ERR_CUDA;
    uint32_t *dTot = ctx->dTotals.as<uint32_t>();                              // [0] results, [1] cigar bytes, [4] overflow, [5] unplaced pairs
    unsigned long long *dWork = ctx->dCounters.as<unsigned long long>() + 11; // cells, tasks of this stage
    MP_CUDA(cudaMemsetAsync(ctx->dS2Counts.p, 0, ((size_t)nReads + 1) * 4, st));
    MP_CUDA(cudaMemsetAsync(dTot, 0, 8 * 4, st));
    MP_CUDA(cudaMemsetAsync(dWork, 0, 16, st));
    (++g_mp_launches), k_single_merge<<<(nReads + 127) / 128, 128, 0, st>>>(ctx->dAligned.as<uint8_t>(), nReads, ctx->dHitStart.as<uint32_t>(),
        ctx->dNPos.as<uint32_t>(), ctx->dNNeg.as<uint32_t>(), ctx->dSeedPos.as<mp_seed_pos>(), ctx->dHits.as<SSeed>(), ctx->dS2Counts.as<uint32_t>(), dTot + 5);
    if (scan_u32_s(ctx, ctx->dS2Counts.as<uint32_t>(), ctx->dS2Start.as<uint32_t>(), (uint64_t)nReads + 1)) return MP_ERR_CUDA;
    uint32_t nTasks = 0, nUnplaced = 0;
    MP_CUDA(cudaMemcpyAsync(&nTasks, ctx->dS2Start.as<uint32_t>() + nReads, 4, cudaMemcpyDeviceToHost, st));
    MP_CUDA(cudaMemcpyAsync(&nUnplaced, dTot + 5, 4, cudaMemcpyDeviceToHost, st));
    MP_CUDA(cudaStreamSynchronize(st));
    tr.mark("  s2 merge (device)");
    if (nUnplaced == 0) return 0;                          // every pair was placed by the deep DP
    const uint32_t patStride = (maxDNALengthS + maxReadLength + 3) & ~3u;
    const uint32_t CH = 1u << 18;
    const uint32_t chunkCap = std::min<uint32_t>(CH, nTasks + 1);
    if (ctx->dS2Tasks.reserve(((size_t)nTasks + 1) * sizeof(MpDpTask)) || ctx->dS2Res.reserve(((size_t)nTasks + 1) * sizeof(mp_single_result)) ||
        ctx->dLO.reserve((size_t)chunkCap * sizeof(MpDpOut)) || ctx->dLP.reserve((size_t)chunkCap * patStride) ||
        ctx->dOk.reserve(((size_t)chunkCap + 1) * 4) || ctx->dBytes.reserve(((size_t)chunkCap + 1) * 8) ||
        ctx->dIdx.reserve(((size_t)chunkCap + 1) * 4) || ctx->dOff.reserve(((size_t)chunkCap + 1) * 4)) return MP_ERR_CUDA;
    if (nTasks)
        (++g_mp_launches), k_single_tasks<<<(nReads + 127) / 128, 128, 0, st>>>(nReads, ctx->dHitStart.as<uint32_t>(), ctx->dHits.as<SSeed>(),
            ctx->dS2Counts.as<uint32_t>(), ctx->dS2Start.as<uint32_t>(), ctx->dLens.as<uint32_t>(), fullLen, ctx->dS2Tasks.as<MpDpTask>(), dWork);
    const uint32_t cigArenaBase = (uint32_t)HC.size();
    size_t cigCap = std::max<size_t>(ctx->dCig.cap, (size_t)1 << 20);
    uint32_t tot[8] = { 0 };
    for (int attempt = 0; attempt < 3 && nTasks; ++attempt) {
        if (ctx->dCig.reserve(cigCap)) return MP_ERR_CUDA;
        MP_CUDA(cudaMemsetAsync(dTot, 0, 5 * 4, st));
        for (uint32_t base = 0; base < nTasks; base += CH) {
            const uint32_t n = std::min<uint32_t>(CH, nTasks - base);
            const MpDpTask *tk = ctx->dS2Tasks.as<MpDpTask>() + base;
            const unsigned g = (n + 127) / 128;
            if (int rc = mpd_run_tasks(ctx, tk, n, maxDNALengthS, maxReadLength, dp, ctx->dLO.as<MpDpOut>(), ctx->dLP.as<uint8_t>(), patStride)) return rc;
            MP_CUDA(cudaMemsetAsync(ctx->dOk.p, 0, ((size_t)n + 1) * 4, st));
            MP_CUDA(cudaMemsetAsync(ctx->dBytes.p, 0, ((size_t)n + 1) * 4, st));
            (++g_mp_launches), k_single_measure<<<g, 128, 0, st>>>(n, tk, ctx->dLO.as<MpDpOut>(), ctx->dLP.as<uint8_t>(), patStride, P->openGapScore, P->extendGapScore,
                                                                   ctx->dOk.as<uint32_t>(), ctx->dBytes.as<uint32_t>());
            if (scan_u32_s(ctx, ctx->dOk.as<uint32_t>(), ctx->dIdx.as<uint32_t>(), (uint64_t)n + 1)) return MP_ERR_CUDA;
            if (scan_u32_s(ctx, ctx->dBytes.as<uint32_t>(), ctx->dOff.as<uint32_t>(), (uint64_t)n + 1)) return MP_ERR_CUDA;
            (++g_mp_launches), k_single_write<<<g, 128, 0, st>>>(n, tk, ctx->dLO.as<MpDpOut>(), ctx->dLP.as<uint8_t>(), patStride, P->matchScore, P->mismatchScore,
                P->openGapScore, P->extendGapScore, maxDNALengthS, ctx->dOk.as<uint32_t>(), ctx->dIdx.as<uint32_t>(), ctx->dOff.as<uint32_t>(), dTot,
                (uint32_t)std::min<size_t>(cigCap, 0xFFFFFFF0u), cigArenaBase, ctx->dS2Res.as<mp_single_result>(), ctx->dCig.as<char>());
            (++g_mp_launches), k_add_totals2<<<1, 1, 0, st>>>(ctx->dIdx.as<uint32_t>() + n, ctx->dOff.as<uint32_t>() + n, dTot);
            MP_CUDA(cudaGetLastError());
        }
        MP_CUDA(cudaMemcpyAsync(tot, dTot, sizeof tot, cudaMemcpyDeviceToHost, st));
        MP_CUDA(cudaStreamSynchronize(st));
        if (!tot[4]) break;
        cigCap = (size_t)tot[1] + tot[1] / 8 + (1 << 20);
        if (attempt == 2) { mp_set_error("CIGAR arena of stage S2 overflowed repeatedly"); return MP_ERR_CAPACITY; }
    }
    tr.mark("  s2 dp + assemble (device)");
    S.resize(tot[0]);
    if (HC.resize((size_t)cigArenaBase + tot[1])) return MP_ERR_CUDA;
    if (tot[0]) MP_CUDA(cudaMemcpyAsync(S.data(), ctx->dS2Res.p, (size_t)tot[0] * sizeof(mp_single_result), cudaMemcpyDeviceToHost, st));
    if (tot[1]) MP_CUDA(cudaMemcpyAsync(HC.data() + cigArenaBase, ctx->dCig.p, tot[1], cudaMemcpyDeviceToHost, st));
    {
        unsigned long long hw[2];
        MP_CUDA(cudaMemcpyAsync(hw, dWork, sizeof hw, cudaMemcpyDeviceToHost, st));
        MP_CUDA(cudaStreamSynchronize(st));
        cells += hw[0]; tasksRun += hw[1];
    }
    tr.mark("  s2 download");
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
