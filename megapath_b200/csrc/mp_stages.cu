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
// compact result list only.  Stage S3 (mate rescue) stays on the device as well: k_rescue_select orders a read's single-end hits the
// way the reference's two sorts do and keeps "the last four" (DV-DPfunctions.cpp:1053-1055), one rescue window per kept hit is
// aligned with the 752-wide instantiation of the same DP kernels, k_rescue_write assembles the AlgnmtDPResult records + CIGAR text
// and k_rescue_ready does the per-pair sort / de-duplication of OutputBuffer::ready; the host receives the final list.
#include "mp_context.h"
#include "mp_cigar.h"
#include "mp_stdsort.h"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <tuple>
#include <string.h>


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
    okFlag[c] = ok; cigBytes[c] = ok ? (uint32_t)outs[c].cigLen + 1 : 0u;       // the traceback measured its CIGAR while it ran
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
    const MpDpTask t = tasks[c]; const MpDpOut o = outs[c];
    const CigStats st = leg_stats(o);
    leg_text(o, pats + (size_t)c * patStride, patStride, open, ext, cig + off);
    cig[off + textLen] = 0;
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

// ---- stage S3 on the device ----
struct RescueInfo { uint32_t refer; uint32_t leftOrRight; };
// S3 order of a read's single-end hits: (score descending, startPos ascending), ties in list order (the reference sorts by
// (readID, score desc) and then by (readID, score desc, startPos), DV-SemiDP.cpp:170, 257)
__device__ __forceinline__ bool rescue_before(const mp_single_result &x, const mp_single_result &y)
{
    if (x.score != y.score) return x.score > y.score;
    return x.startPos < y.startPos;
}
// One thread per single-end result; the thread at the head of a read's run (results are ordered by read) handles the run:
//  * DPSOutputThread's counters: reads with a result, results after per-read de-duplication by (algnmt, score)
//  * HalfEndOccStream::fetchNextSingleAlgnResult (DV-DPfunctions.cpp:1048-1079): the LAST four hits of the read in S3 order
//  * HalfEndAlgnBatch::pack (:1151-1231): one rescue window per kept hit, on the side the hit's strand asks for
// Tasks go to slots [head, head + 4) of slot arrays as long as the result list (a run of n hits yields at most min(n, 4) tasks), flagged for
// the compaction that follows.
__global__ void k_rescue_select(const mp_single_result *__restrict__ S, uint32_t nS, const uint32_t *__restrict__ lens, uint64_t fullLen,
                                int insert_low, int insert_high, int strandLeft, int strandRight, uint32_t maxDNALengthR, int makeTasks,
                                MpDpTask *__restrict__ slotTasks, RescueInfo *__restrict__ slotInfo, uint32_t *__restrict__ slotFlag,
                                uint32_t *__restrict__ totals, unsigned long long *__restrict__ work)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0; uint32_t nt = 0, readsWith = 0, uniq = 0;
    if (i < nS && (i == 0 || S[i - 1].readID != S[i].readID)) {
        const uint32_t readID = S[i].readID;
        uint32_t e = i + 1;
        while (e < nS && S[e].readID == readID) ++e;
        readsWith = 1;
        for (uint32_t a = i; a < e; ++a) {
            bool dup = false;
            const uint64_t al = S[a].algnmt; const int sc = S[a].score;
            for (uint32_t b = i; b < a && !dup; ++b) dup = S[b].algnmt == al && S[b].score == sc;
            uniq += !dup;
        }
        if (makeTasks) {
            uint32_t top[4]; int m = 0;                       // the last (up to) four of the run in S3 order, ascending
            for (uint32_t a = i; a < e; ++a) {
                const mp_single_result x = S[a];
                int pos = m;                                  // a goes after everything that is not strictly behind it (stable)
                while (pos > 0 && rescue_before(x, S[top[pos - 1]])) --pos;
                if (m < 4) { for (int k = m; k > pos; --k) top[k] = top[k - 1]; top[pos] = a; ++m; }
                else if (pos > 0) { for (int k = 0; k + 1 < pos; ++k) top[k] = top[k + 1]; top[pos - 1] = a; }
            }
            uint32_t slot = i;
            for (int k = 0; k < m; ++k) {
                const mp_single_result sr = S[top[k]];
                const uint32_t alignedReadID = sr.readID, unalignedReadID = alignedReadID ^ 1u;
                const uint64_t alignedPos = sr.algnmt;
                const uint32_t alignedLen = lens[alignedReadID], unalignedLen = lens[unalignedReadID];
                MpDpTask t; memset(&t, 0, sizeof t);
                t.readID = unalignedReadID; t.readLen = (uint16_t)unalignedLen; t.valid = 1; t.cutoff = dp_cutoff(unalignedLen);
                t.diag = -1;                                  // a rescue window has no seed
                bool have = false; uint32_t side = 0;
                if ((int)sr.strand == strandLeft) {           // aligned read on the left, mate on the right
                    const uint64_t rightEnd = alignedPos + (uint64_t)(int64_t)insert_high;
                    uint64_t rightStart = alignedPos + (uint64_t)(int64_t)insert_low - unalignedLen;
                    if (rightStart < alignedPos) rightStart = alignedPos;
                    if (rightStart < fullLen && rightEnd <= fullLen) {
                        t.refStart = rightStart; t.refLen = (uint32_t)(rightEnd - rightStart); t.strand = (uint8_t)strandRight; have = true; side = 1;
                    }
                } else if ((int)sr.strand == strandRight) {   // aligned read on the right, mate on the left
                    const uint64_t leftStart = alignedPos + alignedLen - (uint64_t)(int64_t)insert_high;
                    uint64_t leftEnd = alignedPos + alignedLen - (uint64_t)(int64_t)insert_low + unalignedLen;
                    if (leftEnd >= alignedPos + alignedLen) leftEnd = alignedPos + alignedLen - 1;
                    if (leftStart < fullLen && leftEnd <= fullLen) {
                        t.refStart = leftStart; t.refLen = (uint32_t)(leftEnd - leftStart); t.strand = (uint8_t)strandLeft; have = true; side = 0;
                    }
                }
                if (!have) continue;
                if (t.refLen > maxDNALengthR) { totals[5] = 1; continue; }     // cannot happen for reads shorter than -L; reported, never run
                slotTasks[slot] = t;
                RescueInfo ri; ri.refer = top[k]; ri.leftOrRight = side; slotInfo[slot] = ri;
                slotFlag[slot] = 1; ++slot;
                cells += (unsigned long long)t.refLen * t.readLen; ++nt;
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        cells += __shfl_xor_sync(0xffffffffu, cells, d); nt += __shfl_xor_sync(0xffffffffu, nt, d);
        readsWith += __shfl_xor_sync(0xffffffffu, readsWith, d); uniq += __shfl_xor_sync(0xffffffffu, uniq, d);
    }
    if ((threadIdx.x & 31) == 0 && readsWith) {
        atomicAdd(&totals[6], readsWith); atomicAdd(&totals[7], uniq);
        if (nt) { atomicAdd(&work[0], cells); atomicAdd(&work[1], (unsigned long long)nt); }
    }
}
__global__ void k_rescue_compact(uint32_t nS, const uint32_t *__restrict__ slotFlag, const uint32_t *__restrict__ pos, const MpDpTask *__restrict__ slotTasks,
                                 const RescueInfo *__restrict__ slotInfo, MpDpTask *__restrict__ tasks, RescueInfo *__restrict__ info)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nS && slotFlag[i]) { tasks[pos[i]] = slotTasks[i]; info[pos[i]] = slotInfo[i]; }
}
// DP_Space::algnmtCPUThread (DV-DPfunctions.cpp:1476-1623): one AlgnmtDPResult per rescue task -- the single-end hit on one side, the DP
// result (or "no alignment") on the other.  pad carries (which | alignedIsMate << 4) for k_rescue_ready: which = 0 / 1 the mate the DP
// placed, 2 = the DP stayed below its cutoff.
__global__ void k_rescue_write(uint32_t n, const MpDpTask *__restrict__ tasks, const RescueInfo *__restrict__ info, const MpDpOut *__restrict__ outs,
                               const uint8_t *__restrict__ pats, uint32_t patStride, const mp_single_result *__restrict__ S, const uint32_t *__restrict__ lens,
                               int match, int mm, int open, int ext, int insert_low, int insert_high, int strandLeft, int strandRight,
                               uint32_t maxDNALengthR, const uint32_t *__restrict__ cigOff, uint32_t *__restrict__ totals, uint32_t cigCap,
                               uint32_t cigArenaBase, mp_pair_result *__restrict__ rec, char *__restrict__ cig)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const MpDpTask t = tasks[c]; const MpDpOut o = outs[c]; const RescueInfo ri = info[c];
    const mp_single_result sr = S[ri.refer];
    const uint32_t alignedID = sr.readID, alignedIsMate = alignedID & 1u, pairID = alignedID - alignedIsMate;
    const uint32_t canPos32 = (uint32_t)sr.algnmt;                // `uint canInfoAmbPosition` (DV-DPfunctions.cpp:1516)
    const int legStrand = ri.leftOrRight == 0 ? strandLeft : strandRight;
    mp_pair_result f; memset(&f, 0, sizeof f);
    f.readID = pairID;
    uint64_t dpPos = ~0ull; const int dpScore = o.score;
    uint32_t dpCigar = 0, which = 2; int dpEdit = 0; int32_t dpSame = 0;
    if (o.score >= t.cutoff) {
        const uint32_t off = totals[1] + cigOff[c];
        if ((uint64_t)totals[1] + cigOff[c + 1] > cigCap) { totals[4] = 1; return; }
        const int textLen = (int)(cigOff[c + 1] - cigOff[c]) - 1;
        const CigStats st = leg_stats(o);
        leg_text(o, pats + (size_t)c * patStride, patStride, open, ext, cig + off);
        cig[off + textLen] = 0;
        dpCigar = cigArenaBase + off;
        const int L = (int)t.readLen - st.nI - st.nS;
        const int numMis = (L * match + st.gapPenalty - o.score) / (match - mm);
        dpEdit = st.nI + st.nD + numMis;
        dpPos = t.refStart + o.hitLoc;
        which = 1u - alignedIsMate;
        if (dpPos < (uint64_t)canPos32) f.insertSize = (int32_t)((uint64_t)canPos32 - dpPos + lens[alignedID]);
        else f.insertSize = (int32_t)(dpPos - (uint64_t)canPos32 + t.readLen + st.nD - st.nI - st.nS);
        dpSame = (int32_t)o.count;
    }
    const uint32_t lA = ri.leftOrRight == 1 ? maxDNALengthR : (uint32_t)(insert_high - insert_low + 1);
    const uint32_t rA = ri.leftOrRight == 1 ? (uint32_t)t.readLen : 0u;
    if (alignedIsMate == 0) {          // aligned is read (mate 1), DP result is mate 2
        f.algnmt_1 = sr.algnmt; f.strand_1 = sr.strand; f.score_1 = sr.score; f.editdist_1 = sr.editdist; f.cigar_1 = sr.cigar;
        f.num_sameScore_1 = sr.num_sameScore; f.startPos_1 = (uint32_t)sr.startPos; f.refDpLength_1 = sr.refDpLength;
        f.peLeftAnchor_1 = sr.peLeftAnchor; f.peRightAnchor_1 = 0;
        f.algnmt_2 = dpPos; f.strand_2 = (uint8_t)legStrand; f.score_2 = dpScore; f.editdist_2 = dpEdit; f.cigar_2 = dpCigar;
        f.num_sameScore_2 = dpSame; f.startPos_2 = t.refStart; f.refDpLength_2 = t.refLen;
        f.peLeftAnchor_2 = lA; f.peRightAnchor_2 = rA;
    } else {                           // aligned is mate 2, DP result is mate 1
        f.algnmt_1 = dpPos; f.strand_1 = (uint8_t)legStrand; f.score_1 = dpScore; f.editdist_1 = dpEdit; f.cigar_1 = dpCigar;
        f.num_sameScore_1 = dpSame; f.startPos_1 = (uint32_t)t.refStart; f.refDpLength_1 = t.refLen;
        f.peLeftAnchor_1 = lA; f.peRightAnchor_1 = rA;
        f.algnmt_2 = sr.algnmt; f.strand_2 = sr.strand; f.score_2 = sr.score; f.editdist_2 = sr.editdist; f.cigar_2 = sr.cigar;
        f.num_sameScore_2 = sr.num_sameScore; f.startPos_2 = sr.startPos; f.refDpLength_2 = sr.refDpLength;
        f.peLeftAnchor_2 = sr.peLeftAnchor; f.peRightAnchor_2 = 0;
    }
    f.pad = (uint16_t)(which | (alignedIsMate << 4));
    rec[c] = f;
}
// sort key of an AlgnmtDPResult (ResultCompare, DV-DPfunctions.cpp:253-258): the single-end side enters with its 32-bit position
struct RescueKey { uint64_t a1, a2; int s1, s2; };
__device__ __forceinline__ RescueKey rescue_key(const mp_pair_result &f)
{
    RescueKey k; k.s1 = f.score_1; k.s2 = f.score_2;
    if ((f.pad >> 4) == 0) { k.a1 = (uint32_t)f.algnmt_1; k.a2 = f.algnmt_2; } else { k.a1 = f.algnmt_1; k.a2 = (uint32_t)f.algnmt_2; }
    return k;
}
__device__ __forceinline__ bool rescue_key_less(const RescueKey &a, const RescueKey &b)
{
    if (a.a1 != b.a1) return a.a1 < b.a1;
    if (a.a2 != b.a2) return a.a2 < b.a2;
    if (a.s1 != b.s1) return a.s1 < b.s1;
    return a.s2 < b.s2;
}
// DPOutputThread (DV-DPfunctions.cpp:1625-1747) per pair: OutputBuffer::ready(1) = sort + drop duplicates, then the half-aligned entries
// go.  A pair has at most eight records (four per mate), which std::sort orders by plain insertion: first of equal keys stays first.
__global__ void k_rescue_ready(mp_pair_result *__restrict__ rec, uint32_t n, uint32_t *__restrict__ keep, uint32_t *__restrict__ totals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t pairs = 0, kept = 0;
    if (i < n && (i == 0 || rec[i - 1].readID != rec[i].readID)) {
        uint32_t e = i + 1;
        while (e < n && rec[e].readID == rec[i].readID) ++e;
        for (uint32_t a = i + 1; a < e; ++a) {
            const mp_pair_result key = rec[a]; const RescueKey kk = rescue_key(key); uint32_t b = a;
            while (b > i && rescue_key_less(kk, rescue_key(rec[b - 1]))) { rec[b] = rec[b - 1]; --b; }
            if (b != a) rec[b] = key;
        }
        uint32_t last = i;
        for (uint32_t a = i; a < e; ++a) {
            const bool fresh = a == i || rescue_key_less(rescue_key(rec[last]), rescue_key(rec[a]));
            if (fresh) last = a;
            const bool k = fresh && (rec[a].pad & 0xF) < 2;
            keep[a] = k; kept += k;
        }
        pairs = kept ? 1 : 0;
    }
    pairs = __reduce_add_sync(0xffffffffu, pairs); kept = __reduce_add_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0 && kept) { atomicAdd(&totals[3], pairs); atomicAdd(&totals[2], kept); }
}
__global__ void k_rescue_out(const mp_pair_result *__restrict__ rec, uint32_t n, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pos,
                             mp_pair_result *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keep[i]) { mp_pair_result r = rec[i]; r.pad = 0; out[pos[i]] = r; }
}
__global__ void k_add_cig_total(const uint32_t *__restrict__ offTotal, uint32_t *__restrict__ totals) { totals[1] += *offTotal; }
static int scan_u32_s(mp_context *ctx, const uint32_t *in, uint32_t *out, uint64_t n)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int64_t)n, ctx->stream);
    if (ctx->dScanTmp.reserve(tb)) return MP_ERR_CUDA;
    cub::DeviceScan::ExclusiveSum(ctx->dScanTmp.p, tb, in, out, (int64_t)n, ctx->stream);
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
    MpTrace tr;
    const uint32_t inputMax = (uint32_t)P->maxReadLength;
    const uint32_t maxReadLength = (inputMax / 4 + 1) * 4;
    const uint32_t maxDNALengthS = maxReadLength + 2 * MP_MARGIN(inputMax) + 8;
    MpDpParams dp; dp.mismatch = P->mismatchScore; dp.open = P->openGapScore; dp.clipLt = P->softClipLeft; dp.clipRt = P->softClipRight;
    PinnedBuf<mp_single_result> &S = ctx->hSingles;
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
    if (S.resize(tot[0]) || HC.resize((size_t)cigArenaBase + tot[1])) return MP_ERR_CUDA;
    if (tot[0] && !ctx->resultsOnDevice) MP_CUDA(cudaMemcpyAsync(S.data(), ctx->dS2Res.p, (size_t)tot[0] * sizeof(mp_single_result), cudaMemcpyDeviceToHost, st));
    if (tot[1] && !ctx->resultsOnDevice) MP_CUDA(cudaMemcpyAsync(HC.data() + cigArenaBase, ctx->dCig.p, tot[1], cudaMemcpyDeviceToHost, st));
    {
        unsigned long long hw[2];
        MP_CUDA(cudaMemcpyAsync(hw, dWork, sizeof hw, cudaMemcpyDeviceToHost, st));
        MP_CUDA(cudaStreamSynchronize(st));
        cells += hw[0]; tasksRun += hw[1];
    }
    tr.mark("  s2 download");
    // ---- S3 (and the counters DPSOutputThread keeps: reads with >= 1 result, results after per-read de-duplication) ----
    const uint32_t nS = tot[0];
    if (nS == 0) return 0;
    const int doRescue = P->skipDefaultDP ? 0 : 1;
    const int insert_high = P->insert_high, insert_low = P->insert_low;
    const uint32_t maxDNALengthR = (uint32_t)(insert_high - insert_low) + inputMax + 1;
    const uint32_t rStride = (maxDNALengthR + maxReadLength + 3) & ~3u;
    uint32_t *dT3 = dTot + 8;                                                  // [1] cigar bytes, [2] kept, [3] pairs, [4] overflow, [5] window error, [6] reads, [7] unique
    if (ctx->dRsSlotTasks.reserve((size_t)nS * sizeof(MpDpTask)) || ctx->dRsSlotInfo.reserve((size_t)nS * sizeof(RescueInfo)) ||
        ctx->dRsFlag.reserve(((size_t)nS + 1) * 4) || ctx->dRsPos.reserve(((size_t)nS + 1) * 4)) return MP_ERR_CUDA;
    MP_CUDA(cudaMemsetAsync(dT3, 0, 8 * 4, st));
    MP_CUDA(cudaMemsetAsync(dWork, 0, 16, st));
    MP_CUDA(cudaMemsetAsync(ctx->dRsFlag.p, 0, ((size_t)nS + 1) * 4, st));
    (++g_mp_launches), k_rescue_select<<<(nS + 127) / 128, 128, 0, st>>>(ctx->dS2Res.as<mp_single_result>(), nS, ctx->dLens.as<uint32_t>(), fullLen, insert_low, insert_high,
        P->peStrandLeftLeg, P->peStrandRightLeg, maxDNALengthR, doRescue, ctx->dRsSlotTasks.as<MpDpTask>(), ctx->dRsSlotInfo.as<RescueInfo>(),
        ctx->dRsFlag.as<uint32_t>(), dT3, dWork);
    if (scan_u32_s(ctx, ctx->dRsFlag.as<uint32_t>(), ctx->dRsPos.as<uint32_t>(), (uint64_t)nS + 1)) return MP_ERR_CUDA;
    uint32_t nR = 0, t3[8] = { 0 };
    MP_CUDA(cudaMemcpyAsync(&nR, ctx->dRsPos.as<uint32_t>() + nS, 4, cudaMemcpyDeviceToHost, st));
    MP_CUDA(cudaMemcpyAsync(t3, dT3, sizeof t3, cudaMemcpyDeviceToHost, st));
    MP_CUDA(cudaStreamSynchronize(st));
    out->numSingleDPAligned += t3[6]; out->numSingleDPAlignment += t3[7];
    if (t3[5]) { mp_set_error("default DP window exceeds maxDNALength %u", maxDNALengthR); return MP_ERR_CAPACITY; }
    tr.mark("  s3 select (device)");
    if (!doRescue || nR == 0) return 0;
    const uint32_t CH3 = 1u << 17;
    const uint32_t chunkCap3 = std::min<uint32_t>(CH3, nR);
    if (ctx->dRsTasks.reserve((size_t)nR * sizeof(MpDpTask)) || ctx->dRsInfo.reserve((size_t)nR * sizeof(RescueInfo)) ||
        ctx->dRsRec.reserve((size_t)nR * sizeof(mp_pair_result)) || ctx->dRsOut.reserve((size_t)nR * sizeof(mp_pair_result)) ||
        ctx->dRsKeep.reserve(((size_t)nR + 1) * 4) || ctx->dRsKeepPos.reserve(((size_t)nR + 1) * 4) ||
        ctx->dLO.reserve((size_t)chunkCap3 * sizeof(MpDpOut)) || ctx->dLP.reserve((size_t)chunkCap3 * rStride) ||
        ctx->dOk.reserve(((size_t)chunkCap3 + 1) * 4) || ctx->dBytes.reserve(((size_t)chunkCap3 + 1) * 8) ||
        ctx->dOff.reserve(((size_t)chunkCap3 + 1) * 4)) return MP_ERR_CUDA;
    (++g_mp_launches), k_rescue_compact<<<(nS + 127) / 128, 128, 0, st>>>(nS, ctx->dRsFlag.as<uint32_t>(), ctx->dRsPos.as<uint32_t>(), ctx->dRsSlotTasks.as<MpDpTask>(),
        ctx->dRsSlotInfo.as<RescueInfo>(), ctx->dRsTasks.as<MpDpTask>(), ctx->dRsInfo.as<RescueInfo>());
    // the CIGAR text of S2 has left the device: the arena is reused from its start, offsets continue behind S2's in the host arena
    const uint32_t cigArenaBase3 = (uint32_t)HC.size();
    size_t cigCap3 = std::max<size_t>(ctx->dCig.cap, std::max<size_t>((size_t)nR * 64, (size_t)1 << 20));
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (cigCap3 > 0xFFFFFFF0ull) { mp_set_error("CIGAR arena of stage S3 exceeds 4 GB; use smaller batches"); return MP_ERR_CAPACITY; }
        if (ctx->dCig.reserve(cigCap3)) return MP_ERR_CUDA;
        MP_CUDA(cudaMemsetAsync(dT3, 0, 5 * 4, st));
        for (uint32_t base = 0; base < nR; base += CH3) {
            const uint32_t n = std::min<uint32_t>(CH3, nR - base);
            const MpDpTask *tk = ctx->dRsTasks.as<MpDpTask>() + base;
            const unsigned g = (n + 127) / 128;
            if (int rc = mpd_run_tasks(ctx, tk, n, maxDNALengthR, maxReadLength, dp, ctx->dLO.as<MpDpOut>(), ctx->dLP.as<uint8_t>(), rStride)) return rc;
            MP_CUDA(cudaMemsetAsync(ctx->dBytes.p, 0, ((size_t)n + 1) * 4, st));
            (++g_mp_launches), k_single_measure<<<g, 128, 0, st>>>(n, tk, ctx->dLO.as<MpDpOut>(), ctx->dLP.as<uint8_t>(), rStride, P->openGapScore, P->extendGapScore,
                                                                   ctx->dOk.as<uint32_t>(), ctx->dBytes.as<uint32_t>());
            if (scan_u32_s(ctx, ctx->dBytes.as<uint32_t>(), ctx->dOff.as<uint32_t>(), (uint64_t)n + 1)) return MP_ERR_CUDA;
            (++g_mp_launches), k_rescue_write<<<g, 128, 0, st>>>(n, tk, ctx->dRsInfo.as<RescueInfo>() + base, ctx->dLO.as<MpDpOut>(), ctx->dLP.as<uint8_t>(), rStride,
                ctx->dS2Res.as<mp_single_result>(), ctx->dLens.as<uint32_t>(), P->matchScore, P->mismatchScore, P->openGapScore, P->extendGapScore,
                insert_low, insert_high, P->peStrandLeftLeg, P->peStrandRightLeg, maxDNALengthR, ctx->dOff.as<uint32_t>(), dT3,
                (uint32_t)std::min<size_t>(cigCap3, 0xFFFFFFF0u), cigArenaBase3, ctx->dRsRec.as<mp_pair_result>() + base, ctx->dCig.as<char>());
            (++g_mp_launches), k_add_cig_total<<<1, 1, 0, st>>>(ctx->dOff.as<uint32_t>() + n, dT3);
            MP_CUDA(cudaGetLastError());
        }
        MP_CUDA(cudaMemsetAsync(ctx->dRsKeep.p, 0, ((size_t)nR + 1) * 4, st));
        (++g_mp_launches), k_rescue_ready<<<(nR + 127) / 128, 128, 0, st>>>(ctx->dRsRec.as<mp_pair_result>(), nR, ctx->dRsKeep.as<uint32_t>(), dT3);
        if (scan_u32_s(ctx, ctx->dRsKeep.as<uint32_t>(), ctx->dRsKeepPos.as<uint32_t>(), (uint64_t)nR + 1)) return MP_ERR_CUDA;
        (++g_mp_launches), k_rescue_out<<<(nR + 127) / 128, 128, 0, st>>>(ctx->dRsRec.as<mp_pair_result>(), nR, ctx->dRsKeep.as<uint32_t>(), ctx->dRsKeepPos.as<uint32_t>(),
                                                                         ctx->dRsOut.as<mp_pair_result>());
        MP_CUDA(cudaGetLastError());
        MP_CUDA(cudaMemcpyAsync(t3, dT3, sizeof t3, cudaMemcpyDeviceToHost, st));
        MP_CUDA(cudaStreamSynchronize(st));
        if (!t3[4]) break;
        cigCap3 = (size_t)t3[1] + t3[1] / 8 + (1 << 20);
        if (attempt == 2) { mp_set_error("CIGAR arena of stage S3 overflowed repeatedly"); return MP_ERR_CAPACITY; }
    }
    tr.mark("  s3 dp + assemble (device)");
    PinnedBuf<mp_pair_result> &R = ctx->hRescued;
    if (R.resize(t3[2]) || HC.resize((size_t)cigArenaBase3 + t3[1])) return MP_ERR_CUDA;
    if (t3[2] && !ctx->resultsOnDevice) MP_CUDA(cudaMemcpyAsync(R.data(), ctx->dRsOut.p, (size_t)t3[2] * sizeof(mp_pair_result), cudaMemcpyDeviceToHost, st));
    if (t3[1] && !ctx->resultsOnDevice) MP_CUDA(cudaMemcpyAsync(HC.data() + cigArenaBase3, ctx->dCig.p, t3[1], cudaMemcpyDeviceToHost, st));
    {
        unsigned long long hw[2];
        MP_CUDA(cudaMemcpyAsync(hw, dWork, sizeof hw, cudaMemcpyDeviceToHost, st));
        MP_CUDA(cudaStreamSynchronize(st));
        cells += hw[0]; tasksRun += hw[1];
    }
    out->numRescuedPair += t3[3]; out->numRescuedAlignment += t3[2];
    tr.mark("  s3 download");
    return 0;
}
