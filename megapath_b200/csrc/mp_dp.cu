// mp_dp.cu -- semi-global affine-gap DP with traceback on the device.
// Replaces SemiGlobalAligner::performAlignment -> callDP -> SemiGlobalAlignment ->
// GenerateDPTable + GPUBacktrack (CPU_DPfunctions.cpp:270-312; CPU_DP.cpp:881-978, 788-871,
// 122-619, 622-786) and the task packing of PairEndAlgnBatch::packRead/repackDNA
// (DV-DPfunctions.cpp:3009-3073).
//
// The reference encodes the recurrence as 8-bit saturating deltas with "score trimming";
// observable results equal the plain recurrence below (oracle/mp_oracle_dp.cpp,
// checked against the reference's own callDP).  Integer ALU / DPX work, no tensor cores.
//
// k_dp_fill<K>: one warp per PAIR of tasks.  The two tasks travel in the low and high 16-bit halves of
// every register, so each recurrence step is one packed DPX instruction for both (VIADDMNMX.U16x2,
// VIMNMX3.U16x2, VIMNMX.U16x2); scores carry a +16384 bias so that unsigned halves never borrow.
// Lane l owns read columns l*K+1 .. l*K+K and walks the reference rows with a one-row skew per lane
// (anti-diagonal wavefront); the right-most H / I of a lane's strip travel to lane l+1 by shuffle, the
// reference rows come from shared memory.  Every cell leaves one traceback byte
//    (H-Hdiag-mm)*42 + (H-Hleft-open)*3 + g          g 0 = D==H, 1 = neither, 2 = raised by the clip floor
// (the information the reference keeps, CPU_DP.cpp:183, 529-533) in a step-major table in HBM (layout at cell_offset):
// a warp's stores of one step are one full 128-byte line per task.
// The answer cell (first strict maximum in row-major order, tie count, CPU_DP.cpp:545-590) is kept per
// column in packed registers and reduced across the warp at the end.
// k_dp_tb<K>: one thread per task walks its own table (GPUBacktrack, CPU_DP.cpp:622-786) -- thousands of
// independent walks in flight hide the dependent-load latency that a single walking lane cannot.
// k_dp_exact<K>: runs first; tasks whose read occurs unchanged in its window get their (provably identical) answer
// without the DP, the others are compacted for the two kernels above.
#include "mp_context.h"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <type_traits>

#define DP_BIAS   0x4000
#define DP_BIAS2  0x40004000u
#define DP_ONE2   0x00010001u
#define DP_NEG2   0x20C020C0u      /* (bias - 8000) in both halves: "minus infinity" that never underflows */

__device__ __forceinline__ int h0_value(int j, int clipLt, int open)       // row 0 (CPU_DP.cpp:397-429)
{
    return j <= clipLt ? 0 : open - (j - clipLt - 1);
}
__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xffffu) * 0x00010001u; }

struct FillOut { int32_t score; uint32_t row, col, cnt; };

// Trace table of one task PAIR (S steps, S a multiple of 4; 2*S*32*K bytes).  Step t = row + lane.
//   word region  : step-major blocks of WORDS*256 bytes: [word w][task half h][lane] u32; byte c holds column 4w+c of the lane's
//                  strip for the row it had 3-c steps earlier, i.e. four cells of one DIAGONAL: the traceback's usual move stays
//                  inside one 32-byte sector
//   remainder    : the K%4 left-over columns of four consecutive steps share u32 words: group g = (t-1)/4 holds REM*256 bytes
//                  [word][half][lane] u32, byte position ((t-1)%4)*REM + (k-4*WORDS)
// so that every store of a warp is a full 128-byte line per task and one running pointer serves both tasks.
// cell_offset is relative to (pair table + half*128).
template <int K>
__device__ __forceinline__ size_t cell_offset(int r, int c, int S)
{
    const int lane = (c - 1) / K, k = (c - 1) - lane * K, t = r + lane;
    constexpr int WORDS = K / 4, REM = K % 4, WB = WORDS * 256;
    if (k < 4 * WORDS) return (size_t)(t + 3 - (k & 3)) * WB + (size_t)(k >> 2) * 256 + lane * 4 + (k & 3);
    const int b = ((t - 1) & 3) * REM + (k - 4 * WORDS);
    return (size_t)S * WB + (size_t)((t - 1) >> 2) * (REM * 256) + (size_t)(b >> 2) * 256 + lane * 4 + (b & 3);
}
// third digit of a trace byte -> the reference's flag (0 = raised by the clip floor, 1 = D==H, 2 = otherwise; CPU_DP.cpp:529-533)
__device__ __forceinline__ int trace_flag(uint32_t cell) { const int g = (int)(cell % 3); return g == 2 ? 0 : g + 1; }
// MM / OPEN: mismatch score and gap-open score as compile-time constants (0 = take them from P at run time), so that the
// packed constants become immediates of the DPX / IADD3 instructions
template <int K, int MM, int OPEN>
__global__ void __launch_bounds__(128)
k_dp_fill(const uint8_t *__restrict__ refSeq, const uint32_t *__restrict__ refLens, uint32_t refStride,
          const uint8_t *__restrict__ readSeq, const uint32_t *__restrict__ readLens, uint32_t readStride,
          const int32_t *__restrict__ cutoffs, uint32_t taskBase, uint32_t nTasks, MpDpParams P,
          uint8_t *__restrict__ tables, size_t tableStride, int S, FillOut *__restrict__ fill,
          const uint32_t *__restrict__ active, const uint32_t *__restrict__ nActive)
{
    extern __shared__ uint32_t refShared[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t pairId = blockIdx.x * 4 + wib;
    const uint32_t lA = pairId * 2, lB = lA + 1;             // slots inside this launch: tables are indexed by slot
    if (active) nTasks = *nActive;                           // only the tasks the exact-occurrence test left over (k_dp_exact)
    if (lA >= nTasks) return;
    const bool hasB = lB < nTasks;
    const uint32_t tA = taskBase + (active ? active[lA] : lA), tB = hasB ? taskBase + (active ? active[lB] : lB) : tA;
    const int mm = MM ? MM : P.mismatch, open = OPEN ? OPEN : P.open, clipLt = P.clipLt;
    int NA = (int)refLens[tA], LA = (int)readLens[tA], cutA = cutoffs[tA];
    int NB = hasB ? (int)refLens[tB] : 0, LB = hasB ? (int)readLens[tB] : 0, cutB = hasB ? cutoffs[tB] : 0;
    // CPU_DP.cpp:296-324: outside these bounds the reference aborts the SIMD group
    const bool okA = !(cutA > LA || cutA <= 0 || LA >= 255 + open - 1 + cutA || LA > 32 * K);
    const bool okB = hasB && !(cutB > LB || cutB <= 0 || LB >= 255 + open - 1 + cutB || LB > 32 * K);
    if (!okA) { NA = 0; LA = 0; }
    if (!okB) { NB = 0; LB = 0; }
    const int maxN = max(NA, NB), maxL = max(LA, LB);
    uint32_t *refS = refShared + (size_t)wib * S;
    {
        const uint8_t *fa = refSeq + (size_t)tA * refStride, *fb = refSeq + (size_t)tB * refStride;
        for (int r = lane; r < maxN; r += 32) {
            uint32_t a = r < NA ? fa[r] : 4u, b = r < NB ? fb[r] : 4u;
            refS[r] = a | (b << 16);
        }
    }
    __syncwarp();
    const uint8_t *ra = readSeq + (size_t)tA * readStride, *rbp = readSeq + (size_t)tB * readStride;
    const int j0 = lane * K + 1;
    uint32_t rb[K], Hp[K], Dp[K], fl[K], bestH[K], bestRow[K], cnt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = j0 + k;
        uint32_t a = j <= LA ? ra[j - 1] : 8u, b = j <= LB ? rbp[j - 1] : 8u;
        rb[k] = a | (b << 16);
        Hp[k] = pack2(h0_value(j, clipLt, open) + DP_BIAS);
        Dp[k] = DP_NEG2;
        fl[k] = j <= clipLt ? DP_BIAS2 : 0u;
        bestH[k] = 0; bestRow[k] = 0; cnt[k] = 0;
    }
    uint32_t prevHleft = pack2(h0_value(j0 - 1, clipLt, open) + DP_BIAS);
    const uint32_t OPENABS2 = pack2(-open), MMABS2 = pack2(-mm), MINUS1 = 0xFFFFFFFFu;
    const uint32_t DELTA = (uint32_t)(1 - mm);                 // match score - mismatch score
    constexpr int WORDS = K / 4, REM = K % 4, WB = WORDS * 256;
    uint8_t *tabP = tables + (size_t)lA * tableStride;          // pair table: 2 * tableStride bytes
    uint8_t *pw = tabP + WB + lane * 4;                         // word block of step 1
    uint8_t *pr = tabP + (size_t)S * WB + lane * 4;             // remainder group 0
    uint32_t sendH = 0, sendI = 0;
    const int lastLane = maxL > 0 ? (maxL - 1) / K : 0;
    const int steps = maxN + lastLane;
    const uint32_t amask = lastLane >= 31 ? 0xffffffffu : ((2u << lastLane) - 1u);     // lanes that own read columns
    uint32_t accA[REM ? REM : 1] = {0}, accB[REM ? REM : 1] = {0};
    uint32_t hist[WORDS][3][4];                                 // codes of the last steps, slot = step % 4 (see cell_offset)
#pragma unroll
    for (int w = 0; w < WORDS; ++w)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) hist[w][c][q] = 0;
    // One wavefront step of this lane; u = (t-1) % 4 is a compile-time constant in every caller.
    // CHECKED = false in the steady phase, where every owning lane has a row that exists in BOTH tasks: no range test, no row mask,
    // and the four unrolled steps need no register shuffling between them.
    auto step = [&](const int t, auto uTag, auto checked) {
        constexpr int u = decltype(uTag)::value;
        constexpr bool CHECKED = decltype(checked)::value;
        const uint32_t rH = __shfl_up_sync(amask, sendH, 1);
        const uint32_t rI = __shfl_up_sync(amask, sendI, 1);
        const int i = t - lane;
        uint32_t code[K] = {};
        const bool compute = !CHECKED || (i >= 1 && i <= maxN);
        if (compute) {
            uint32_t Hleft = lane == 0 ? DP_BIAS2 : rH;
            uint32_t Il = lane == 0 ? DP_NEG2 : rI;
            const uint32_t ref2 = refS[i - 1];
            const uint32_t rowMask = (i <= NA ? 0x0000FFFFu : 0u) | (i <= NB ? 0xFFFF0000u : 0u);
            const uint32_t row2 = (uint32_t)i * 0x00010001u;
            uint32_t Hdiag = prevHleft;
            prevHleft = Hleft;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint32_t m = __vminu2(ref2 ^ rb[k], DP_ONE2);               // 1 where the bases differ
                const uint32_t hs = Hdiag + DP_ONE2 - m * DELTA;                  // Hdiag + s(i,j)
                const uint32_t t1 = Hp[k] - OPENABS2;
                const uint32_t d = __viaddmax_u16x2(Dp[k], MINUS1, t1);           // max(D + ext, Hup + open)
                const uint32_t t2 = Hleft - OPENABS2;
                Il = __viaddmax_u16x2(Il, MINUS1, t2);                            // max(I + ext, Hleft + open)
                const uint32_t h = __vimax3_u16x2(hs, d, Il);
                const uint32_t hf = __vmaxu2(h, fl[k]);                           // clip floor for j <= clipLt (CPU_DP.cpp:505-510)
                const uint32_t a = hf - Hdiag + MMABS2;                           // H - Hdiag - mm   >= 0
                const uint32_t b = hf - t2;                                       // H - Hleft - open >= 0
                const uint32_t zd = __vminu2(hf - d, DP_ONE2);                    // 0 where D == H
                const uint32_t zr = __vminu2(hf - h, DP_ONE2);                    // 1 where the floor raised the cell (then zd == 1 too)
                code[k] = a * 42u + (b * 3u + (zd + zr));                         // third digit: 0 = D==H, 1 = neither, 2 = raised by the clip floor
                // answer cell per column: first strict maximum, ties counted (CPU_DP.cpp:545-590).  Column eligibility is applied
                // when the columns are reduced; rows past the end of the shorter task only occur in the CHECKED variant.
                const uint32_t hE = CHECKED ? (hf & rowMask) : hf;
                const uint32_t q = bestH[k] + 0x80008000u - hE;                   // per half: bit 15 set <=> best >= this cell (no borrow: values < 0x8000)
                uint32_t keep;                                                    // 0xFFFF in the halves that did NOT improve
                asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(keep) : "r"(q));        // replicate the sign bits of bytes 1 and 3
                const uint32_t lower = __viaddmin_s16x2_relu(q, 0x80008000u, DP_ONE2);   // 1 where this cell is below the best so far
                bestH[k] = __vmaxu2(bestH[k], hE);
                bestRow[k] = (bestRow[k] & keep) | (row2 & ~keep);
                cnt[k] = ((cnt[k] + DP_ONE2 - lower) & keep) | (DP_ONE2 & ~keep);
                Hdiag = Hp[k]; Hp[k] = hf; Dp[k] = d; Hleft = hf;
            }
            sendH = Hleft; sendI = Il;
#pragma unroll
            for (int kk = 0; kk < REM; ++kk) {
                const int bpos = u * REM + kk, pos = bpos & 3;                     // byte position inside the group's remainder words
                const uint32_t selA = 0x3210u ^ ((uint32_t)(4 ^ pos) << (4 * pos)), selB = 0x3210u ^ ((uint32_t)(6 ^ pos) << (4 * pos));
#pragma unroll
                for (int q = 0; q < REM; ++q)
                    if (q == (bpos >> 2)) {
                        accA[q] = __byte_perm(accA[q], code[4 * WORDS + kk], selA);
                        accB[q] = __byte_perm(accB[q], code[4 * WORDS + kk], selB);
                    }
            }
#pragma unroll
            for (int w = 0; w < WORDS; ++w)
#pragma unroll
                for (int c = 0; c < 3; ++c) hist[w][c][u] = code[4 * w + c];
        }
        // trace words: byte c of word w holds column 4w+c of row i-(3-c), so that a diagonal run of four cells shares one word.
        // Low halves -> task A, high halves -> task B.  Rows maxN+1..maxN+3 only flush the older bytes.
        if (!CHECKED || (i >= 1 && i <= maxN + 3)) {
#pragma unroll
            for (int w = 0; w < WORDS; ++w) {
                const uint32_t c0 = hist[w][0][(u + 1) & 3], c1 = hist[w][1][(u + 2) & 3], c2 = hist[w][2][(u + 3) & 3], c3 = code[4 * w + 3];
                const uint32_t p01 = __byte_perm(c0, c1, 0x6240), p23 = __byte_perm(c2, c3, 0x6240);
                // p01 bytes: [c0.lo, c1.lo, c0.hi, c1.hi]
                const uint32_t wa = __byte_perm(p01, p23, 0x5410), wb = __byte_perm(p01, p23, 0x7632);
                *(uint32_t *)(pw + u * WB + w * 256) = wa;
                if (hasB) *(uint32_t *)(pw + u * WB + w * 256 + 128) = wb;
            }
        }
    };
    // remainder words of the four-step group that ends at step t
    auto flush = [&](const int t) {
        if (REM && t - lane >= 1 && t - 3 - lane <= maxN) {
#pragma unroll
            for (int q = 0; q < REM; ++q) {
                *(uint32_t *)(pr + q * 256) = accA[q];
                if (hasB) *(uint32_t *)(pr + q * 256 + 128) = accB[q];
            }
        }
        pr += REM * 256;
    };
    auto group = [&](const int t, auto checked) {
        step(t, std::integral_constant<int, 0>(), checked);
        step(t + 1, std::integral_constant<int, 1>(), checked);
        step(t + 2, std::integral_constant<int, 2>(), checked);
        step(t + 3, std::integral_constant<int, 3>(), checked);
        flush(t + 3);
        pw += 4 * WB;
    };
    if (lane <= lastLane && maxL > 0) {
        const int steadyN = (NA > 0 && NB > 0) ? min(NA, NB) : maxN;    // rows present in every live task of the pair
        const int stepsR = (steps + 3 + 3) & ~3;                        // + 3 flush steps, whole groups
        const int rampEnd = min((lastLane + 3) & ~3, stepsR);           // from here on every owning lane has started
        const int steadyEnd = max(steadyN & ~3, rampEnd);               // up to here no owning lane has run out of rows
        int t = 1;
#pragma unroll 1
        for (; t <= rampEnd; t += 4) group(t, std::true_type());
#pragma unroll 1
        for (; t <= steadyEnd; t += 4) group(t, std::false_type());
#pragma unroll 1
        for (; t <= stepsR; t += 4) group(t, std::true_type());
    }
    __syncwarp();
    // ---- winner per task: max score over the eligible columns, then smallest (row, col); ties summed ----
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int sh = half * 16;
        const int Lh = half ? LB : LA, minCol = max(Lh - P.clipRt, 1);
        int best = 0;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (j0 + k >= minCol && j0 + k <= Lh) best = max(best, (int)((bestH[k] >> sh) & 0xffffu));
        int gbest = best;
#pragma unroll
        for (int dlt = 16; dlt; dlt >>= 1) gbest = max(gbest, __shfl_xor_sync(0xffffffffu, gbest, dlt));
        uint32_t key = 0xffffffffu; uint32_t c2 = 0;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (j0 + k >= minCol && j0 + k <= Lh && (int)((bestH[k] >> sh) & 0xffffu) == gbest) {
                key = min(key, (((bestRow[k] >> sh) & 0xffffu) << 12) | (uint32_t)(j0 + k));
                c2 += (cnt[k] >> sh) & 0xffffu;
            }
#pragma unroll
        for (int dlt = 16; dlt; dlt >>= 1) { key = min(key, __shfl_xor_sync(0xffffffffu, key, dlt)); c2 += __shfl_xor_sync(0xffffffffu, c2, dlt); }
        if (lane == 0 && (half == 0 || hasB)) {
            FillOut f; f.score = gbest - DP_BIAS; f.row = key >> 12; f.col = key & 0xfffu; f.cnt = c2;
            if (gbest == 0 || !(half ? okB : okA)) { f.score = 0; f.row = 0; f.col = 0; f.cnt = 0; }
            fill[half == 0 ? tA : tB] = f;
        }
    }
}

// ---- GPUBacktrack (CPU_DP.cpp:622-786), literal; one thread per task ----
// The walk is bound by memory TRANSACTIONS (every lane of a warp touches its own sector), so all three streams go through
// one-word register caches: the trace table (a diagonal run of four cells shares a word, see cell_offset), the two sequences
// (aligned 4-base words) and the pattern (four bytes per store; WORDPAT needs patStride % 4 == 0).
template <int K, bool WORDPAT>
__global__ void __launch_bounds__(128)
k_dp_tb(const uint8_t *__restrict__ refSeq, const uint32_t *__restrict__ refLens, uint32_t refStride,
        const uint8_t *__restrict__ readSeq, const uint32_t *__restrict__ readLens, uint32_t readStride,
        const int32_t *__restrict__ cutoffs, uint32_t taskBase, uint32_t nTasks, MpDpParams P,
        const uint8_t *__restrict__ tables, size_t tableStride, int S, const FillOut *__restrict__ fill,
        MpDpOut *__restrict__ outs, uint8_t *__restrict__ patterns, uint32_t patStride,
        const uint32_t *__restrict__ active, const uint32_t *__restrict__ nActive)
{
    const uint32_t lt = blockIdx.x * blockDim.x + threadIdx.x;
    if (active) nTasks = *nActive;
    if (lt >= nTasks) return;
    const uint32_t task = taskBase + (active ? active[lt] : lt);
    const int L = (int)readLens[task], cutoff = cutoffs[task];
    const int mm = P.mismatch, open = P.open, ext = -1, clipLt = P.clipLt;
    MpDpOut o; o.score = 0; o.hitLoc = 0; o.count = 0; o.patLen = 0;
    const FillOut f = fill[task];
    if (cutoff > L || cutoff <= 0 || L >= 255 + open - 1 + cutoff || L > 32 * K || f.score < cutoff) { outs[task] = o; return; }
    const uint8_t *tab = tables + (size_t)(lt & ~1u) * tableStride + (lt & 1u) * 128;   // pair table + this task's half
    uint8_t *pat = patterns + (size_t)task * patStride;
    // ---- cached accessors ----
    uint32_t tabW = 0; size_t tabIdx = ~(size_t)0;
    auto cellAt = [&](int r, int c) -> uint32_t {           // trace byte of (row r >= 1, column c >= 1)
        const size_t off = cell_offset<K>(r, c, S);
        if ((off >> 2) != tabIdx) { tabIdx = off >> 2; tabW = __ldcs((const uint32_t *)tab + tabIdx); }   // read once or twice: keep it out of L1
        return (tabW >> ((off & 3) * 8)) & 0xffu;
    };
    auto flagAt = [&](int r, int c) -> int {                // flag of any cell including the virtual row 0 / column 0
        if (c == 0) return 0;                               // column 0 cells are stored as 0 (CPU_DP.cpp:447-450)
        if (r == 0) return c <= clipLt ? 0 : 1;             // row 0 (CPU_DP.cpp:405-427)
        return trace_flag(cellAt(r, c));
    };
    auto hdAt = [&](int r, int c) -> int {                  // H[r][c] - H[r][c-1] for any row including row 0
        if (r == 0) return h0_value(c, clipLt, open) - h0_value(c - 1, clipLt, open);
        return open + (int)(cellAt(r, c) / 3 % 14);
    };
    const size_t rsOff = (size_t)task * readStride, fsOff = (size_t)task * refStride;
    uint32_t rsW = 0, fsW = 0; size_t rsIdx = ~(size_t)0, fsIdx = ~(size_t)0;
    auto readBase = [&](int x) -> uint32_t {
        const size_t a = rsOff + x;
        if ((a >> 2) != rsIdx) { rsIdx = a >> 2; rsW = __ldg((const uint32_t *)readSeq + rsIdx); }
        return (rsW >> ((a & 3) * 8)) & 0xffu;
    };
    auto refBase = [&](int x) -> uint32_t {
        const size_t a = fsOff + x;
        if ((a >> 2) != fsIdx) { fsIdx = a >> 2; fsW = __ldg((const uint32_t *)refSeq + fsIdx); }
        return (fsW >> ((a & 3) * 8)) & 0xffu;
    };
    uint32_t p = 0, pacc = 0;
    auto emit = [&](uint32_t byte) {
        if (WORDPAT) {
            pacc |= byte << ((p & 3) * 8);
            if ((p & 3) == 3) { ((uint32_t *)pat)[p >> 2] = pacc; pacc = 0; }
        } else pat[p] = (uint8_t)byte;
        ++p;
    };
    const int hitRow = (int)f.row, hitCol = (int)f.col;
    o.score = f.score; o.count = min(f.cnt, 255u);
    int clipR = L - hitCol;
    if (clipR > 0) { emit('S'); emit('V'); emit((uint32_t)clipR & 0xffu); }
    int i = L - clipR, j = hitRow;
    enum { NORMAL, I_EXT, D_EXT, SM_EXIT, SI_EXIT, SD_EXIT };
    int state = NORMAL;
    int accum = 0;
    // `diagCell` carries the byte of (j-1, i-1) from the clip check of one step to the next step, which usually moves there
    uint32_t cell = cellAt(j, i);
    while (i > 0 && j > 0) {
        int flag = trace_flag(cell);
        int hd = open + (int)(cell / 3 % 14);
        int dd = mm + (int)(cell / 42);
        if (state == NORMAL) {
            bool eq = refBase(j - 1) == readBase(i - 1);
            int ms = eq ? 1 : mm;
            if (dd == ms) {
                // flag of the diagonal predecessor, including the virtual row 0 / column 0 (CPU_DP.cpp:405-450)
                uint32_t diagCell = 0; int dflag;
                if (i - 1 == 0) dflag = 0;
                else if (j - 1 == 0) dflag = (i - 1) <= clipLt ? 0 : 1;
                else { diagCell = cellAt(j - 1, i - 1); dflag = trace_flag(diagCell); }
                if (i != 1 && dflag == 0) { state = SM_EXIT; break; }
                emit(eq ? 'M' : 'm'); --j; --i;
                cell = diagCell;                                     // valid whenever the loop continues (i > 0 && j > 0)
                continue;
            } else if (flag == 1) {
                int vd = dd - hdAt(j - 1, i);
                emit('D'); --j;
                if (vd != open) { accum = (int8_t)(vd - ext); state = D_EXT; }
            } else {
                emit('I'); --i;
                if (hd != open) { accum = (int8_t)(hd - ext); state = I_EXT; }
            }
        } else if (state == D_EXT) {
            int vd = dd - hdAt(j - 1, i);
            if (vd + accum == open && flagAt(j - 1, i) == 0) { state = SD_EXIT; break; }
            emit('D'); --j;
            if (vd + accum == open) state = NORMAL; else accum = (int8_t)(accum + vd - ext);
        } else {
            if (hd + accum == open && flagAt(j, i - 1) == 0) { state = SI_EXIT; break; }
            emit('I'); --i;
            if (hd + accum == open) state = NORMAL; else accum = (int8_t)(accum + hd - ext);
        }
        if (i > 0 && j > 0) cell = cellAt(j, i);
    }
    bool discard = false;
    if (j == 0) {
        int sc = min(clipLt & 0xff, i);
        if (sc < i) { emit('I'); emit('V'); emit((uint32_t)(i - sc) & 0xffu); }
        emit('S'); emit('V'); emit((uint32_t)sc & 0xffu);
    } else if (state == SI_EXIT) {
        emit('I'); emit('S'); emit('V'); emit((uint32_t)(i - 1) & 0xffu);
    } else if (state == SD_EXIT) {
        emit('D'); emit('S'); emit('V'); emit((uint32_t)(i - 1) & 0xffu);
        discard = true;                                   // CPU_DP.cpp:842-857
    } else if (state == SM_EXIT) {
        emit((refBase(j - 1) == readBase(i - 1)) ? 'M' : 'm');
        emit('S'); emit('V'); emit((uint32_t)(i - 1) & 0xffu);
        j -= 1;
    }
    o.patLen = p;
    emit(0);                                              // terminator
    if (WORDPAT && (p & 3)) ((uint32_t *)pat)[p >> 2] = pacc;
    if (discard) { o.score = 0; o.hitLoc = 0; } else o.hitLoc = (uint32_t)j;
    outs[task] = o;
}

// ------------------------------------------------------------------------------------
// Exact-occurrence test.  If the read occurs in its reference window without a single difference, the DP answer is known:
//  * no cell can exceed its column index (a column adds at most the match score 1; clips, gaps and mismatches add <= 0), so the
//    maximum over the eligible columns is L, reached only in column L, exactly at the rows where an occurrence ends;
//  * the answer cell is the first such row (row-major order), the tie count the number of occurrences (CPU_DP.cpp:545-590);
//  * on the diagonal of an occurrence H = column index, so every traceback step sees H - Hdiag = 1 with equal bases, no cell of the
//    diagonal was raised by the clip floor (H >= 1 there), and GPUBacktrack emits L times 'M'; it stops at column 0, and when that
//    is also row 0 it appends the zero-length left clip "SV\0" (CPU_DP.cpp:788-871).
// Such tasks skip k_dp_fill / k_dp_tb; the others are compacted into `active`.  One warp per task, one lane per window offset.
// ------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128)
k_dp_exact(const uint8_t *__restrict__ refSeq, const uint32_t *__restrict__ refLens, uint32_t refStride,
           const uint8_t *__restrict__ readSeq, const uint32_t *__restrict__ readLens, uint32_t readStride,
           const int32_t *__restrict__ cutoffs, uint32_t taskBase, uint32_t nTasks, MpDpParams P,
           MpDpOut *__restrict__ outs, uint8_t *__restrict__ patterns, uint32_t patStride, uint32_t *__restrict__ needDp)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lt = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (lt >= nTasks) return;
    const uint32_t task = taskBase + lt;
    const int N = (int)refLens[task], L = (int)readLens[task], cutoff = cutoffs[task];
    // same admission test as the DP kernels; scores must be the usual kind (match 1, everything else <= 0)
    const bool ok = !(cutoff > L || cutoff <= 0 || L >= 255 + P.open - 1 + cutoff || L > 32 * K) && L > 0 && N >= L &&
                    P.mismatch <= 0 && P.open <= 0;
    uint32_t occ = 0; int first = -1;
    if (ok) {
        const uint8_t *fs = refSeq + (size_t)task * refStride, *rs = readSeq + (size_t)task * readStride;
        const int shifts = N - L + 1;
        // every lane screens one window offset with the first bases of the read; the few offsets that survive are then verified by
        // the whole warp, one base per lane and trip
        const int PRE = L < 8 ? L : 8;
        for (int base = 0; base < shifts; base += 32) {
            const int o = base + lane;
            bool alive = o < shifts;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (alive && k < PRE) alive = fs[o + k] == rs[k];
            uint32_t cand = __ballot_sync(0xffffffffu, alive);
            while (cand) {
                const int oc = base + __ffs(cand) - 1;
                cand &= cand - 1;
                bool same = true;
                for (int k = PRE + lane; k < L; k += 32) same = same && fs[oc + k] == rs[k];
                if (__all_sync(0xffffffffu, same)) { if (first < 0) first = oc; ++occ; }
            }
        }
    }
    if (occ == 0) { if (lane == 0) needDp[lt] = 1; return; }
    uint8_t *pat = patterns + (size_t)task * patStride;
    for (int k = lane; k < L; k += 32) pat[k] = 'M';
    if (lane == 0) {
        uint32_t p = (uint32_t)L;
        if (first == 0) { pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = 0; }      // j == 0: min(clipLt, i = 0) = 0 clipped bases
        pat[p] = 0;
        MpDpOut o; o.score = L; o.hitLoc = (uint32_t)first; o.count = min(occ, 255u); o.patLen = p;
        outs[task] = o;
        needDp[lt] = 0;
    }
}
// slots of the tasks that still need the DP, in order; also adds up what the fill kernel is about to do (work accounting)
__global__ void k_dp_compact(const uint32_t *__restrict__ needDp, const uint32_t *__restrict__ pos, uint32_t nTasks, uint32_t taskBase,
                             const uint32_t *__restrict__ refLens, const uint32_t *__restrict__ readLens,
                             uint32_t *__restrict__ active, unsigned long long *__restrict__ counters)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0; uint32_t exact = 0;
    if (t < nTasks) {
        if (needDp[t]) { active[pos[t]] = t; cells = (unsigned long long)refLens[taskBase + t] * readLens[taskBase + t]; }
        else exact = 1;
    }
    cells = __reduce_add_sync(0xffffffffu, (uint32_t)cells);        // refLen * readLen < 2^24 per task: 32 of them fit 32 bits
    exact = __reduce_add_sync(0xffffffffu, exact);
    if ((threadIdx.x & 31) == 0 && counters) { if (cells) atomicAdd(&counters[15], cells); if (exact) atomicAdd(&counters[14], (unsigned long long)exact); }
}

// ------------------------------------------------------------------------------------
// task sequence extraction (replaces packRead / repackDNA): one byte per base
// ------------------------------------------------------------------------------------
__global__ void k_extract(MpIndexView ix, const uint32_t *__restrict__ reads, uint32_t wpq, const MpDpTask *__restrict__ tasks,
                          uint32_t nTasks, uint8_t *__restrict__ refSeq, uint32_t refStride, uint8_t *__restrict__ readSeq,
                          uint32_t readStride, uint32_t *__restrict__ refLens, uint32_t *__restrict__ readLens, int32_t *__restrict__ cutoffs)
{
    const uint32_t task = blockIdx.x;
    if (task >= nTasks) return;
    MpDpTask tk = tasks[task];
    if (!tk.valid) { tk.refLen = 0; }
    for (uint32_t a = threadIdx.x; a < tk.refLen; a += blockDim.x)
        refSeq[(size_t)task * refStride + a] = (uint8_t)mp_text_base(ix, tk.refStart + a);
    const uint32_t *rd = reads + (size_t)tk.readID * wpq;
    for (uint32_t a = threadIdx.x; a < tk.readLen; a += blockDim.x) {
        uint32_t p = tk.strand == 1 ? a : tk.readLen - 1 - a;
        uint32_t b = (rd[p >> 4] >> ((p & 15) << 1)) & 3;
        readSeq[(size_t)task * readStride + a] = (uint8_t)(tk.strand == 1 ? b : 3 - b);
    }
    if (threadIdx.x == 0) { refLens[task] = tk.refLen; readLens[task] = tk.valid ? tk.readLen : 0; cutoffs[task] = tk.cutoff; }
}

static int launch_dp(mp_context *ctx, const uint8_t *dRef, const uint32_t *dRefLens, uint32_t refStride,
                     const uint8_t *dRead, const uint32_t *dReadLens, uint32_t readStride, const int32_t *dCutoffs,
                     uint32_t nTasks, uint32_t maxRefLen, uint32_t maxReadLen, const MpDpParams &P,
                     MpDpOut *dOuts, uint8_t *dPatterns, uint32_t patStride)
{
    if (nTasks == 0) return 0;
    const int K = maxReadLen <= 160 ? 5 : maxReadLen <= 256 ? 8 : 10;
    if (maxReadLen > 320) { mp_set_error("read length %u exceeds the DP kernel bound 320", maxReadLen); return MP_ERR_ARG; }
    const int S = ((int)maxRefLen + 44) & ~3;                 // steps 1 .. maxRefLen + 31 + 3 flush steps, rounded up to groups of four, plus slack
    const size_t tableStride = (size_t)S * 32 * K;
    if (ctx->dFill.reserve((size_t)nTasks * sizeof(FillOut))) return MP_ERR_CUDA;
    // the traceback tables of one sub-batch stay in HBM between the two kernels.  A chunk of stage S1 has up to 2^18 tasks:
    // size for that even when this launch is smaller, and only talk to the allocator when the buffer really is too small
    const size_t typical = nTasks >= (1u << 15) ? (size_t)(1u << 18) * tableStride : (size_t)(nTasks + 1) * tableStride;
    const size_t ideal = std::max<size_t>((size_t)(nTasks + 1) * tableStride, typical);
    if (ctx->dTable.cap < ideal) {
        size_t freeB = 0, totalB = 0; cudaMemGetInfo(&freeB, &totalB);
        const size_t maxBytes = std::min<size_t>((size_t)24 << 30, (freeB + ctx->dTable.cap) / 2);
        const size_t want = std::min<size_t>(ideal, maxBytes);
        if (ctx->dTable.cap < want && ctx->dTable.reserve(want)) return MP_ERR_CUDA;
    }
    // tables are laid out per task PAIR: an odd final task still needs a whole pair table
    const uint32_t per = (uint32_t)std::min<size_t>(nTasks, (ctx->dTable.cap / tableStride) & ~(size_t)1);
    if (per == 0) { mp_set_error("not enough device memory for the DP traceback tables"); return MP_ERR_CUDA; }
    uint8_t *tab = ctx->dTable.as<uint8_t>();
    FillOut *fill = ctx->dFill.as<FillOut>();
    const size_t smem = (size_t)4 * S * 4;
    // exact-occurrence shortcut (k_dp_exact): MP_DP_EXACT=0 sends every task through the DP kernels
    static const bool useExact = !(getenv("MP_DP_EXACT") && getenv("MP_DP_EXACT")[0] == '0');
    const uint32_t perMax = std::min<uint32_t>(per, nTasks);
    if (useExact && (ctx->dExFlag.reserve(((size_t)perMax + 1) * 4) || ctx->dExPos.reserve(((size_t)perMax + 1) * 4) ||
                     ctx->dExIdx.reserve(((size_t)perMax + 1) * 4) || ctx->dCounters.reserve(16 * 8))) return MP_ERR_CUDA;
    for (uint32_t base = 0; base < nTasks; base += per) {
        const uint32_t n = std::min<uint32_t>(per, nTasks - base);
        dim3 gridF((n + 7) / 8), gridT((n + 127) / 128), block(128);
        const uint32_t *active = nullptr, *nActive = nullptr;
#define LAUNCH_EXACT(KK) do { \
        cudaEvent_t stop_ = ctx->ev_begin(2); \
        MP_CUDA(cudaMemsetAsync(ctx->dExFlag.p, 0, ((size_t)n + 1) * 4, ctx->stream)); \
        (++g_mp_launches), k_dp_exact<KK><<<(n + 3) / 4, block, 0, ctx->stream>>>(dRef, dRefLens, refStride, dRead, dReadLens, readStride, dCutoffs, \
            base, n, P, dOuts, dPatterns, patStride, ctx->dExFlag.as<uint32_t>()); \
        { size_t tb_ = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb_, ctx->dExFlag.as<uint32_t>(), ctx->dExPos.as<uint32_t>(), (int64_t)n + 1, ctx->stream); \
          if (ctx->dScanTmp.reserve(tb_)) return MP_ERR_CUDA; \
          cub::DeviceScan::ExclusiveSum(ctx->dScanTmp.p, tb_, ctx->dExFlag.as<uint32_t>(), ctx->dExPos.as<uint32_t>(), (int64_t)n + 1, ctx->stream); } \
        (++g_mp_launches), k_dp_compact<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->dExFlag.as<uint32_t>(), ctx->dExPos.as<uint32_t>(), n, base, \
            dRefLens, dReadLens, ctx->dExIdx.as<uint32_t>(), ctx->dCounters.as<unsigned long long>()); \
        active = ctx->dExIdx.as<uint32_t>(); nActive = ctx->dExPos.as<uint32_t>() + n; \
        ctx->ev_end(stop_); } while (0)
#define LAUNCH(KK) do { \
        if (useExact) LAUNCH_EXACT(KK); \
        cudaEvent_t stop_ = ctx->ev_begin(0); \
        if (P.mismatch == -2 && P.open == -3) \
            (++g_mp_launches), k_dp_fill<KK, -2, -3><<<gridF, block, smem, ctx->stream>>>(dRef, dRefLens, refStride, dRead, dReadLens, readStride, dCutoffs, \
                base, n, P, tab, tableStride, S, fill, active, nActive); \
        else \
            (++g_mp_launches), k_dp_fill<KK, 0, 0><<<gridF, block, smem, ctx->stream>>>(dRef, dRefLens, refStride, dRead, dReadLens, readStride, dCutoffs, \
                base, n, P, tab, tableStride, S, fill, active, nActive); \
        ctx->ev_end(stop_); stop_ = ctx->ev_begin(1); \
        if ((patStride & 3) == 0 && ((uintptr_t)dPatterns & 3) == 0) \
            (++g_mp_launches), k_dp_tb<KK, true><<<gridT, block, 0, ctx->stream>>>(dRef, dRefLens, refStride, dRead, dReadLens, readStride, dCutoffs, \
                base, n, P, tab, tableStride, S, fill, dOuts, dPatterns, patStride, active, nActive); \
        else \
            (++g_mp_launches), k_dp_tb<KK, false><<<gridT, block, 0, ctx->stream>>>(dRef, dRefLens, refStride, dRead, dReadLens, readStride, dCutoffs, \
                base, n, P, tab, tableStride, S, fill, dOuts, dPatterns, patStride, active, nActive); \
        ctx->ev_end(stop_); } while (0)
        if (K == 5) LAUNCH(5); else if (K == 8) LAUNCH(8); else LAUNCH(10);
#undef LAUNCH
#undef LAUNCH_EXACT
        MP_CUDA(cudaGetLastError());
    }
    return 0;
}

int mpd_run_explicit(mp_context *ctx, const uint8_t *dRef, const uint32_t *dRefLens, uint32_t maxRefLen,
                     const uint8_t *dRead, const uint32_t *dReadLens, uint32_t maxReadLen, const int32_t *dCutoffs,
                     uint32_t nTasks, const MpDpParams &P, MpDpOut *dOuts, uint8_t *dPatterns, uint32_t patStride)
{
    return launch_dp(ctx, dRef, dRefLens, maxRefLen, dRead, dReadLens, maxReadLen, dCutoffs, nTasks, maxRefLen, maxReadLen, P,
                     dOuts, dPatterns, patStride);
}

int mpd_run_tasks(mp_context *ctx, const MpDpTask *dTasks, uint32_t nTasks, uint32_t maxRefLen, uint32_t maxReadLen,
                  const MpDpParams &P, MpDpOut *dOuts, uint8_t *dPatterns, uint32_t patStride)
{
    if (nTasks == 0) return 0;
    size_t refB = ((size_t)nTasks * maxRefLen + 15) & ~(size_t)15, readB = (size_t)nTasks * maxReadLen;
    if (ctx->dRefSeq.reserve(refB + (size_t)nTasks * 12) || ctx->dReadSeq.reserve(readB)) return MP_ERR_CUDA;
    uint8_t *dRef = ctx->dRefSeq.as<uint8_t>();
    uint32_t *dRefLens = (uint32_t *)(dRef + refB);
    uint32_t *dReadLens = dRefLens + nTasks;
    int32_t *dCutoffs = (int32_t *)(dReadLens + nTasks);
    (++g_mp_launches), k_extract<<<nTasks, 64, 0, ctx->stream>>>(ctx->ix, ctx->dReads.as<uint32_t>(), ctx->wpq, dTasks, nTasks, dRef, maxRefLen,
                                             ctx->dReadSeq.as<uint8_t>(), maxReadLen, dRefLens, dReadLens, dCutoffs);
    MP_CUDA(cudaGetLastError());
    return launch_dp(ctx, dRef, dRefLens, maxRefLen, ctx->dReadSeq.as<uint8_t>(), dReadLens, maxReadLen, dCutoffs, nTasks,
                     maxRefLen, maxReadLen, P, dOuts, dPatterns, patStride);
}
