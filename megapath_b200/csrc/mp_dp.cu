// mp_dp.cu -- semi-global affine-gap DP with traceback on the device.
// Replaces SemiGlobalAligner::performAlignment -> callDP -> SemiGlobalAlignment ->
// GenerateDPTable + GPUBacktrack (CPU_DPfunctions.cpp:270-312; CPU_DP.cpp:881-978, 788-871,
// 122-619, 622-786) and the task packing of PairEndAlgnBatch::packRead/repackDNA
// (DV-DPfunctions.cpp:3009-3073).
//
// The reference encodes the recurrence as 8-bit saturating deltas with "score trimming";
// observable results equal the plain recurrence below (oracle/mp_oracle_dp.cpp,
// checked against the reference's own callDP).  Integer ALU work, no tensor cores.
//
// k_dp<K>: one warp per task.  Lane l owns read columns l*K+1 .. l*K+K in registers and walks
// the reference rows with a one-row skew per lane (anti-diagonal wavefront); the right-most
// H / I of a lane's strip and the row's reference base travel to lane l+1 in one shuffle.
// Every cell leaves one traceback byte  (H-Hdiag-mm)*42 + (H-Hleft-open)*3 + flag
// (flag 0 = raised by the clip floor, 1 = D==H, 2 = otherwise -- same information the
// reference keeps, CPU_DP.cpp:183, 529-533) in a step-major table so that a warp's stores of
// one step are one contiguous 32*SLOT-byte segment.  The warp then back-tracks its own table.
#include "mp_context.h"

#define DP_NEG (-20000)

template <int K> struct Slot { static const int BYTES = K <= 4 ? 4 : (K <= 8 ? 8 : 16); };

__device__ __forceinline__ int h0_value(int j, int clipLt, int open)       // row 0 (CPU_DP.cpp:397-429)
{
    return j <= clipLt ? 0 : open - (j - clipLt - 1);
}

struct CellInfo { int dd, hd, flag; };   // H - Hdiag, H - Hleft, flag

template <int K>
__device__ __forceinline__ uint8_t load_cell(const uint8_t *__restrict__ tab, int r, int c)
{
    int lane = (c - 1) / K, k = (c - 1) - lane * K;
    return tab[((size_t)(r + lane) * 32 + lane) * Slot<K>::BYTES + k];
}
// flag of any cell including the virtual row 0 / column 0
template <int K>
__device__ __forceinline__ int cell_flag(const uint8_t *__restrict__ tab, int r, int c, int clipLt)
{
    if (c == 0) return 0;                       // column 0 cells are stored as 0 (CPU_DP.cpp:447-450)
    if (r == 0) return c <= clipLt ? 0 : 1;     // row 0 (CPU_DP.cpp:405-427)
    return load_cell<K>(tab, r, c) % 3;
}
// H[r][c] - H[r][c-1] for any row including row 0
template <int K>
__device__ __forceinline__ int cell_hd(const uint8_t *__restrict__ tab, int r, int c, int clipLt, int open)
{
    if (r == 0) return h0_value(c, clipLt, open) - h0_value(c - 1, clipLt, open);
    return open + (int)(load_cell<K>(tab, r, c) / 3 % 14);
}

template <int K>
__global__ void __launch_bounds__(128)
k_dp(const uint8_t *__restrict__ refSeq, const uint32_t *__restrict__ refLens, uint32_t refStride,
     const uint8_t *__restrict__ readSeq, const uint32_t *__restrict__ readLens, uint32_t readStride,
     const int32_t *__restrict__ cutoffs, uint32_t nTasks, MpDpParams P,
     uint8_t *__restrict__ tables, size_t tableStride, MpDpOut *__restrict__ outs,
     uint8_t *__restrict__ patterns, uint32_t patStride)
{
    const int SLOT = Slot<K>::BYTES;
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int mm = P.mismatch, open = P.open, ext = -1, clipLt = P.clipLt;
    uint8_t *tab = tables + (size_t)warpId * tableStride;

    for (uint32_t task = warpId; task < nTasks; task += warpsPerGrid) {
        const int N = (int)refLens[task], L = (int)readLens[task], cutoff = cutoffs[task];
        MpDpOut o; o.score = 0; o.hitLoc = 0; o.count = 0; o.patLen = 0;
        uint8_t *pat = patterns + (size_t)task * patStride;
        // CPU_DP.cpp:296-324: outside these bounds the reference aborts the SIMD group
        if (cutoff > L || cutoff <= 0 || L >= 255 + open - 1 + cutoff || L > 32 * K) {
            if (lane == 0) outs[task] = o;
            continue;
        }
        const uint8_t *rs = readSeq + (size_t)task * readStride;
        const uint8_t *fs = refSeq + (size_t)task * refStride;
        const int j0 = lane * K + 1;
        int rb[K], Hp[K], Dp[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int j = j0 + k;
            rb[k] = j <= L ? (int)rs[j - 1] : 8;
            Hp[k] = h0_value(j, clipLt, open);
            Dp[k] = DP_NEG;
        }
        int prevHleft = h0_value(j0 - 1, clipLt, open);      // H[i-1][j0-1]
        const int minCol = max(L - P.clipRt, 1);
        int best = cutoff - 1, bestRow = 0, bestCol = 0, cnt = 0;
        uint32_t sendHI = 0; int sendRef = 4;
        const int steps = N + 31;
        for (int t = 1; t <= steps; ++t) {
            uint32_t rHI = __shfl_up_sync(0xffffffffu, sendHI, 1);
            int rRef = __shfl_up_sync(0xffffffffu, sendRef, 1);
            const int i = t - lane;
            int Hleft, Il, refc;
            if (lane == 0) { Hleft = 0; Il = DP_NEG; refc = t <= N ? (int)fs[t - 1] : 4; }
            else { Hleft = (int)(int16_t)(rHI & 0xffff); Il = (int)(int16_t)(rHI >> 16); refc = rRef; }
            if (i >= 1 && i <= N) {
                int Hdiag = prevHleft;
                prevHleft = Hleft;
                uint32_t lo = 0, hi = 0, hi2 = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int j = j0 + k;
                    int s = (refc == rb[k]) ? 1 : mm;
                    int d = max(Dp[k] + ext, Hp[k] + open);
                    Il = max(Il + ext, Hleft + open);
                    int h = max(Hdiag + s, max(d, Il));
                    int flag = (d == h) ? 1 : 2;
                    if (j <= clipLt && h < 0) { h = 0; flag = 0; }                   // CPU_DP.cpp:505-510
                    uint32_t code = (uint32_t)((h - Hdiag - mm) * 42 + (h - Hleft - open) * 3 + flag);
                    if (k < 4) lo |= code << (8 * k); else if (k < 8) hi |= code << (8 * (k - 4)); else hi2 |= code << (8 * (k - 8));
                    if (h >= cutoff && j >= minCol && j <= L) {                      // CPU_DP.cpp:545-590
                        if (h > best) { best = h; bestRow = i; bestCol = j; cnt = 1; }
                        else if (h == best) ++cnt;
                    }
                    Hdiag = Hp[k]; Hp[k] = h; Dp[k] = d; Hleft = h;
                }
                uint8_t *dst = tab + ((size_t)t * 32 + lane) * SLOT;
                if (SLOT == 4) *(uint32_t *)dst = lo;
                else if (SLOT == 8) *(uint2 *)dst = make_uint2(lo, hi);
                else *(uint4 *)dst = make_uint4(lo, hi, hi2, 0);
                sendHI = ((uint32_t)Hleft & 0xffffu) | ((uint32_t)max(Il, -20000) << 16);
                sendRef = refc;
            }
        }
        // ---- winner: max score, then smallest (row, col); ties counted (CPU_DP.cpp:569-590) ----
        int gbest = best;
#pragma unroll
        for (int d = 16; d; d >>= 1) gbest = max(gbest, __shfl_xor_sync(0xffffffffu, gbest, d));
        if (gbest < cutoff) { if (lane == 0) outs[task] = o; __syncwarp(); continue; }
        uint32_t key = best == gbest ? ((uint32_t)bestRow << 12) | (uint32_t)bestCol : 0xffffffffu;
        int c2 = best == gbest ? cnt : 0;
#pragma unroll
        for (int d = 16; d; d >>= 1) { key = min(key, __shfl_xor_sync(0xffffffffu, key, d)); c2 += __shfl_xor_sync(0xffffffffu, c2, d); }
        __syncwarp();
        if (lane == 0) {
            // ---- GPUBacktrack (CPU_DP.cpp:622-786), literal ----
            const int hitRow = (int)(key >> 12), hitCol = (int)(key & 0xfff);
            o.score = gbest; o.count = min(c2, 255);
            uint32_t p = 0;
            int clipR = L - hitCol;
            if (clipR > 0) { pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)clipR; }
            int i = L - clipR, j = hitRow;
            enum { NORMAL, I_EXT, D_EXT, SM_EXIT, SI_EXIT, SD_EXIT };
            int state = NORMAL;
            int accum = 0;
            while (i > 0 && j > 0) {
                uint32_t cell = load_cell<K>(tab, j, i);
                int flag = cell % 3;
                int hd = open + (int)(cell / 3 % 14);
                int dd = mm + (int)(cell / 42);
                int vd = dd - cell_hd<K>(tab, j - 1, i, clipLt, open);
                bool eq = fs[j - 1] == rs[i - 1];
                int ms = eq ? 1 : mm;
                if (state == NORMAL) {
                    if (dd == ms && i != 1 && cell_flag<K>(tab, j - 1, i - 1, clipLt) == 0) { state = SM_EXIT; break; }
                    else if (dd == ms) { pat[p++] = eq ? 'M' : 'm'; --j; --i; }
                    else if (flag == 1) {
                        pat[p++] = 'D'; --j;
                        if (vd != open) { accum = (int8_t)(vd - ext); state = D_EXT; }
                    } else {
                        pat[p++] = 'I'; --i;
                        if (hd != open) { accum = (int8_t)(hd - ext); state = I_EXT; }
                    }
                } else if (state == D_EXT) {
                    if (vd + accum == open && cell_flag<K>(tab, j - 1, i, clipLt) == 0) { state = SD_EXIT; break; }
                    pat[p++] = 'D'; --j;
                    if (vd + accum == open) state = NORMAL; else accum = (int8_t)(accum + vd - ext);
                } else {
                    if (hd + accum == open && cell_flag<K>(tab, j, i - 1, clipLt) == 0) { state = SI_EXIT; break; }
                    pat[p++] = 'I'; --i;
                    if (hd + accum == open) state = NORMAL; else accum = (int8_t)(accum + hd - ext);
                }
            }
            bool discard = false;
            if (j == 0) {
                int sc = min(clipLt & 0xff, i);
                if (sc < i) { pat[p++] = 'I'; pat[p++] = 'V'; pat[p++] = (uint8_t)(i - sc); }
                pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)sc;
            } else if (state == SI_EXIT) {
                pat[p++] = 'I'; pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)(i - 1);
            } else if (state == SD_EXIT) {
                pat[p++] = 'D'; pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)(i - 1);
                discard = true;                                   // CPU_DP.cpp:842-857
            } else if (state == SM_EXIT) {
                pat[p++] = (fs[j - 1] == rs[i - 1]) ? 'M' : 'm';
                pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)(i - 1);
                j -= 1;
            }
            pat[p] = 0;
            o.patLen = p;
            if (discard) { o.score = 0; o.hitLoc = 0; } else o.hitLoc = (uint32_t)j;
            outs[task] = o;
        }
        __syncwarp();
    }
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
    int K = maxReadLen <= 128 ? 4 : maxReadLen <= 160 ? 5 : maxReadLen <= 256 ? 8 : 10;
    if (maxReadLen > 320) { mp_set_error("read length %u exceeds the DP kernel bound 320", maxReadLen); return MP_ERR_ARG; }
    int slot = K <= 4 ? 4 : (K <= 8 ? 8 : 16);
    int dev = 0, nSM = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, dev);
    size_t tableStride = ((size_t)maxRefLen + 34) * 32 * slot;
    // persistent grid: 4 warps per CTA, up to 8 CTAs per SM, bounded by the task count and by table memory
    uint32_t warps = (uint32_t)nSM * 8 * 4;
    if (warps > nTasks) warps = (nTasks + 3) / 4 * 4;
    size_t freeB = 0, totalB = 0; cudaMemGetInfo(&freeB, &totalB);
    size_t budget = ctx->dTable.cap > freeB / 2 ? ctx->dTable.cap : freeB / 2;
    while ((size_t)warps * tableStride > budget && warps > 4) warps = (warps / 2 + 3) / 4 * 4;
    if (ctx->dTable.reserve((size_t)warps * tableStride)) return MP_ERR_CUDA;
    dim3 grid(warps / 4), block(128);
    uint8_t *tab = ctx->dTable.as<uint8_t>();
#define LAUNCH(KK) (++g_mp_launches), k_dp<KK><<<grid, block, 0, ctx->stream>>>(dRef, dRefLens, refStride, dRead, dReadLens, readStride, dCutoffs, \
        nTasks, P, tab, tableStride, dOuts, dPatterns, patStride)
    if (K == 4) LAUNCH(4); else if (K == 5) LAUNCH(5); else if (K == 8) LAUNCH(8); else LAUNCH(10);
#undef LAUNCH
    MP_CUDA(cudaGetLastError());
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
