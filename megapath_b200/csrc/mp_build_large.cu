// mp_build_large.cu -- FM-index construction for texts of 2^32 bases and more (the NT-scale references of MegaPath: BASELINE
// config 3), where neither a 32-bit suffix array nor a whole-text rank array fits the prefix-doubling builder of mp_build.cu.
// Replaces the same reference code (2bwt-lib/BWTConstruct.c, 64-bit throughout) and writes the same files.
//
// The suffix array is never materialised.  Suffixes are handled one BUCKET at a time (bucket = first B symbols, B = 3 or 4, "past
// the end" being a symbol of its own that sorts first):
//   collect   one scan of the packed text gathers the bucket's suffix positions (64-bit)
//   sort      radix sort by the next 21 symbols (63-bit keys); groups of equal keys are refined by the following 21 symbols, and so
//             on: per round one 64-bit sort by the new key and one stable 32-bit sort by the group's first slot, so that elements stay
//             inside their group; resolved suffixes drop out.  Random and near-duplicate texts (few-percent divergence) need a handful
//             of rounds; the number of rounds grows with the longest exact repeat / 21, which is why texts below 2^32 keep the
//             prefix-doubling builder
//   emit      BWT symbol text[SA-1] of every suffix into a 2-bit array indexed by SA index, every 16th SA value (the .sa samples),
//             the SA index of suffix 0 (inverseSa0)
// then, as in mp_build.cu: the '$'-less BWT words, symbol counts, occurrence blocks, LKT.
// Peak memory: the text + n/4 bytes of BWT + n/2 bytes of SA samples + ~60 bytes per suffix of ONE bucket (n/64 .. n/256 suffixes).
#include "mp_context.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <vector>
#include <algorithm>

int mpi_relayout_words(mp_context *ctx, const uint32_t *dWords, uint64_t n);

namespace {

__device__ __forceinline__ uint32_t sym_code(const uint8_t *__restrict__ pac, uint64_t n, uint64_t p)      // 1 + base, 0 past the end
{
    return p < n ? ((pac[p >> 2] >> ((3 - (p & 3)) << 1)) & 3u) + 1u : 0u;
}
__device__ __forceinline__ uint32_t bucket_of(const uint8_t *__restrict__ pac, uint64_t n, uint64_t i, int B)
{
    uint32_t id = 0;
    for (int s = 0; s < B; ++s) id = id * 5 + sym_code(pac, n, i + s);
    return id;
}
__device__ __forceinline__ uint64_t key21(const uint8_t *__restrict__ pac, uint64_t n, uint64_t p)
{
    uint64_t key = 0;
#pragma unroll
    for (int s = 0; s < 21; ++s) key = (key << 3) | sym_code(pac, n, p + s);
    return key;
}

__global__ void k_bucket_hist(const uint8_t *__restrict__ pac, uint64_t n, int B, int nBuckets, unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned int sh[];
    for (int k = threadIdx.x; k < nBuckets; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) atomicAdd(&sh[bucket_of(pac, n, i, B)], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < nBuckets; k += blockDim.x) if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}
// suffix positions of one bucket (any order) + their first sort key (the 21 symbols behind the bucket prefix)
__global__ void k_bucket_collect(const uint8_t *__restrict__ pac, uint64_t n, int B, uint32_t bucket, uint64_t *__restrict__ suf,
                                 uint64_t *__restrict__ key, unsigned int *__restrict__ cursor)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 <= n; i0 += stride) {
        const uint64_t i = i0 + threadIdx.x;
        const bool mine = i <= n && bucket_of(pac, n, i, B) == bucket;
        const unsigned ball = __ballot_sync(0xffffffffu, mine);
        if (ball) {
            unsigned base = 0;
            const int lane = threadIdx.x & 31;
            if (lane == 0) base = atomicAdd(cursor, (unsigned)__popc(ball));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (mine) { const unsigned slot = base + __popc(ball & ((1u << lane) - 1u)); suf[slot] = i; key[slot] = key21(pac, n, i + B); }
        }
    }
}
__global__ void k_iota(uint32_t *__restrict__ v, uint32_t m) { const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; if (t < m) v[t] = t; }
__global__ void k_next_keys(const uint8_t *__restrict__ pac, uint64_t n, const uint64_t *__restrict__ suf, uint32_t m, uint64_t h, uint64_t *__restrict__ key)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) key[t] = key21(pac, n, suf[t] + h);
}
__global__ void k_gather_u32(const uint32_t *__restrict__ src, const uint32_t *__restrict__ perm, uint32_t m, uint32_t *__restrict__ dst)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) dst[t] = src[perm[t]];
}
// after both sorts: element t of the final order is perm[t].  oldHead[t] = t + 1 where a group of the previous round starts
// (to be max-scanned into "first position of my old group"), head[t] = 1 where a group of this round starts.
__global__ void k_round_heads(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ grp, const uint64_t *__restrict__ key, uint32_t m,
                              uint32_t *__restrict__ oldHead, uint32_t *__restrict__ head)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const uint32_t e = perm[t];
    bool oh = t == 0, nh = t == 0;
    if (t) { const uint32_t f = perm[t - 1]; oh = grp[e] != grp[f]; nh = oh || key[e] != key[f]; }
    oldHead[t] = oh ? t + 1 : 0;
    head[t] = nh ? 1u : 0u;
}
struct MaxU32L { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };
// slot of every element = first slot of its old group + rank inside it; new group id = slot of the new group's first element;
// the bucket's suffix array takes the element; unresolved = member of a group of two or more
__global__ void k_round_place(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ grp, const uint64_t *__restrict__ suf,
                              const uint32_t *__restrict__ firstOfOld, const uint32_t *__restrict__ head, uint32_t m,
                              uint64_t *__restrict__ saLocal, uint32_t *__restrict__ slotOut, uint32_t *__restrict__ unres)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const uint32_t e = perm[t];
    const uint32_t slot = grp[e] + (t - (firstOfOld[t] - 1));
    saLocal[slot] = suf[e];
    slotOut[t] = head[t] ? slot + 1 : 0;                      // max-scanned into the new group id + 1
    const bool single = head[t] && (t + 1 == m || head[t + 1]);
    unres[t] = single ? 0u : 1u;
}
__global__ void k_round_compact(const uint32_t *__restrict__ perm, const uint64_t *__restrict__ suf, const uint32_t *__restrict__ newGrp1,
                                const uint32_t *__restrict__ unres, const uint32_t *__restrict__ dst, uint32_t m,
                                uint64_t *__restrict__ sufOut, uint32_t *__restrict__ grpOut)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m || !unres[t]) return;
    sufOut[dst[t]] = suf[perm[t]];
    grpOut[dst[t]] = newGrp1[t] - 1;
}
// bucket done: BWT symbols (2 bit, indexed by SA index, '$' position included), SA samples, inverseSa0
__global__ void k_bucket_emit(const uint8_t *__restrict__ pac, const uint64_t *__restrict__ saLocal, uint32_t nb, uint64_t off,
                              uint32_t *__restrict__ rawBits, uint64_t *__restrict__ saSamples, uint32_t saShift, unsigned long long *__restrict__ isa0)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    const uint64_t i = saLocal[j], g = off + j;
    uint32_t sym = 0;
    if (i == 0) *isa0 = g; else sym = (pac[(i - 1) >> 2] >> ((3 - ((i - 1) & 3)) << 1)) & 3u;
    if (sym) atomicOr(&rawBits[g >> 4], sym << ((15 - (uint32_t)(g & 15)) << 1));
    if ((g & ((1ull << saShift) - 1)) == 0) saSamples[g >> saShift] = g == 0 ? ~0ull : i;       // "saValue[0] = -1" (BWT.c:241)
}
// BWT words without the '$' position (BWT.c:132-157)
__global__ void k_drop_dollar(const uint32_t *__restrict__ rawBits, uint64_t n, uint64_t inverseSa0, uint32_t *__restrict__ words, uint64_t nWords)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nWords) return;
    uint32_t word = 0;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const uint64_t jj = w * 16 + s;
        if (jj < n) {
            const uint64_t j = jj + (jj >= inverseSa0);
            word |= ((rawBits[j >> 4] >> ((15 - (uint32_t)(j & 15)) << 1)) & 3u) << ((15 - s) << 1);
        }
    }
    words[w] = word;
}
__global__ void k_count_syms_l(const uint32_t *__restrict__ words, uint64_t nWords, unsigned long long *__restrict__ cnt)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (w < nWords) { const uint32_t x = words[w]; c0 = mp_word_count(x, 0, 16); c1 = mp_word_count(x, 1, 16); c2 = mp_word_count(x, 2, 16); c3 = mp_word_count(x, 3, 16); }
    c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1);
    c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
    if ((threadIdx.x & 31) == 0) {
        if (c0) atomicAdd(&cnt[0], (unsigned long long)c0);
        if (c1) atomicAdd(&cnt[1], (unsigned long long)c1);
        if (c2) atomicAdd(&cnt[2], (unsigned long long)c2);
        if (c3) atomicAdd(&cnt[3], (unsigned long long)c3);
    }
}
__global__ void k_lkt_hist_l(const uint8_t *__restrict__ pac, uint64_t n, unsigned long long *__restrict__ hist)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t key = 0;
#pragma unroll
    for (int s = 0; s < 13; ++s) { const uint64_t p = i + s; key = (key << 2) | (p < n ? ((pac[p >> 2] >> ((3 - (p & 3)) << 1)) & 3u) : 0u); }
    atomicAdd(&hist[key], 1ull);
}

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

// ctx->dPac holds the text.  Fills the resident index (blocks, SA samples, LKT) like mp_index_build's 32-bit path.
int mpb_build_large(mp_context *ctx, uint64_t n)
{
    cudaStream_t st = ctx->stream;
    const uint8_t *pac = ctx->dPac.as<uint8_t>();
    const uint64_t N = n + 1;
    const int B = n < (1ull << 34) ? 3 : 4;
    int nBuckets = 1; for (int s = 0; s < B; ++s) nBuckets *= 5;
    // SA sampling of the RESIDENT index: every lookup walks LF steps until it reaches a sampled SA index (geometric, mean interval - 1
    // dependent 64-byte gathers), so the interval is as small as HBM allows: 4 when n/4 samples of 8 bytes fit in 30 % of what is free
    // (8 Gbp: 16 GB), else 8, else the file format's 16.  mp_index_save subsamples to the 1/16 the .sa file holds.
    uint32_t saShift = 4;
    {
        size_t freeB = 0, totalB = 0; cudaMemGetInfo(&freeB, &totalB);
        for (uint32_t sh = 2; sh <= 4; ++sh) if ((double)((n >> sh) + 2) * 8.0 <= 0.30 * (double)freeB) { saShift = sh; break; }
        if (const char *e = getenv("MP_SA_SHIFT")) { const int v = atoi(e); if (v >= 0 && v <= 4) saShift = (uint32_t)v; }
    }
    DevBuf dHist, dRaw, dIsa, bSuf, bSuf2, bKey, bKey2, bPerm, bPerm2, bGrp, bGrp2, bG2, bG2b, bA, bB, bC, bD, bSa, sortTmp, scanTmp, dCur;
    auto fail = [&](int rc) { for (DevBuf *b : { &dHist, &dRaw, &dIsa, &bSuf, &bSuf2, &bKey, &bKey2, &bPerm, &bPerm2, &bGrp, &bGrp2, &bG2, &bG2b, &bA, &bB, &bC, &bD, &bSa, &sortTmp, &scanTmp, &dCur }) b->release(); return rc; };
    // ---- bucket sizes ----
    if (dHist.reserve((size_t)nBuckets * 8) || dCur.reserve(16) || dIsa.reserve(8)) return fail(MP_ERR_CUDA);
    MP_CUDA(cudaMemsetAsync(dHist.p, 0, (size_t)nBuckets * 8, st));
    ++g_mp_launches; k_bucket_hist<<<148 * 8, 256, (size_t)nBuckets * 4, st>>>(pac, n, B, nBuckets, dHist.as<unsigned long long>());
    std::vector<unsigned long long> hist(nBuckets);
    MP_CUDA(cudaMemcpyAsync(hist.data(), dHist.p, (size_t)nBuckets * 8, cudaMemcpyDeviceToHost, st));
    MP_CUDA(cudaStreamSynchronize(st));
    unsigned long long maxB = 0, tot = 0;
    for (unsigned long long h : hist) { maxB = std::max(maxB, h); tot += h; }
    if (tot != N) { mp_set_error("mp_index_build: bucket histogram does not add up"); return fail(MP_ERR_STATE); }
    if (maxB >= 0xFFFFFFF0ull) { mp_set_error("mp_index_build: a suffix bucket holds %llu suffixes (32-bit bucket indexing)", maxB); return fail(MP_ERR_CAPACITY); }
    // ---- outputs that live for the whole build ----
    const uint64_t rawWords = (N + 15) / 16, nSa = (n >> saShift) + 1;
    if (dRaw.reserve(rawWords * 4) || ctx->dSa.reserve(nSa * 8)) return fail(MP_ERR_CUDA);
    MP_CUDA(cudaMemsetAsync(dRaw.p, 0, rawWords * 4, st));
    MP_CUDA(cudaMemsetAsync(dIsa.p, 0xFF, 8, st));
    // ---- per-bucket work space ----
    const size_t mb = (size_t)maxB;
    if (bSuf.reserve(mb * 8) || bSuf2.reserve(mb * 8) || bKey.reserve(mb * 8) || bKey2.reserve(mb * 8) || bPerm.reserve(mb * 4) || bPerm2.reserve(mb * 4) ||
        bGrp.reserve(mb * 4) || bGrp2.reserve(mb * 4) || bG2.reserve(mb * 4) || bG2b.reserve(mb * 4) || bA.reserve(mb * 4) || bB.reserve(mb * 4) ||
        bC.reserve(mb * 4) || bD.reserve((mb + 1) * 4) || bSa.reserve(mb * 8)) return fail(MP_ERR_CUDA);
    size_t sortBytes = 0, b2 = 0, scanBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, bKey.as<uint64_t>(), bKey2.as<uint64_t>(), bPerm.as<uint32_t>(), bPerm2.as<uint32_t>(), (int64_t)mb, 0, 63, st);
    cub::DeviceRadixSort::SortPairs(nullptr, b2, bG2.as<uint32_t>(), bG2b.as<uint32_t>(), bPerm.as<uint32_t>(), bPerm2.as<uint32_t>(), (int64_t)mb, 0, 32, st);
    sortBytes = std::max(sortBytes, b2);
    cub::DeviceScan::InclusiveScan(nullptr, scanBytes, bA.as<uint32_t>(), bA.as<uint32_t>(), MaxU32L(), (int64_t)mb, st);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, bC.as<uint32_t>(), bD.as<uint32_t>(), (int64_t)mb + 1, st);
    scanBytes = std::max(scanBytes, b2);
    if (sortTmp.reserve(sortBytes) || scanTmp.reserve(scanBytes)) return fail(MP_ERR_CUDA);

    uint64_t off = 0;
    for (int bucket = 0; bucket < nBuckets; ++bucket) {
        const uint32_t nb = (uint32_t)hist[bucket];
        if (nb == 0) continue;
        MP_CUDA(cudaMemsetAsync(dCur.p, 0, 4, st));
        ++g_mp_launches; k_bucket_collect<<<148 * 16, 256, 0, st>>>(pac, n, B, (uint32_t)bucket, bSuf.as<uint64_t>(), bKey.as<uint64_t>(), dCur.as<unsigned int>());
        MP_CUDA(cudaMemsetAsync(bGrp.p, 0, (size_t)nb * 4, st));               // one group: the whole bucket, first slot 0
        uint64_t *suf = bSuf.as<uint64_t>(), *sufAlt = bSuf2.as<uint64_t>();
        uint32_t *grp = bGrp.as<uint32_t>(), *grpAlt = bGrp2.as<uint32_t>();
        uint32_t m = nb;
        uint64_t hoff = (uint64_t)B;                                           // offset of the symbols this round's key holds
        for (int round = 0; m; ++round) {
            if (round > 1000000) { mp_set_error("mp_index_build: suffix refinement did not converge"); return fail(MP_ERR_STATE); }
            const unsigned g = grid_for(m, 256);
            if (round) { hoff += 21; ++g_mp_launches; k_next_keys<<<g, 256, 0, st>>>(pac, n, suf, m, hoff, bKey.as<uint64_t>()); }
            ++g_mp_launches; k_iota<<<g, 256, 0, st>>>(bPerm.as<uint32_t>(), m);
            // order by the new key, then (stable) by group: elements stay inside their group, ordered by key
            cub::DeviceRadixSort::SortPairs(sortTmp.p, sortBytes, bKey.as<uint64_t>(), bKey2.as<uint64_t>(), bPerm.as<uint32_t>(), bPerm2.as<uint32_t>(), (int64_t)m, 0, 63, st);
            uint32_t *perm = bPerm2.as<uint32_t>();
            if (round) {
                ++g_mp_launches; k_gather_u32<<<g, 256, 0, st>>>(grp, perm, m, bG2.as<uint32_t>());
                cub::DeviceRadixSort::SortPairs(sortTmp.p, sortBytes, bG2.as<uint32_t>(), bG2b.as<uint32_t>(), bPerm2.as<uint32_t>(), bPerm.as<uint32_t>(), (int64_t)m, 0, 32, st);
                perm = bPerm.as<uint32_t>();
            }
            ++g_mp_launches; k_round_heads<<<g, 256, 0, st>>>(perm, grp, bKey.as<uint64_t>(), m, bA.as<uint32_t>(), bB.as<uint32_t>());
            cub::DeviceScan::InclusiveScan(scanTmp.p, scanBytes, bA.as<uint32_t>(), bA.as<uint32_t>(), MaxU32L(), (int64_t)m, st);
            ++g_mp_launches; k_round_place<<<g, 256, 0, st>>>(perm, grp, suf, bA.as<uint32_t>(), bB.as<uint32_t>(), m, bSa.as<uint64_t>(), bG2.as<uint32_t>(), bC.as<uint32_t>());
            cub::DeviceScan::InclusiveScan(scanTmp.p, scanBytes, bG2.as<uint32_t>(), bG2.as<uint32_t>(), MaxU32L(), (int64_t)m, st);
            MP_CUDA(cudaMemsetAsync(bC.as<uint32_t>() + m, 0, 4, st));
            cub::DeviceScan::ExclusiveSum(scanTmp.p, scanBytes, bC.as<uint32_t>(), bD.as<uint32_t>(), (int64_t)m + 1, st);
            uint32_t m2 = 0;
            MP_CUDA(cudaMemcpyAsync(&m2, bD.as<uint32_t>() + m, 4, cudaMemcpyDeviceToHost, st));
            MP_CUDA(cudaStreamSynchronize(st));
            if (m2) { ++g_mp_launches; k_round_compact<<<g, 256, 0, st>>>(perm, suf, bG2.as<uint32_t>(), bC.as<uint32_t>(), bD.as<uint32_t>(), m, sufAlt, grpAlt); }
            std::swap(suf, sufAlt); std::swap(grp, grpAlt);
            m = m2;
        }
        ++g_mp_launches; k_bucket_emit<<<grid_for(nb, 256), 256, 0, st>>>(pac, bSa.as<uint64_t>(), nb, off, dRaw.as<uint32_t>(), ctx->dSa.as<uint64_t>(), saShift,
                                                                         dIsa.as<unsigned long long>());
        MP_CUDA(cudaGetLastError());
        off += nb;
    }
    unsigned long long inverseSa0 = 0;
    MP_CUDA(cudaMemcpyAsync(&inverseSa0, dIsa.p, 8, cudaMemcpyDeviceToHost, st));
    MP_CUDA(cudaStreamSynchronize(st));
    for (DevBuf *b : { &bSuf, &bSuf2, &bKey, &bKey2, &bPerm, &bPerm2, &bGrp, &bGrp2, &bG2, &bG2b, &bA, &bB, &bC, &bD, &bSa, &sortTmp }) b->release();
    if (inverseSa0 > n) { mp_set_error("mp_index_build: suffix 0 was not placed"); return fail(MP_ERR_STATE); }
    // ---- '$'-less BWT words padded to whole occ blocks, symbol counts ----
    const uint64_t nBlocks = n / MP_BLK_SYMS + 1, nWords = (n + 15) / 16;
    DevBuf dWords;
    if (dWords.reserve(nBlocks * 48)) return fail(MP_ERR_CUDA);
    MP_CUDA(cudaMemsetAsync(dWords.p, 0, nBlocks * 48, st));
    ++g_mp_launches; k_drop_dollar<<<grid_for(nWords, 256), 256, 0, st>>>(dRaw.as<uint32_t>(), n, inverseSa0, dWords.as<uint32_t>(), nWords);
    unsigned long long *dCnt = dHist.as<unsigned long long>();
    MP_CUDA(cudaMemsetAsync(dCnt, 0, 32, st));
    ++g_mp_launches; k_count_syms_l<<<grid_for(nWords, 256), 256, 0, st>>>(dWords.as<uint32_t>(), nWords, dCnt);
    unsigned long long hc[4];
    MP_CUDA(cudaMemcpyAsync(hc, dCnt, 32, cudaMemcpyDeviceToHost, st));
    MP_CUDA(cudaStreamSynchronize(st));
    dRaw.release();
    hc[0] -= nWords * 16 - n;                        // zero padding of the last word was counted as 'A'
    ctx->ix.n = n; ctx->ix.inverseSa0 = inverseSa0;
    ctx->ix.cum[0] = 0;
    for (int c = 0; c < 4; ++c) ctx->ix.cum[c + 1] = ctx->ix.cum[c] + hc[c];
    if (ctx->ix.cum[4] != n) { dWords.release(); mp_set_error("mp_index_build: symbol counts do not add up"); return fail(MP_ERR_STATE); }
    ctx->ix.sa = ctx->dSa.as<uint64_t>(); ctx->ix.saShift = saShift; ctx->saInterval = 1ull << saShift;
    ctx->ix.sa32 = nullptr; ctx->dSa32.release();
    if (int rc = mpi_relayout_words(ctx, dWords.as<uint32_t>(), n)) { dWords.release(); return fail(rc); }
    dWords.release();
    // ---- LKT: inclusive cumulative 13-mer counts ----
    const uint64_t nLkt = 1ull << 26;
    if (ctx->dLkt.reserve(nLkt * 8)) return fail(MP_ERR_CUDA);
    MP_CUDA(cudaMemsetAsync(ctx->dLkt.p, 0, nLkt * 8, st));
    ++g_mp_launches; k_lkt_hist_l<<<grid_for(n, 256), 256, 0, st>>>(pac, n, ctx->dLkt.as<unsigned long long>());
    {
        size_t tb = 0;
        cub::DeviceScan::InclusiveSum(nullptr, tb, ctx->dLkt.as<uint64_t>(), ctx->dLkt.as<uint64_t>(), (int64_t)nLkt, st);
        if (scanTmp.reserve(tb)) return fail(MP_ERR_CUDA);
        cub::DeviceScan::InclusiveSum(scanTmp.p, tb, ctx->dLkt.as<uint64_t>(), ctx->dLkt.as<uint64_t>(), (int64_t)nLkt, st);
    }
    MP_CUDA(cudaGetLastError());
    MP_CUDA(cudaStreamSynchronize(st));
    fail(0);
    ctx->ix.lkt = ctx->dLkt.as<uint64_t>();
    ctx->ix.pac = pac;
    if (int rc = mpi_finish_sa(ctx)) return rc;
    ctx->hbmBytes = ctx->dBlocks.cap + ctx->dSuper.cap + ctx->dSa.cap + ctx->dSa40Lo.cap + ctx->dSa40Hi.cap + ctx->dLkt.cap + ctx->dPac.cap;
    ctx->hasIndex = true; ctx->hasBatch = false; ctx->seeded = false;
    return 0;
}
