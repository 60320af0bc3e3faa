// mp_index.cuh -- HBM-resident FM-index of the soap4 hot path (device view + primitives).
//
// Replaces the in-memory 2BWT index of the reference (2bwt-lib/BWT.h:72-98, 2bwt-flex/LT.h:51-55):
// BWTOccValue / BWTOccValueExplicit / BWTDecode (2bwt-lib/BWT.c:597, 783, 328),
// BWTSaValue / BWTPsiMinusValue / BWTOccValueOnSpot (BWT.c:968, 915, 689), LT lookups
// (DV-DPfunctions.cpp:2240-2241).
//
// Layout (DESIGN.md "Index in HBM"):
//   occ blocks : 64 bytes = one 2-sector aligned fetch per Occ() evaluation
//                  [ 0..15]  u32 cnt[4]  count of symbol c in BWT[0, 192*b) minus the
//                            superblock base (superblock = 2^24 blocks, u64 x4, L1/L2 resident)
//                  [16..63]  192 BWT symbols, 2 bit, 16 per u32, first symbol in the top bits
//                            (the .bwt word order, so block b is .bwt words [12b, 12b+12))
//   sa         : sampled suffix array, one value per saInterval SA indices (by index), as in the .sa file
//   sa32       : the full suffix array as u32 when the text is shorter than 2^32 (12.4 GB for 3.1 Gbp, HBM has
//                180 GB): an SA lookup is then one gather instead of a geometric(1/16) LF walk
//   sa40lo/hi  : texts of 2^32 bases and more: 40-bit samples (u32 + u8 arrays) at the densest rate that fits the HBM
//                budget -- every index (5 bytes per base: 40 GB for 8 Gbp) down to every 8th (37 GB for 60 Gbp)
//   lkt        : 4^13 cumulative 13-mer counts
//   pac        : packed text, 4 bases per byte, first base in the top 2 bits
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define MP_BLK_SYMS 192u
#define MP_SUPER_SHIFT 24

struct MpIndexView {
    uint64_t n;              // text length
    uint64_t inverseSa0;
    uint64_t cum[5];
    const uint4 *blocks;     // nBlocks * 4 uint4
    uint64_t nBlocks;
    const uint64_t *super;   // (nBlocks >> 24) + 1 rows of 4
    const uint64_t *sa;      // (n + saInterval) / saInterval values, sa[0] = -1
    uint32_t saShift;        // log2(saInterval)
    const uint32_t *sa32;    // optional: the whole suffix array (n + 1 entries, texts < 4.29 Gbp); one gather per SA lookup
    const uint32_t *sa40lo;  // optional (texts >= 2^32): 40-bit samples, one per 2^saShift SA indices, low 32 bits ...
    const uint8_t *sa40hi;   // ... and bits 32..39; they replace `sa` (saShift 0 = the whole suffix array)
    const uint64_t *lkt;     // 4^13
    const uint8_t *pac;
};

// cum[c] without dynamic indexing of kernel parameters (that would force a local copy)
__device__ __forceinline__ uint64_t mp_cum(const MpIndexView &ix, uint32_t c)
{
    return c == 0 ? ix.cum[0] : c == 1 ? ix.cum[1] : c == 2 ? ix.cum[2] : ix.cum[3];
}

// number of symbols == c among the first r (0..16) symbols of a .bwt word
__device__ __forceinline__ uint32_t mp_word_count(uint32_t w, uint32_t c, uint32_t r)
{
    uint32_t x = w ^ (0x55555555u * c);          // pair == 00 where the symbol equals c
    uint32_t m = ~(x | (x >> 1)) & 0x55555555u;  // low bit of each matching pair
    uint32_t keep = r >= 16 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> (2 * r));
    return __popc(m & keep);
}

// Occ over the $-less BWT positions [0, idx) -- one thread, whole 64-byte block.
__device__ __forceinline__ uint64_t mp_occ_raw(const MpIndexView &ix, uint64_t idx, uint32_t c)
{
    uint64_t b = idx / MP_BLK_SYMS;
    uint32_t off = (uint32_t)(idx - b * MP_BLK_SYMS);
    const uint4 *p = ix.blocks + b * 4;
    uint4 h = __ldg(p);
    uint32_t cnt = c == 0 ? h.x : c == 1 ? h.y : c == 2 ? h.z : h.w;
    uint64_t v = ix.super[(b >> MP_SUPER_SHIFT) * 4 + c] + cnt;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int base = 64 * q;
        if ((int)off > base) {
            uint4 w = __ldg(p + 1 + q);
            int r = (int)off - base;
            v += mp_word_count(w.x, c, min(max(r, 0), 16));
            v += mp_word_count(w.y, c, min(max(r - 16, 0), 16));
            v += mp_word_count(w.z, c, min(max(r - 32, 0), 16));
            v += mp_word_count(w.w, c, min(max(r - 48, 0), 16));
        }
    }
    return v;
}

// BWTOccValue (BWT.c:597-634): Occ(c, idx) with the '$' adjustment
__device__ __forceinline__ uint64_t mp_occ(const MpIndexView &ix, uint64_t idx, uint32_t c)
{
    idx -= (idx > ix.inverseSa0);
    return mp_occ_raw(ix, idx, c);
}

// One LF step -- BWTPsiMinusValue (BWT.c:915-938): needs the BWT symbol before the position
// and its rank; both live in the same 64-byte block.
__device__ __forceinline__ uint64_t mp_lf(const MpIndexView &ix, uint64_t i)
{
    if (i == ix.inverseSa0) return 0;
    uint64_t i1 = i + 1;
    i1 -= (i1 > ix.inverseSa0);
    uint64_t p = i1 - 1;                       // position of the symbol
    uint64_t b = p / MP_BLK_SYMS;
    uint32_t off = (uint32_t)(p - b * MP_BLK_SYMS);
    const uint32_t *wp = (const uint32_t *)(ix.blocks + b * 4);
    uint32_t w = __ldg(wp + 4 + (off >> 4));
    uint32_t c = (w >> ((15 - (off & 15)) << 1)) & 3;
    return mp_cum(ix, c) + mp_occ_raw(ix, p, c) + 1;
}

// SA value of an index that has a resident sample (any index when the array is dense; SA[0] reads as -1, BWT.c:241)
__device__ __forceinline__ bool mp_sa_dense(const MpIndexView &ix) { return ix.sa32 != nullptr || (ix.sa40lo != nullptr && ix.saShift == 0); }
__device__ __forceinline__ uint64_t mp_sa_sample(const MpIndexView &ix, uint64_t saIndex)
{
    if (ix.sa32) return saIndex == 0 ? ~0ull : (uint64_t)__ldg(ix.sa32 + saIndex);
    if (ix.sa40lo) {
        const uint64_t k = saIndex >> ix.saShift;
        return saIndex == 0 ? ~0ull : (uint64_t)__ldg(ix.sa40lo + k) | ((uint64_t)__ldg(ix.sa40hi + k) << 32);
    }
    return __ldg(ix.sa + (saIndex >> ix.saShift));
}
// BWTSaValue (BWT.c:968-998)
__device__ __forceinline__ uint64_t mp_sa(const MpIndexView &ix, uint64_t saIndex, uint32_t *steps = nullptr)
{
    // dense SA: any sampling rate gives the same SA[i], this one costs one gather
    uint64_t skipped = 0;
    if (!mp_sa_dense(ix)) {
        const uint64_t mask = (1ull << ix.saShift) - 1;
        while (saIndex & mask) { ++skipped; saIndex = mp_lf(ix, saIndex); }
    }
    if (steps) *steps = (uint32_t)skipped;
    return mp_sa_sample(ix, saIndex) + skipped;
}

__device__ __forceinline__ uint32_t mp_text_base(const MpIndexView &ix, uint64_t pos)
{
    return (__ldg(ix.pac + (pos >> 2)) >> ((3 - (pos & 3)) << 1)) & 3;
}
