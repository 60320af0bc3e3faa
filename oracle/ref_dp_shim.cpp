// oracle/ref_dp_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin C entry point around the reference's own callDP() (soap4/CPU_DP.cpp:881-978),
// compiled together with that file (from where it lies) into oracle/_ref/libref_dp.so.
// It only converts plain per-task byte arrays into the reference's 32-task interleaved
// 2-bit layout (the layout PairEndAlgnBatch::packRead/repackDNA produce,
// soap4/DV-DPfunctions.cpp:3009-3073) and allocates the aligned scratch tables.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <immintrin.h>
typedef unsigned int uint;
typedef unsigned char uchar;
#include "CPU_DP.h"

extern "C" int ref_dp_lanes() {
#ifdef __AVX2__
    return 32;
#else
    return 16;
#endif
}

extern "C" int ref_dp_run(int n, int maxDNALength, int maxReadLength,
                          const uint8_t *ref, const uint32_t *dnaLens,
                          const uint8_t *reads, const uint32_t *readLens,
                          int clipLt, int clipRt, int mismatch, int gapOpen,
                          int *scores, uint32_t *hitLocs, uint32_t *counts, uint8_t *pattern)
{
    const int wDNA = (maxDNALength + 15) >> 4, wRead = (maxReadLength + 15) >> 4;
    const int nPad = ((n + 31) / 32) * 32;
    std::vector<uint> pDNA((size_t)nPad * wDNA, 0), pRead((size_t)nPad * wRead, 0);
    std::vector<uint> dl(nPad, 0), rl(nPad, 0), hl(nPad, 0), mc(nPad, 0), al(nPad, maxDNALength), ar(nPad, 0);
    std::vector<int> co(nPad, 0), sc(nPad, 0);
    std::vector<uchar> pat((size_t)nPad * (maxDNALength + maxReadLength), 0);
    for (int t = 0; t < n; ++t) {
        size_t dT = (size_t)(t / 32) * 32 * wDNA + (t % 32);
        size_t rT = (size_t)(t / 32) * 32 * wRead + (t % 32);
        dl[t] = dnaLens[t]; rl[t] = readLens[t];
        co[t] = (int)std::max(readLens[t] * 0.2, 30.0);
        for (uint i = 1; i <= dnaLens[t]; ++i)
            pDNA[dT + ((i >> 4) << 5)] |= (uint)(ref[(size_t)t * maxDNALength + i - 1] & 3) << ((15 - (i & 15)) << 1);
        for (uint i = 1; i <= readLens[t]; ++i)
            pRead[rT + ((i >> 4) << 5)] |= (uint)(reads[(size_t)t * maxReadLength + i - 1] & 3) << ((15 - (i & 15)) << 1);
    }
    void *table = _mm_malloc((size_t)(maxDNALength + 2) * (maxReadLength + 1) * sizeof(__m256i), 32);
    __m256i *tmp = (__m256i *)_mm_malloc((size_t)(maxReadLength + 1) * sizeof(__m256i), 32);
    callDP(pDNA.data(), dl.data(), maxDNALength, pRead.data(), rl.data(), maxReadLength,
           clipLt, clipRt, al.data(), ar.data(), 1, (uint)mismatch, (uint)gapOpen, (uint)-1,
           co.data(), table, tmp, n, sc.data(), hl.data(), mc.data(), pat.data());
    _mm_free(table); _mm_free(tmp);
    for (int t = 0; t < n; ++t) { scores[t] = sc[t]; hitLocs[t] = hl[t]; counts[t] = mc[t]; }
    memcpy(pattern, pat.data(), (size_t)n * (maxDNALength + maxReadLength));
    return 0;
}
