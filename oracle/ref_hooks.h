// oracle/ref_hooks.h -- TEST INFRASTRUCTURE ONLY.
//
// Dump hooks force-included (-include) into two reference translation units when
// building oracle/_ref/soap4_dump (see Makefile.ref).  They write the data crossing
// three seams of the reference hot path to $MPH_DUMP_DIR so that the CPU restatement
// (oracle/mp_oracle.cpp) and the CUDA path can be compared seam by seam:
//
//   seedpos.bin  : after PairEndSeedingBatch::mmpSeeding   (DV-DPfunctions.cpp:2404-2615)
//   cand.bin     : after mergeAndPairPairedEnd              (DV-DPfunctions.cpp:2088-2119)
//   dp.bin       : every callDP invocation                  (CPU_DP.cpp:881-978)
//
// All records are little-endian, appended, and self-describing (see tools/refdump.py).
// Nothing here changes what the reference computes.
#ifndef MP_REF_HOOKS_H
#define MP_REF_HOOKS_H
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>

static inline FILE *mph_open(const char *name)
{
    const char *d = getenv("MPH_DUMP_DIR");
    if (!d) return NULL;
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", d, name);
    return fopen(path, "ab");
}

static pthread_mutex_t mph_mutex = PTHREAD_MUTEX_INITIALIZER;

// SeedPos = { u64 pos; u32 strand_readID; u32 paired_seedLength } (SeedPool.h:52-57), 16 bytes.
#define MPH_DUMP_SEEDPOS(rp, nr, mp, nm) do {                                   \
    FILE *f_ = mph_open("seedpos.bin");                                         \
    if (f_) { uint64_t h_[2] = { (uint64_t)(nr), (uint64_t)(nm) };              \
        fwrite(h_, 8, 2, f_);                                                   \
        fwrite((rp), 16, (nr), f_); fwrite((mp), 16, (nm), f_); fclose(f_); }   \
} while (0)

// CandidateInfo (DeepDP_Space, DV-DPfunctions.h:1232-1240): only readIDLeft and pos[2]
// are defined at this seam; written as { u32 readIDLeft; u32 0; u64 pos0; u64 pos1 }.
#define MPH_DUMP_CAND(v) do {                                                   \
    FILE *f_ = mph_open("cand.bin");                                            \
    if (f_) { uint64_t n_ = (v).size(); fwrite(&n_, 8, 1, f_);                  \
        for (uint64_t i_ = 0; i_ < n_; ++i_) {                                  \
            uint32_t a_[2] = { (v)[i_].readIDLeft, 0 };                         \
            uint64_t p_[2] = { (v)[i_].pos[0], (v)[i_].pos[1] };                \
            fwrite(a_, 4, 2, f_); fwrite(p_, 8, 2, f_); }                       \
        fclose(f_); }                                                           \
} while (0)

// One record per callDP invocation:
//   header u32[8] = { magic 'MPDP', n, maxDNALength, maxReadLength, clipLt, clipRt,
//                     mismatch (as passed, two's complement), gapOpen }
//   per task i<n: u32 dnaLen, u32 readLen, i32 cutoff, i32 score, u32 hitLoc, u32 count,
//                 u8 ref[dnaLen] (codes 0..3), u8 read[readLen], u16 patLen, u8 pat[patLen]
//   (pattern only when score >= cutoff, else patLen = 0)
#define MPH_DUMP_DP() do {                                                      \
    if (getenv("MPH_DUMP_DIR")) {                                               \
      pthread_mutex_lock(&mph_mutex);                                           \
      FILE *f_ = mph_open("dp.bin");                                            \
      if (f_) {                                                                 \
        uint32_t h_[8] = { 0x5044504du, numDPInstances, maxDNALength, maxReadLength, \
            (uint32_t)clipLtSizes, (uint32_t)clipRtSizes, MismatchScore, GapOpenScore }; \
        fwrite(h_, 4, 8, f_);                                                   \
        for (uint32_t t_ = 0; t_ < numDPInstances; ++t_) {                      \
            uint32_t dnaTPARA = (t_ >> 5) * (MC_CeilDivide16(maxDNALength) << 5) + (t_ & 0x1F); \
            uint32_t readTPARA = (t_ >> 5) * (MC_CeilDivide16(maxReadLength) << 5) + (t_ & 0x1F); \
            uint32_t r_[6] = { DNALengths[t_], readLengths[t_], (uint32_t)cutOffThresholds[t_], \
                               (uint32_t)scores[t_], hitLocs[t_], maxScoreCounts[t_] }; \
            fwrite(r_, 4, 6, f_);                                               \
            for (uint32_t j_ = 1; j_ <= DNALengths[t_]; ++j_) {                 \
                uint8_t c_ = MC_DnaUnpack(DNASequences, j_); fputc(c_, f_); }   \
            for (uint32_t j_ = 1; j_ <= readLengths[t_]; ++j_) {                \
                uint8_t c_ = MC_ReadUnpack(readSequences, j_); fputc(c_, f_); } \
            uint16_t pl_ = 0;                                                   \
            const uchar *p_ = pattern + (size_t)t_ * (maxDNALength + maxReadLength); \
            if (scores[t_] >= cutOffThresholds[t_]) {                           \
                while (p_[pl_] != 0) { pl_ += (p_[pl_] == 'V') ? 2 : 1; }       \
            }                                                                   \
            fwrite(&pl_, 2, 1, f_); fwrite(p_, 1, pl_, f_);                     \
        }                                                                       \
        fclose(f_); }                                                           \
      pthread_mutex_unlock(&mph_mutex);                                         \
    }                                                                           \
} while (0)

#endif
