// oracle/mp_oracle.h -- TEST INFRASTRUCTURE ONLY (the checker, never the product).
//
// Plain scalar CPU restatement of the soap4 alignment hot path of HKU-BAL/MegaPath.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
// Parity pinning: the reference ships no golden vectors for this path (SURVEY.md
// section 4), so every function here is pinned against outputs of the reference
// itself, built by oracle/Makefile.ref into oracle/_ref/ (seam dumps of soap4_dump,
// libref_dp.so, libref_bwt.so) -- see tests/test_oracle_vs_ref.py and tests/golden/.
//
// All file:line citations are relative to /root/reference/soap4/.
#ifndef MP_ORACLE_H
#define MP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcIndex OrcIndex;

// MmpProperties (IniParam.h; parsed at IniParam.cpp:427-435)
typedef struct {
    int32_t seedSAsizeThreshold;   // mmpSeedSAsizeThreshold (30)
    int32_t seedMinLength;         // mmpSeedMinLength       (22 / 17 nt2)
    int32_t uniqThreshold;         // mmpUniqThreshold       (6)
    int32_t indelFuzz;             // mmpIndelFuzz           (5)
    int32_t goodSeedLen;           // mmpGoodSeedLen         (27)
    int32_t reseedLen;             // mmpReseedLen           (23 / 18 nt2)
    double  reseedRLTratio;        // mmpReseedRLTratio      (0.7)
    int32_t reseedAbsDiff;         // mmpReseedAbsDiff       (4)
    double  shortSeedRatio;        // mmpShortSeedRatio      (0.5)
} OrcMmpParams;

// SeedSAalign (DV-DPfunctions.cpp:2161-2171)
typedef struct { uint32_t query_offset; uint64_t sa_l; uint32_t sa_diff; uint32_t seed_len; } OrcSeedSA;
// SeedPos (SeedPool.h:52-57)
typedef struct { uint64_t pos; uint32_t strand_readID; uint32_t paired_seedLength; } OrcSeedPos;
// DeepDP_Space::CandidateInfo at the pairing seam (DV-DPfunctions.h:1232-1240)
typedef struct { uint32_t readIDLeft; uint32_t pad; uint64_t pos[2]; } OrcCandidate;

// ---- index (2bwt-lib/BWT.c, 2bwt-flex/LT.c, 2bwt-lib/TextConverter.c) ----
OrcIndex *orc_index_load(const char *prefix);
void      orc_index_free(OrcIndex *);
uint64_t  orc_text_length(const OrcIndex *);
uint64_t  orc_inverse_sa0(const OrcIndex *);
void      orc_cum_freq(const OrcIndex *, uint64_t out[5]);
uint64_t  orc_occ(const OrcIndex *, uint64_t idx, uint32_t c);          // BWTOccValue, BWT.c:597-634
uint64_t  orc_sa(const OrcIndex *, uint64_t saIndex);                   // BWTSaValue,  BWT.c:968-998
void      orc_lkt(const OrcIndex *, uint32_t key, uint64_t *l, uint64_t *r); // DV-DPfunctions.cpp:2240-2241
uint32_t  orc_text_base(const OrcIndex *, uint64_t pos);                // .pac symbol, TextConverter.c:427-479
void      orc_text_window(const OrcIndex *, uint64_t start, uint32_t len, uint8_t *out);
void      orc_occ_many(const OrcIndex *, int n, const uint64_t *idx, const uint32_t *c, uint64_t *out);
void      orc_sa_many(const OrcIndex *, int n, const uint64_t *idx, uint64_t *out);
// work counters accumulated by the calls below: occ, onspot-occ, sa, lkt
void      orc_counters(uint64_t out[4], int reset);

// ---- MMP seeding (DV-DPfunctions.cpp:2188-2377): read = codes 0..3, forward orientation ----
// strand 0 = mmp<0> ('+'), 1 = mmp<2> ('-').  Returns number of seeds written (<= cap).
int orc_mmp(const OrcIndex *, const uint8_t *read, int len, int strand, const OrcMmpParams *, OrcSeedSA *out, int cap);

// ---- seeding post-processing + SeedPos arrays (DV-DPfunctions.cpp:2404-2615) ----
// reads: nPairs*2 reads (mate1, mate2 interleaved: read id 2p, 2p+1), stride maxLen codes.
// Outputs the two arrays exactly as mmpSeeding builds them (with both sentinels);
// *nReadPos / *nMatePos include the two sentinels.  Caller frees with orc_free.
void orc_seed_pairs(const OrcIndex *, const uint8_t *reads, const uint32_t *lens, int maxLen, int nPairs,
                    const OrcMmpParams *, OrcSeedPos **readPos, uint64_t *nReadPos,
                    OrcSeedPos **matePos, uint64_t *nMatePos);

// ---- candidate pairing (DV-DPfunctions.cpp:1968-2119) ----
void orc_pair_candidates(OrcSeedPos *readPos, uint64_t nReadPos, OrcSeedPos *matePos, uint64_t nMatePos,
                         const uint32_t *lens, int insert_low, int insert_high,
                         OrcCandidate **cands, uint64_t *nCands);
void orc_free(void *);

// ---- semi-global DP + traceback (CPU_DP.cpp:122-619, 622-786, 788-871) ----
// ref/read: codes 0..3.  pattern must hold dnaLen+readLen+8 bytes.
// Outputs what callDP writes for one task: score (0 if below cutoff or discarded),
// hitLoc, maxScoreCount, pattern (only when score >= cutoff).
void orc_dp(const uint8_t *ref, int dnaLen, const uint8_t *read, int readLen,
            int clipLt, int clipRt, int mismatch, int gapOpen, int cutoff,
            int *score, uint32_t *hitLoc, uint32_t *count, uint8_t *pattern);

#ifdef __cplusplus
}
#endif
#endif
