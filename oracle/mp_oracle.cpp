// oracle/mp_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
//
// Index primitives, MMP seeding, seed post-processing and paired-end candidate
// generation of soap4, restated as plain scalar C++.  file:line citations are relative
// to /root/reference/soap4/.
#include "mp_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

struct OrcIndex {
    uint64_t n = 0, inverseSa0 = 0, cum[5] = {0, 0, 0, 0, 0}, saInterval = 16;
    std::vector<uint32_t> bwt;       // 16 symbols / word, first symbol in the top 2 bits (BWT.c:132-157)
    std::vector<uint32_t> occMinor;  // two u16 samples per word per symbol, even sample high (BWT.c:783-797)
    std::vector<uint64_t> occMajor;  // u64 x4 per 65536 symbols
    std::vector<uint64_t> sa;        // one value per saInterval SA indices; [0] = -1 (BWT.c:200-243)
    std::vector<uint64_t> lkt;       // 4^13 inclusive cumulative counts (LTConstruct.c:46-96)
    std::vector<uint8_t>  pac;       // 4 bases / byte, first base in top 2 bits (TextConverter.c:427-479)
};

static uint64_t g_cnt[4];

static std::vector<uint8_t> slurp(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "[oracle] cannot open %s\n", path.c_str()); exit(1); }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> b(sz);
    if (sz && fread(b.data(), 1, sz, f) != (size_t)sz) { fprintf(stderr, "[oracle] short read %s\n", path.c_str()); exit(1); }
    fclose(f);
    return b;
}

extern "C" OrcIndex *orc_index_load(const char *prefix)
{
    std::string p(prefix);
    OrcIndex *ix = new OrcIndex;
    {   // .bwt : u64 inverseSa0, u64 cumFreq[1..4], then packed words
        std::vector<uint8_t> b = slurp(p + ".bwt");
        const uint64_t *h = (const uint64_t *)b.data();
        ix->inverseSa0 = h[0];
        for (int i = 1; i <= 4; ++i) ix->cum[i] = h[i];
        ix->n = ix->cum[4];
        size_t words = (ix->n + 15) / 16;
        ix->bwt.assign(words + 8, 0);
        memcpy(ix->bwt.data(), b.data() + 40, words * 4);
        // BWTClearTrailingBwtCode (BWT.c:870-889)
        if (ix->n % 16) ix->bwt[ix->n / 16] &= ~0u << (32 - 2 * (ix->n % 16));
    }
    {   // .fmv : same 40-byte header, minor table, major table (BWT.c:159-186)
        std::vector<uint8_t> b = slurp(p + ".fmv");
        uint64_t nocc = (ix->n + 255) / 256 + 1;
        uint64_t minorWords = (nocc + 1) / 2 * 4;
        uint64_t majorWords = (nocc + 255) / 256 * 4;
        ix->occMinor.resize(minorWords);
        ix->occMajor.resize(majorWords);
        memcpy(ix->occMinor.data(), b.data() + 40, minorWords * 4);
        memcpy(ix->occMajor.data(), b.data() + 40 + minorWords * 4, majorWords * 8);
    }
    {   // .sa : 40-byte header, u64 saInterval, values
        std::vector<uint8_t> b = slurp(p + ".sa");
        ix->saInterval = *(const uint64_t *)(b.data() + 40);
        uint64_t cnt = (ix->n + ix->saInterval) / ix->saInterval;
        ix->sa.resize(cnt);
        memcpy(ix->sa.data(), b.data() + 48, cnt * 8);
        ix->sa[0] = (uint64_t)-1;
    }
    ix->pac = slurp(p + ".pac");
    FILE *lf = fopen((p + ".lkt").c_str(), "rb");
    if (lf) {   // .lkt : i32 tableSize(13), 4^13 u64
        fclose(lf);
        std::vector<uint8_t> b = slurp(p + ".lkt");
        int ts = *(const int32_t *)b.data();
        uint64_t cnt = 1ull << (2 * ts);
        ix->lkt.resize(cnt);
        memcpy(ix->lkt.data(), b.data() + 4, cnt * 8);
    } else {    // no .lkt (the 512 MiB table is not kept with the small golden fixtures): rebuild it as
                // BuildLookupTable does (2bwt-lib/LTConstruct.c:46-96): count the 13-mer starting at every
                // text position, windows running past the end padded with 'A', then inclusive prefix sums
        const uint64_t cnt = 1ull << 26;
        ix->lkt.assign(cnt, 0);
        uint64_t window = 0;
        for (uint64_t i = 0; i < ix->n + 12; ++i) {
            uint32_t c = i < ix->n ? ((ix->pac[i >> 2] >> ((3 - (i & 3)) << 1)) & 3) : 0;
            window = ((window << 2) | c) & (cnt - 1);
            if (i >= 12) ++ix->lkt[window];
        }
        for (uint64_t k = 1; k < cnt; ++k) ix->lkt[k] += ix->lkt[k - 1];
    }
    return ix;
}
extern "C" void orc_index_free(OrcIndex *ix) { delete ix; }
extern "C" uint64_t orc_text_length(const OrcIndex *ix) { return ix->n; }
extern "C" uint64_t orc_inverse_sa0(const OrcIndex *ix) { return ix->inverseSa0; }
extern "C" void orc_cum_freq(const OrcIndex *ix, uint64_t out[5]) { for (int i = 0; i < 5; ++i) out[i] = ix->cum[i]; }
extern "C" void orc_counters(uint64_t out[4], int reset)
{
    for (int i = 0; i < 4; ++i) out[i] = g_cnt[i];
    if (reset) memset(g_cnt, 0, sizeof g_cnt);
}

static inline uint32_t bwt_sym(const OrcIndex *ix, uint64_t i) { return (ix->bwt[i >> 4] >> ((15 - (i & 15)) << 1)) & 3; }

// occ over the $-less BWT positions [0, idx) -- explicit sample + scalar count
// (BWTOccValueExplicit BWT.c:783-797 + BWTDecode BWT.c:328-432, bidirectional)
static uint64_t occ_raw(const OrcIndex *ix, uint64_t idx, uint32_t c)
{
    uint64_t k = (idx + 127) / 256;                  // nearest explicit sample
    uint64_t major = ix->occMajor[(k * 256 / 65536) * 4 + c];
    uint32_t w = ix->occMinor[(k / 2) * 4 + c];
    uint64_t v = major + ((k % 2 == 0) ? (w >> 16) : (w & 0xffff));
    uint64_t at = k * 256;
    if (at <= idx) { for (uint64_t i = at; i < idx; ++i) v += bwt_sym(ix, i) == c; }
    else           { for (uint64_t i = idx; i < at; ++i) v -= bwt_sym(ix, i) == c; }
    return v;
}
extern "C" uint64_t orc_occ(const OrcIndex *ix, uint64_t idx, uint32_t c)
{
    ++g_cnt[0];
    idx -= (idx > ix->inverseSa0);                   // BWT.c:605
    return occ_raw(ix, idx, c);
}
// BWTPsiMinusValue (BWT.c:915-938) via BWTOccValueOnSpot (BWT.c:689-729)
static uint64_t psi_minus(const OrcIndex *ix, uint64_t index)
{
    if (index == ix->inverseSa0) return 0;
    ++g_cnt[1];
    uint64_t i = index + 1;
    i -= (i > ix->inverseSa0);
    uint32_t c = bwt_sym(ix, i - 1);
    return ix->cum[c] + occ_raw(ix, i, c);
}
extern "C" uint64_t orc_sa(const OrcIndex *ix, uint64_t saIndex)
{
    ++g_cnt[2];
    uint64_t skipped = 0;
    while (saIndex % ix->saInterval != 0) { ++skipped; saIndex = psi_minus(ix, saIndex); }
    return ix->sa[saIndex / ix->saInterval] + skipped;
}
extern "C" void orc_lkt(const OrcIndex *ix, uint32_t key, uint64_t *l, uint64_t *r)
{
    ++g_cnt[3];
    *l = key == 0 ? 1 : ix->lkt[key - 1] + 1;
    *r = ix->lkt[key];
}
extern "C" uint32_t orc_text_base(const OrcIndex *ix, uint64_t pos) { return (ix->pac[pos >> 2] >> ((3 - (pos & 3)) << 1)) & 3; }
extern "C" void orc_text_window(const OrcIndex *ix, uint64_t start, uint32_t len, uint8_t *out)
{ for (uint32_t i = 0; i < len; ++i) out[i] = (uint8_t)orc_text_base(ix, start + i); }
extern "C" void orc_occ_many(const OrcIndex *ix, int n, const uint64_t *idx, const uint32_t *c, uint64_t *out)
{ for (int i = 0; i < n; ++i) out[i] = orc_occ(ix, idx[i], c[i]); }
extern "C" void orc_sa_many(const OrcIndex *ix, int n, const uint64_t *idx, uint64_t *out)
{ for (int i = 0; i < n; ++i) out[i] = orc_sa(ix, idx[i]); }

// ---------------------------------------------------------------------------------
// MMP seeding: mmp<0> (DV-DPfunctions.cpp:2226-2267) and mmp<2> (:2319-2377) share one
// loop once the scan order is abstracted: scan position i walks a sequence q[0..len)
// where q is the read reversed ('+', strand 0) or complemented ('-', strand 1).
// CHECK_AND_SET_LAST :2188-2195, CHECK_AND_ADD_RANGE :2197-2219.
// ---------------------------------------------------------------------------------
extern "C" int orc_mmp(const OrcIndex *ix, const uint8_t *read, int len, int strand,
                       const OrcMmpParams *P, OrcSeedSA *out, int cap)
{
    const int K = 13;                                  // LOOKUP_SIZE (2bwt-flex/LT.h:49)
    std::vector<uint8_t> q(len);
    for (int i = 0; i < len; ++i) q[i] = strand == 0 ? read[len - 1 - i] : (uint8_t)(3 - read[i]);
    // lkp[i] = 13-mer q[i..i+12]; the first scanned symbol ends up in the LOW 2 bits (:2233-2239)
    const uint64_t n = ix->n;
    int nOut = 0;
    int i = 0, seed_len = 0;
    uint64_t l = 0, r = n, nextl = 0, nextr = 0, last_l = 0, last_r = n, last_seed_len = 0;
    auto add_range = [&](int x) {
        int diff = 0;
        if (seed_len >= P->seedMinLength) {
            if (seed_len >= P->reseedLen && last_r - last_l + 1 <= (uint64_t)P->seedSAsizeThreshold &&
                ((uint64_t)seed_len - last_seed_len <= (uint64_t)P->reseedAbsDiff ||
                 seed_len * P->reseedRLTratio < (double)last_seed_len)) {
                diff = seed_len - (int)last_seed_len;
                l = last_l; r = last_r; seed_len = (int)last_seed_len;
            }
            if (nOut < cap) {
                out[nOut].query_offset = (uint32_t)x & 0x3ff;                       // bitfield :10
                out[nOut].sa_l = l;
                out[nOut].sa_diff = (uint32_t)std::min<uint64_t>(P->seedSAsizeThreshold, r - l) & 0x3ff;
                out[nOut].seed_len = (uint32_t)seed_len & 0xfff;
            }
            ++nOut;
        }
        i -= diff;
        i -= std::min(seed_len, P->seedMinLength);
        l = 0; r = n; seed_len = 0; last_l = l; last_r = r; last_seed_len = 0;
    };
    for (i = 0; i < len; ++i) {
        if (seed_len == 0) {
            if (len - i < P->seedMinLength) break;
            uint32_t key = 0;
            for (int k = 0; k < K; ++k) key |= (uint32_t)q[i + k] << (2 * k);
            orc_lkt(ix, key, &nextl, &nextr);
            i += K - 1;
            seed_len = K - 1;
        } else {
            uint32_t c = q[i];
            nextl = ix->cum[c] + orc_occ(ix, l, c) + 1;
            nextr = ix->cum[c] + orc_occ(ix, r + 1, c);
        }
        if (nextl <= nextr) {
            if (seed_len >= P->seedMinLength && nextr - nextl < r - l) { last_r = r; last_l = l; last_seed_len = seed_len; }
            l = nextl; r = nextr; ++seed_len;
        } else {
            add_range(strand == 0 ? len - i : i - seed_len);
        }
    }
    add_range(strand == 0 ? 0 : len - seed_len);
    return nOut;
}

// ---------------------------------------------------------------------------------
// mmpSeeding post-processing (DV-DPfunctions.cpp:2474-2553) and array layout (:2555-2594)
// ---------------------------------------------------------------------------------
namespace {
struct SeedAlign { uint64_t offset; uint32_t multiplicity; int length; uint32_t query_offset; };
}

static void post_process_read(const OrcIndex *ix, const OrcMmpParams *P, const std::vector<OrcSeedSA> seeds[2],
                              uint32_t readLen, uint32_t evenReadID, std::vector<OrcSeedPos> &pos, std::vector<OrcSeedPos> &neg)
{
    std::vector<SeedAlign> sa[2];
    for (int a = 0; a < 2; ++a) {
        for (const OrcSeedSA &res : seeds[a]) {
            uint64_t l = res.sa_l, r = res.sa_l + res.sa_diff;
            if (r > l + P->seedSAsizeThreshold) r = P->seedSAsizeThreshold + l - 1;
            uint32_t off = res.query_offset, seedlen = res.seed_len;
            for (uint64_t k = l; k <= r; ++k) {
                uint64_t t = a == 0 ? (orc_sa(ix, k) - off) : (orc_sa(ix, k) - (uint64_t)(readLen - seedlen - off));
                SeedAlign s;
                s.offset = t;
                s.multiplicity = ((int)seedlen >= P->goodSeedLen || seedlen >= readLen / 2) ? 1 : (uint32_t)((r - l + 1) & 0xffff);
                s.length = (int)(int16_t)seedlen; s.query_offset = off & 0xffff;
                sa[a].push_back(s);
            }
        }
        std::stable_sort(sa[a].begin(), sa[a].end(), [](const SeedAlign &x, const SeedAlign &y) { return x.offset < y.offset; });
    }
    std::vector<OrcSeedPos> s_pos;
    uint32_t max_seed_len = 0;
    for (int a = 0; a < 2; ++a) {
        for (size_t m = 0; m < sa[a].size(); ++m) {
            bool has_unique = (int)sa[a][m].multiplicity <= P->uniqThreshold && sa[a][m].length >= P->seedMinLength;
            OrcSeedPos sp; sp.pos = sa[a][m].offset;
            std::vector<std::pair<uint32_t, uint32_t> > itv(1, std::make_pair(sa[a][m].query_offset, (uint32_t)(sa[a][m].length + sa[a][m].query_offset)));
            while (m + 1 < sa[a].size() && sa[a][m + 1].offset <= sp.pos + (uint64_t)P->indelFuzz) {
                ++m;
                has_unique |= (int)sa[a][m].multiplicity <= P->uniqThreshold && sa[a][m].length >= P->seedMinLength;
                itv.push_back(std::make_pair(sa[a][m].query_offset, (uint32_t)(sa[a][m].length + sa[a][m].query_offset)));
            }
            std::sort(itv.begin(), itv.end());
            uint32_t total = 0, cs = 0, ce = 0;
            for (size_t t = 0; t < itv.size(); ++t) {
                if (itv[t].first >= ce) { total += ce - cs; cs = itv[t].first; }
                ce = std::max(ce, itv[t].second);
            }
            total += ce - cs;
            if (max_seed_len < total) max_seed_len = total;
            if (has_unique || (int)total >= P->goodSeedLen) {
                sp.paired_seedLength = total;
                sp.strand_readID = evenReadID | ((uint32_t)(a != 0) << 31);
                s_pos.push_back(sp);
            }
        }
    }
    for (const OrcSeedPos &sp : s_pos)
        if (sp.paired_seedLength >= P->shortSeedRatio * max_seed_len)
            ((sp.strand_readID >> 31) ? neg : pos).push_back(sp);
}

extern "C" void orc_seed_pairs(const OrcIndex *ix, const uint8_t *reads, const uint32_t *lens, int maxLen, int nPairs,
                               const OrcMmpParams *P, OrcSeedPos **readPos, uint64_t *nReadPos,
                               OrcSeedPos **matePos, uint64_t *nMatePos)
{
    std::vector<OrcSeedPos> rp, rn, mp, mn;
    std::vector<OrcSeedSA> buf(4096);
    for (int mate = 0; mate < 2; ++mate) {
        for (int p = 0; p < nPairs; ++p) {
            uint32_t id = 2 * p + mate;
            const uint8_t *rd = reads + (size_t)id * maxLen;
            std::vector<OrcSeedSA> seeds[2];
            for (int a = 0; a < 2; ++a) {
                int ns = orc_mmp(ix, rd, lens[id], a, P, buf.data(), (int)buf.size());
                seeds[a].assign(buf.begin(), buf.begin() + std::min<int>(ns, buf.size()));
            }
            post_process_read(ix, P, seeds, lens[id], 2 * p, mate == 0 ? rp : mp, mate == 0 ? rn : mn);
        }
    }
    auto build = [](std::vector<OrcSeedPos> &a, std::vector<OrcSeedPos> &b, OrcSeedPos **out, uint64_t *n) {
        OrcSeedPos t; t.pos = ~0ull; t.paired_seedLength = 0xffffffff;
        *n = a.size() + b.size() + 2;
        OrcSeedPos *o = (OrcSeedPos *)malloc(*n * sizeof(OrcSeedPos));
        size_t k = 0;
        for (auto &s : a) o[k++] = s;
        t.strand_readID = 0x7fffffff; o[k++] = t;
        for (auto &s : b) o[k++] = s;
        t.strand_readID = 0xffffffff; o[k++] = t;
        *out = o;
    };
    build(rp, rn, readPos, nReadPos);
    build(mp, mn, matePos, nMatePos);
}

// ---------------------------------------------------------------------------------
// pairEndMerge (DV-DPfunctions.cpp:1968-2070), findRevStart (:1844-1871),
// mergeAndPairPairedEnd (:2088-2119)
// ---------------------------------------------------------------------------------
static uint64_t find_rev_start(const OrcSeedPos *arr, uint64_t len)
{
    if (len == 0 || !(arr[len - 1].strand_readID >> 31)) return len;
    uint64_t s = 0, e = len - 1;
    while (s < e) { uint64_t m = (s + e) / 2; if (arr[m].strand_readID >> 31) e = m; else s = m + 1; }
    return s;
}
#define ORC_MARGIN(l) (((l) > 100) ? 30 : 25)      // DP2_MARGIN, DV-DPfunctions.cpp:1760
static void pair_end_merge(std::vector<OrcCandidate> &out, OrcSeedPos *readPos, OrcSeedPos *matePos,
                           int isMatePositive, const uint32_t *lens, int insert_low, int insert_high)
{
    OrcSeedPos *readIter = readPos, *mateIter = matePos;
    while (true) {
        uint32_t mateID = mateIter->strand_readID & 0x7fffffff;
        while ((readIter->strand_readID & 0x7fffffff) < mateID) ++readIter;
        uint32_t readID = readIter->strand_readID & 0x7fffffff;
        while ((mateIter->strand_readID & 0x7fffffff) < readID) ++mateIter;
        mateID = mateIter->strand_readID & 0x7fffffff;
        if (mateID == 0x7fffffff) break;
        else if (readID < mateID) continue;
        OrcSeedPos *readStart = readIter, *mateStart = mateIter;
        while ((readIter->strand_readID & 0x7fffffff) == readID) ++readIter;
        while ((mateIter->strand_readID & 0x7fffffff) == mateID) ++mateIter;
        OrcSeedPos *readEnd = readIter, *mateEnd = mateIter;
        {   // MC_Compress(readStart, readEnd, 5) :2015-2026 -- in place, left leg only
            OrcSeedPos *w = readStart; uint64_t prev = w->pos;
            for (OrcSeedPos *p = readStart + 1; p < readEnd; ++p)
                if (prev + 5 < p->pos) { *(++w) = *p; prev = p->pos; }
            readEnd = w + 1;
        }
        int readLength = (int)lens[readID / 2 * 2 + 1 - isMatePositive];     // length of the '-' read
        int margin = ORC_MARGIN(readLength);
        int length_low = insert_low - readLength - margin; if (length_low < 0) length_low = 0;
        int length_high = insert_high - readLength + margin;
        OrcSeedPos *matePreStart = mateStart;
        for (OrcSeedPos *rp = readStart; rp < readEnd; ++rp) {
            uint64_t readLoc = rp->pos;
            for (OrcSeedPos *mi = matePreStart; mi < mateEnd; ++mi) {
                uint64_t mateLoc = mi->pos;
                if (readLoc + (uint64_t)(int64_t)length_high < mateLoc) break;
                else if (readLoc + (uint64_t)(int64_t)length_low <= mateLoc) {
                    OrcCandidate ci; ci.pos[0] = readLoc; ci.pos[1] = mateLoc; ci.pad = 0;
                    ci.readIDLeft = (rp->strand_readID & 0x7fffffff) + isMatePositive;
                    out.push_back(ci);
                    matePreStart = mi;                                   // updatePreStart is never set (:2044-2063)
                }
            }
        }
    }
}

extern "C" void orc_pair_candidates(OrcSeedPos *readPos, uint64_t nReadPos, OrcSeedPos *matePos, uint64_t nMatePos,
                                    const uint32_t *lens, int insert_low, int insert_high,
                                    OrcCandidate **cands, uint64_t *nCands)
{
    // StrandArrangement "+/-" (soap4.ini): left leg '+', right leg '-'
    OrcSeedPos *readNeg = readPos + find_rev_start(readPos, nReadPos);
    OrcSeedPos *mateNeg = matePos + find_rev_start(matePos, nMatePos);
    std::vector<OrcCandidate> v;
    pair_end_merge(v, readPos, mateNeg, 0, lens, insert_low, insert_high);
    pair_end_merge(v, matePos, readNeg, 1, lens, insert_low, insert_high);
    std::stable_sort(v.begin(), v.end(), [](const OrcCandidate &a, const OrcCandidate &b) { return a.readIDLeft < b.readIDLeft; });
    *nCands = v.size();
    *cands = (OrcCandidate *)malloc((v.size() + 1) * sizeof(OrcCandidate));
    if (!v.empty()) memcpy(*cands, v.data(), v.size() * sizeof(OrcCandidate));
}
extern "C" void orc_free(void *p) { free(p); }
