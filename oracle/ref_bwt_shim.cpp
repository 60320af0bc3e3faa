// oracle/ref_bwt_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the reference's own FM-index primitives so tests can compare
// the restatement / CUDA index against them directly:
//   BWTLoad            soap4/2bwt-lib/BWT.c:100-250
//   BWTOccValue        soap4/2bwt-lib/BWT.c:597-634
//   BWTOccValueOnSpot  soap4/2bwt-lib/BWT.c:689-729
//   BWTSaValue         soap4/2bwt-lib/BWT.c:968-998
//   LTLoad             soap4/2bwt-flex/LT.c:34-57
// Linked against the reference objects compiled from where they lie (Makefile.ref).
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "2bwt-lib/BWT.h"
#include "2bwt-lib/MemManager.h"
#include "2bwt-flex/LT.h"

struct RefIndex { BWT *bwt; LT *lt; MMPool *pool; };

extern "C" void *ref_index_load(const char *prefix)
{
    std::string p(prefix);
    MMMasterInitialize(3, 0, FALSE, NULL);
    RefIndex *ri = new RefIndex;
    ri->pool = MMPoolCreate(2097152);
    ri->bwt = BWTLoad(ri->pool, 0, (p + ".bwt").c_str(), (p + ".fmv").c_str(), (p + ".sa").c_str());
    ri->lt = LTLoad((p + ".lkt").c_str());
    return ri;
}
extern "C" uint64_t ref_text_length(void *h) { return ((RefIndex *)h)->bwt->textLength; }
extern "C" uint64_t ref_inverse_sa0(void *h) { return ((RefIndex *)h)->bwt->inverseSa0; }
extern "C" void ref_cum_freq(void *h, uint64_t *out) { for (int i = 0; i < 5; ++i) out[i] = ((RefIndex *)h)->bwt->cumulativeFreq[i]; }
extern "C" void ref_occ(void *h, int n, const uint64_t *idx, const uint32_t *c, uint64_t *out)
{ for (int i = 0; i < n; ++i) out[i] = BWTOccValue(((RefIndex *)h)->bwt, idx[i], c[i]); }
extern "C" void ref_sa(void *h, int n, const uint64_t *idx, uint64_t *out)
{ for (int i = 0; i < n; ++i) out[i] = BWTSaValue(((RefIndex *)h)->bwt, idx[i]); }
extern "C" void ref_lkt(void *h, int n, const uint32_t *key, uint64_t *l, uint64_t *r)
{
    LT *lt = ((RefIndex *)h)->lt;
    for (int i = 0; i < n; ++i) { l[i] = key[i] == 0 ? 1 : lt->table[key[i] - 1] + 1; r[i] = lt->table[key[i]]; }
}
