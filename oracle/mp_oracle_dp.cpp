// oracle/mp_oracle_dp.cpp -- TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
//
// Semi-global affine-gap DP + traceback, restated as an ordinary un-saturated DP.
// Follows soap4/CPU_DP.cpp:
//   GenerateDPTable      :122-619   (8-bit saturating delta encoding of the same recurrence)
//   GPUBacktrack         :622-786   (traceback state machine, reproduced literally)
//   SemiGlobalAlignment  :788-871   (cutoff test, discard rule)
// Equivalence of the plain recurrence with the delta/"score trimming" encoding is
// checked in tests against oracle/_ref/libref_dp.so (the reference's own callDP).
#include "mp_oracle.h"
#include <vector>
#include <algorithm>
#include <cstring>

namespace {
const int NEG = -100000;
}

extern "C" void orc_dp(const uint8_t *ref, int N, const uint8_t *read, int L,
                       int clipLt, int clipRt, int mm, int open, int cutoff,
                       int *scoreOut, uint32_t *hitLocOut, uint32_t *countOut, uint8_t *pattern)
{
    const int ext = -1, match = 1;
    *scoreOut = 0; *hitLocOut = 0; *countOut = 0;
    // CPU_DP.cpp:296-324 input validation: the reference aborts the whole SIMD group and
    // leaves garbage; such tasks are outside defined behaviour -> score 0.
    if (cutoff > L || cutoff <= 0 || L >= 255 + open - 1 + cutoff) return;
    const int W = L + 1;
    std::vector<int> H((size_t)(N + 1) * W), D((size_t)(N + 1) * W, NEG);
    std::vector<uint8_t> clipped((size_t)(N + 1) * W, 0);
#define AT(M, i, j) M[(size_t)(i) * W + (j)]
    // row 0 (CPU_DP.cpp:397-429): free until column clipLt, then one gap
    AT(H, 0, 0) = 0; AT(clipped, 0, 0) = 1;
    for (int j = 1; j <= L; ++j) {
        if (j <= clipLt) { AT(H, 0, j) = 0; AT(clipped, 0, j) = 1; }
        else if (j == clipLt + 1) AT(H, 0, j) = AT(H, 0, j - 1) + open;
        else AT(H, 0, j) = AT(H, 0, j - 1) + ext;
    }
    int best = 0, bestRow = 0, bestCol = 0; unsigned cnt = 0;
    const int minCol = std::max(L - clipRt, 1);
    for (int i = 1; i <= N; ++i) {
        AT(H, i, 0) = 0; AT(clipped, i, 0) = 1;     // column 0 is stored as all-zero cells (:447-450)
        int I = NEG;
        for (int j = 1; j <= L; ++j) {
            int s = (ref[i - 1] == read[j - 1]) ? match : mm;
            int d = std::max(AT(D, i - 1, j) + ext, AT(H, i - 1, j) + open);
            I = std::max(I + ext, AT(H, i, j - 1) + open);
            int h = std::max(AT(H, i - 1, j - 1) + s, std::max(d, I));
            if (j <= clipLt && h < 0) { h = 0; AT(clipped, i, j) = 1; }   // :505-510
            AT(H, i, j) = h; AT(D, i, j) = d;
            if (j >= minCol && h >= cutoff) {                             // :545-590
                if (h > best) { best = h; bestRow = i; bestCol = j; cnt = 1; }
                else if (h == best) { if (cnt < 255) ++cnt; }
            }
        }
    }
    if (best < cutoff) return;
    *scoreOut = best; *countOut = cnt;
    // ---- GPUBacktrack (:622-786) ----
    unsigned p = 0;
    int clipR = L - bestCol;
    if (clipR > 0) { pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)clipR; }
    int i = L - clipR;   // read position
    int j = bestRow;     // reference row
    enum { NORMAL, I_EXT, D_EXT, SM_EXIT, SI_EXIT, SD_EXIT } state = NORMAL;
    int8_t accum = 0;
    while (i > 0 && j > 0) {
        int8_t hd = (int8_t)(AT(H, j, i) - AT(H, j, i - 1));
        int8_t dd = (int8_t)(AT(H, j, i) - AT(H, j - 1, i - 1));
        int8_t vd = (int8_t)(AT(H, j, i) - AT(H, j - 1, i));
        int flag = AT(clipped, j, i) ? 0 : (AT(D, j, i) == AT(H, j, i) ? 1 : 2);
        bool eq = ref[j - 1] == read[i - 1];
        int8_t ms = eq ? match : mm;
        if (state == NORMAL) {
            if (AT(clipped, j - 1, i - 1) && dd == ms && i != 1) { state = SM_EXIT; break; }
            else if (dd == ms) { pattern[p++] = eq ? 'M' : 'm'; --j; --i; }
            else if (flag == 1) {
                pattern[p++] = 'D'; --j;
                if (vd != open) { accum = vd - ext; state = D_EXT; }
            } else {
                pattern[p++] = 'I'; --i;
                if (hd != open) { accum = hd - ext; state = I_EXT; }
            }
        } else if (state == D_EXT) {
            if (AT(clipped, j - 1, i) && vd + accum == open) { state = SD_EXIT; break; }
            pattern[p++] = 'D'; --j;
            if (vd + accum == open) state = NORMAL; else accum += vd - ext;
        } else {
            if (AT(clipped, j, i - 1) && hd + accum == open) { state = SI_EXIT; break; }
            pattern[p++] = 'I'; --i;
            if (hd + accum == open) state = NORMAL; else accum += hd - ext;
        }
    }
    if (j == 0) {
        int sc = std::min(clipLt & 0xff, i);     // clipLtCheckLoc is a uint8_t parameter (:626)
        if (sc < i) { pattern[p++] = 'I'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)(i - sc); }
        pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)sc;
    } else if (state == SI_EXIT) {
        pattern[p++] = 'I'; pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)(i - 1);
    } else if (state == SD_EXIT) {
        pattern[p++] = 'D'; pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)(i - 1);
        pattern[p++] = 0;
        *scoreOut = 0; *hitLocOut = 0;            // :842-857 alignment discarded, count kept
        return;
    } else if (state == SM_EXIT) {
        pattern[p++] = (ref[j - 1] == read[i - 1]) ? 'M' : 'm';
        pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)(i - 1);
        j -= 1;
    }
    pattern[p++] = 0;
    *hitLocOut = (uint32_t)j;
#undef AT
}
