// fasta_index.h -- FASTA front end of the index builder: the text 2bwt-builder indexes and its .ann / .amb / .tra files.
//
// Restates what HSPParseFASTAToPacked observably does (2bwt-lib/HSP.c:354-699), because every output position of soap4 goes
// through the translate table it writes (getChrAndPos, BGS-IO.cpp:163-190):
//   * sequence name = header up to the first white space (at most 256 chars); gi number when the name starts with "gi|"
//   * bases: A C G T in either case; any other IUPAC code (M R S V W Y H K D B N; every lower-case letter too with -U) starts an
//     AMBIGUITY RUN that lasts until the next A/C/G/T, not counting characters that are no nucleotide code at all (line ends, digits,
//     'L', ...): fewer than 10 codes -> that many 'G's; 10 or more -> the run is cut out of the text and remembered
//     (start in the text, cumulative number of bases cut so far)                                               HSP.c:486-528
//   * the translate table interleaves sequence starts and cut-out runs by text position (a run that starts exactly at a sequence
//     start swallows that sequence's own entry); its `correction` arithmetic is the reference's, unsigned wrap-around included
//                                                                                                           HSP.c:569-640
// The reference reads out of bounds (ambiguity[-1]) when a run is cut out of the third or a later sequence before any run was cut
// out of the sequences up to two before it; such inputs have no defined output there and are refused here.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <ctype.h>
#include <string>
#include <vector>

struct FastaText {
    std::vector<uint8_t> pac;                     // 4 bases per byte, first base in the top 2 bits (no trailer)
    uint64_t n = 0;                               // bases kept
    struct Seq { uint32_t gi; std::string name; uint64_t start, end; int firstAmbiguityIndex; uint64_t actualStart, actualEnd; };
    struct Amb { uint64_t startPos, rightOfEndPos; };
    std::vector<Seq> seqs;
    std::vector<Amb> ambs;
    std::string error;
};

// -> false + text.error on I/O or unsupported input
static inline bool fasta_parse(const char *path, bool maskLowerCase, FastaText &t)
{
    static const char dnaChar[16] = { 'A', 'C', 'G', 'T', 'M', 'R', 'S', 'V', 'W', 'Y', 'H', 'K', 'D', 'B', 'N', 'L' };
    static const char ambiguityCount[16] = { 1, 1, 1, 1, 2, 2, 2, 3, 2, 2, 3, 2, 3, 3, 4, 0 };
    uint8_t charMap[256];
    memset(charMap, 15, sizeof charMap);          // 15 = not a nucleotide code
    for (int i = 0; i < 16; ++i) { charMap[(uint8_t)dnaChar[i]] = (uint8_t)i; charMap[(uint8_t)(dnaChar[i] - 'A' + 'a')] = (uint8_t)i; }
    FILE *f = fopen(path, "rb");
    if (!f) { t.error = std::string("cannot open ") + path; return false; }
    std::vector<char> data;
    { char buf[1 << 16]; size_t k; while ((k = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + k); }
    fclose(f);
    const size_t N = data.size();
    if (N == 0 || data[0] != '>') { t.error = "FASTA file does not begin with '>'"; return false; }
    size_t p = 1;
    uint64_t totalNumChar = 0, totalDiscarded = 0;
    uint8_t cur = 0; int nInByte = 0;
    auto push = [&](unsigned code) { cur = (uint8_t)((cur << 2) | code); if (++nInByte == 4) { t.pac.push_back(cur); cur = 0; nInByte = 0; } };
    while (p <= N) {                              // one sequence per trip; `p` sits behind a '>'
        FastaText::Seq s; s.gi = 0;
        bool comment = false;
        while (p < N && data[p] != '\n') {
            const char c = data[p++];
            if (!comment && isspace((unsigned char)c)) comment = true;
            if (!comment && s.name.size() < 256) s.name.push_back(c);
        }
        if (p < N) ++p;                           // the line end
        if (s.name.size() > 3 && s.name[0] == 'g' && s.name[1] == 'i' && s.name[2] == '|') { unsigned g = 0; sscanf(s.name.c_str() + 3, "%u", &g); s.gi = g; }
        s.start = totalNumChar; s.actualStart = totalNumChar + totalDiscarded;
        uint64_t numChar = 0;
        while (p < N && data[p] != '>') {
            char c = data[p];
            if (c == '\n' || c == '\t') { ++p; continue; }
            if (maskLowerCase && c >= 'a' && c <= 'z') c = 'N';
            const uint8_t m = charMap[(uint8_t)c];
            if (m == 15) { ++p; continue; }
            if (ambiguityCount[m] == 1) { push(m); ++numChar; ++p; continue; }
            // ambiguity run: count the codes up to the next A/C/G/T (or the next header)
            uint64_t nCount = 1;
            ++p;
            while (p < N && data[p] != '>') {
                const uint8_t m2 = charMap[(uint8_t)data[p]];
                if (m2 != 15) { if (ambiguityCount[m2] != 1) ++nCount; else break; }
                ++p;
            }
            if (nCount < 10) { for (uint64_t k = 0; k < nCount; ++k) { push(2); ++numChar; } }
            else {
                totalDiscarded += nCount;
                FastaText::Amb a; a.startPos = totalNumChar + numChar; a.rightOfEndPos = totalNumChar + numChar + totalDiscarded - 1;
                t.ambs.push_back(a);
            }
        }
        s.end = totalNumChar + numChar - 1;
        s.firstAmbiguityIndex = (int)t.ambs.size();
        s.actualEnd = totalNumChar + totalDiscarded + numChar - 1;
        totalNumChar += numChar;
        t.seqs.push_back(s);
        if (p >= N) break;
        ++p;                                      // behind the next '>'
    }
    if (nInByte) t.pac.push_back((uint8_t)(cur << (2 * (4 - nInByte))));
    t.n = totalNumChar;
    return true;
}

// .ann / .amb / .tra exactly as the reference writes them (HSP.c:569-699)
static inline bool fasta_write_annotation(const FastaText &t, const std::string &prefix, std::string &error)
{
    const size_t numSeq = t.seqs.size(), numAmb = t.ambs.size();
    struct Tr { unsigned long long startPos; unsigned chrID; unsigned long long correction; };
    std::vector<Tr> tr(numSeq + numAmb);
    const uint64_t GRID = 262144;
    const unsigned gridEntries = (unsigned)(t.n / GRID) + 1;
    std::vector<unsigned> grid(gridEntries, 0);
    size_t i = 0, j = 0, k = 0;
    auto amb_entry = [&]() -> bool {
        grid[t.ambs[j].startPos / GRID] += 1;
        tr[i].startPos = t.ambs[j].startPos; tr[i].chrID = (unsigned)k;
        const unsigned long long span = t.ambs[j].rightOfEndPos - t.ambs[j].startPos + 1;
        if ((long long)k - 1 > 0) {
            const int idx = t.seqs[k - 2].firstAmbiguityIndex - 1;
            if (idx < 0) { error = "an ambiguity run is cut out of sequence " + std::to_string(k) + " before any run in the sequences up to " + std::to_string(k - 1) +
                                   ": the reference reads out of bounds for this input (HSP.c:583-585), its output is undefined"; return false; }
            const unsigned long long corr = t.ambs[idx].rightOfEndPos - t.ambs[idx].startPos + 1;
            tr[i].correction = t.seqs[k - 1].start + corr - span - 1;
        } else tr[i].correction = 0ull - t.ambs[j].rightOfEndPos + t.ambs[j].startPos - 2;
        ++j; ++i;
        return true;
    };
    auto seq_entry = [&]() {
        grid[t.seqs[k].start / GRID] += 1;
        tr[i].startPos = t.seqs[k].start; tr[i].chrID = (unsigned)k + 1; tr[i].correction = t.seqs[k].start - 1;
        ++k; ++i;
    };
    while (j < numAmb && k < numSeq) {
        if (t.ambs[j].startPos < t.seqs[k].start) { if (!amb_entry()) return false; }
        else if (t.ambs[j].startPos > t.seqs[k].start) seq_entry();
        else ++k;                                  // a run that starts with the sequence swallows the sequence's entry
    }
    while (j < numAmb) if (!amb_entry()) return false;
    while (k < numSeq) seq_entry();
    for (; i < numAmb + numSeq; ++i) if (i > 0) tr[i] = tr[i - 1];
    for (unsigned g = 1; g < gridEntries; ++g) grid[g] += grid[g - 1];
    for (unsigned g = 0; g < gridEntries; ++g) grid[g]--;
    FILE *f = fopen((prefix + ".ann").c_str(), "w");
    if (!f) { error = "cannot create " + prefix + ".ann"; return false; }
    fprintf(f, "%llu %u %u\n", (unsigned long long)t.n, (unsigned)numSeq, 0u);
    for (const FastaText::Seq &s : t.seqs) {
        fprintf(f, "%u %s\n", s.gi, s.name.c_str());
        fprintf(f, "%llu %llu 0\n", (unsigned long long)s.start, (unsigned long long)(s.end - s.start + 1));
    }
    fclose(f);
    f = fopen((prefix + ".amb").c_str(), "w");
    if (!f) { error = "cannot create " + prefix + ".amb"; return false; }
    fprintf(f, "%llu %u %u\n", (unsigned long long)t.n, (unsigned)numSeq, 0u);
    fclose(f);
    f = fopen((prefix + ".tra").c_str(), "w");
    if (!f) { error = "cannot create " + prefix + ".tra"; return false; }
    fprintf(f, "%llu %u %u %u\n", (unsigned long long)t.n, (unsigned)numSeq, (unsigned)numAmb, gridEntries);
    for (unsigned g = 0; g < gridEntries; ++g) fprintf(f, "%u\n", grid[g]);
    for (const Tr &x : tr) fprintf(f, "%llu %u %llu\n", x.startPos, x.chrID, x.correction);
    for (const FastaText::Seq &s : t.seqs) fprintf(f, "%llu %llu\n", (unsigned long long)s.actualStart, (unsigned long long)(s.actualEnd - s.actualStart + 1));
    fclose(f);
    return true;
}
