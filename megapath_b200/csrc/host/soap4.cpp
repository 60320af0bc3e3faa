// soap4.cpp -- host driver with the reference soap4's command line, .ini semantics and MegaPath
// output (annotated interleaved FASTQ on stdout), calling the B200 hot path through the C-ABI of
// libmegapath_b200.so (include/megapath_b200.h).  Drop-in for the `soap4 pair ...` calls of
// runMegaPath.sh:136,199.
//
// Mirrors (reference file:line, relative to soap4/):
//   command line             parseInputArgs                     IniParam.cpp:542-942
//   .ini                     ParseIniFile                       IniParam.cpp:242-438
//   read loading             loadPairReadsKseq / appendToQueryArrays   QueryParser.cpp:160-260, kseq.h
//   batch loop, first-batch read-length detection / insert_low clamp    SOAP4.cpp:424-585
//   stage sequencing         soap3_dp_pair_align                alignment.cpp:29-355
//   per-pair best / FASTQ    outputDeepDPResult2, pairDeepDPOutputFastqAPI, unproperlypairDPOutputFastqAPI,
//                            decideTargetChr, getChrAndPos, getMappingFromHeader
//                            OutputDPResult.cpp:65-265; BGS-IO.cpp:163-190, 1312-1446, 1966-2091
//   unpaired bookkeeping     filterOutUnpairedSingleReads, DPSOutputUnpairedAlignment
//                            SeedPool.cpp:267-322; DV-DPfunctions.cpp:841-920
// BAM output (-b) is not implemented yet; the flag is accepted and reported on stderr.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <zlib.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "megapath_b200.h"

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------------
struct Options {
    std::string indexName, query1, query2, outputPrefix, iniFile;
    int maxReadLength = 120, insert_low = 1, insert_high = 500, outputBAM = 0, numCpuThreads = 4, device = 0;
    int megapathMode = 0, top = 95, ignoreComments = 0, alignmentType = 2, printMDNM = 0;
};

static bool parse_args(int argc, char **argv, Options &o)
{
    if (argc < 2 || strcmp(argv[1], "pair") != 0) {
        fprintf(stderr, "Usage: %s pair <index prefix> <reads_1.fq[.gz]> <reads_2.fq[.gz]> [-o prefix] [-C ini] [-L len] [-T n] [-u n] [-v n] [-b] [-F|-P] [-nc] [-top n] [-c gpu]\n", argv[0]);
        if (argc >= 2) fprintf(stderr, "Only 'pair' mode is supported (the reference asserts readType == PAIR_END_READ, SOAP4.cpp:554).\n");
        return false;
    }
    if (argc < 5) { fprintf(stderr, "Invalid number of command-line arguments.\n"); return false; }
    o.indexName = argv[2]; o.query1 = argv[3]; o.query2 = argv[4]; o.outputPrefix = argv[3];
    for (int i = 5; i < argc; i++) {
        const char *a = argv[i];
        auto need = [&](const char *what) { if (i + 1 >= argc) { fprintf(stderr, "Please specify %s after '%s'\n", what, a); return false; } return true; };
        if (!strcmp(a, "-h")) { if (!need("the output option")) return false; int t = atoi(argv[++i]); if (t < 1 || t > 4) { fprintf(stderr, "The output option should be 1, 2, 3 or 4\n"); return false; } o.alignmentType = t; }
        else if (!strcmp(a, "-l") || !strcmp(a, "-L")) { if (!need("the length")) return false; o.maxReadLength = atoi(argv[++i]);
            if (o.maxReadLength < 0) { fprintf(stderr, "The length should not be less than 0\n"); return false; }
            if (o.maxReadLength > 1024) { fprintf(stderr, "The length should not be greater than %u\n", 1024u); return false; } }
        else if (!strcmp(a, "-u")) { if (!need("the maximum value of insert size")) return false; o.insert_high = atoi(argv[++i]); }
        else if (!strcmp(a, "-v")) { if (!need("the minimum value of insert size")) return false; o.insert_low = atoi(argv[++i]); }
        else if (!strcmp(a, "-b")) o.outputBAM = 1;
        else if (!strcmp(a, "-o")) { if (!need("the output file prefix")) return false; o.outputPrefix = argv[++i]; }
        else if (!strcmp(a, "-c")) { if (!need("the GPU device ID")) return false; o.device = atoi(argv[++i]); if (o.device < 0) { fprintf(stderr, "The GPU device ID should not be less than 0\n"); return false; } }
        else if (!strcmp(a, "-p")) o.printMDNM = 1;
        else if (!strcmp(a, "-T")) { if (!need("the number of CPU threads")) return false; o.numCpuThreads = atoi(argv[++i]); if (o.numCpuThreads <= 0) { fprintf(stderr, "Please specify a positive number of CPU threads after '-T'\n"); return false; } }
        else if (!strcmp(a, "-C")) { if (!need("ini file name")) return false; o.iniFile = argv[++i]; }
        else if (!strcmp(a, "-F")) o.megapathMode = 1;
        else if (!strcmp(a, "-P")) o.megapathMode = 2;
        else if (!strcmp(a, "-top")) { if (!need("the value of '-top'")) return false; o.top = atoi(argv[i + 1]); }   // the reference does not consume the value (IniParam.cpp:885-892)
        else if (!strcmp(a, "-nc")) o.ignoreComments = 1;
        else if (!strcmp(a, "-D") || !strcmp(a, "-A") || !strcmp(a, "-R") || !strcmp(a, "-e")) { if (i + 1 < argc) ++i; }
        else if (!strcmp(a, "-I")) { fprintf(stderr, "illumina quality is not supported yet\n"); return false; }
    }
    if (o.insert_low > o.insert_high) { fprintf(stderr, "The minimum value of insert size should not be greater than the maximum value of insert size.\n"); return false; }
    return true;
}

// ------------------------------------------------------------------------------------------------
// .ini (iniparser semantics: "section:key", case-insensitive keys, ';' / '#' comments)
struct Ini {
    std::map<std::string, std::string> kv;
    static std::string lower(std::string s) { for (auto &c : s) c = (char)tolower(c); return s; }
    static std::string trim(const std::string &s) { size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n"); return a == std::string::npos ? "" : s.substr(a, b - a + 1); }
    bool load(const std::string &path) {
        FILE *f = fopen(path.c_str(), "r"); if (!f) return false;
        char line[4096]; std::string sec;
        while (fgets(line, sizeof line, f)) {
            std::string s = trim(line);
            if (s.empty() || s[0] == ';' || s[0] == '#') continue;
            if (s[0] == '[') { size_t e = s.find(']'); if (e != std::string::npos) sec = lower(trim(s.substr(1, e - 1))); continue; }
            size_t eq = s.find('='); if (eq == std::string::npos) continue;
            std::string k = lower(trim(s.substr(0, eq))), v = trim(s.substr(eq + 1));
            size_t c = v.find_first_of(";#"); if (c != std::string::npos) v = trim(v.substr(0, c));
            if (v.size() >= 2 && (v[0] == '"' || v[0] == '\'') && v.back() == v[0]) v = v.substr(1, v.size() - 2);
            kv[sec + ":" + k] = v;
        }
        fclose(f); return true;
    }
    int geti(const char *k, int d) const { auto it = kv.find(lower(k)); return it == kv.end() ? d : (int)strtol(it->second.c_str(), nullptr, 0); }
    double getd(const char *k, double d) const { auto it = kv.find(lower(k)); return it == kv.end() ? d : atof(it->second.c_str()); }
    std::string gets(const char *k, const char *d) const { auto it = kv.find(lower(k)); return it == kv.end() ? d : it->second; }
};

// ------------------------------------------------------------------------------------------------
// chromosome translation (.ann names, .tra grid + translate table; HSP.c:57-330, BGS-IO.cpp:163-190)
struct Annotation {
    uint64_t dnaLength = 0; uint32_t numSeq = 0;
    std::vector<std::string> names;
    std::vector<uint32_t> grid;
    struct Tr { uint64_t startPos; uint32_t chrID; uint64_t correction; };
    std::vector<Tr> tr;
    bool load(const std::string &prefix) {
        FILE *f = fopen((prefix + ".ann").c_str(), "r"); if (!f) { fprintf(stderr, "Cannot open annotation file!\n"); return false; }
        unsigned long long n; unsigned ns, seed;
        if (fscanf(f, "%llu %u %u\n", &n, &ns, &seed) != 3) { fclose(f); return false; }
        dnaLength = n; numSeq = ns;
        char buf[4096];
        for (uint32_t i = 0; i < ns; ++i) {
            unsigned gi; if (fscanf(f, "%u ", &gi) != 1) break;
            if (!fgets(buf, sizeof buf, f)) break;
            size_t l = strlen(buf); if (l && buf[l - 1] == '\n') buf[l - 1] = 0;
            names.push_back(buf);
            unsigned long long a, b; int c; if (fscanf(f, "%llu %llu %d\n", &a, &b, &c) != 3) break;
        }
        fclose(f);
        if (names.size() != ns) { fprintf(stderr, "Annotation missing entries!\n"); return false; }
        f = fopen((prefix + ".tra").c_str(), "r"); if (!f) { fprintf(stderr, "Cannot open translate file!\n"); return false; }
        unsigned long long n2; int ns2; unsigned removed, gridEntries;
        if (fscanf(f, "%llu %d %u %u\n", &n2, &ns2, &removed, &gridEntries) != 4) { fclose(f); return false; }
        grid.resize(gridEntries);
        for (unsigned j = 0; j < gridEntries; ++j) if (fscanf(f, "%u\n", &grid[j]) != 1) { fclose(f); return false; }
        tr.resize(ns + removed);
        for (size_t j = 0; j < tr.size(); ++j) {
            unsigned long long s, c; unsigned id;
            if (fscanf(f, "%llu %u %llu\n", &s, &id, &c) != 3) { fclose(f); fprintf(stderr, "Translate missing entries!\n"); return false; }
            tr[j].startPos = s; tr[j].chrID = id; tr[j].correction = c;
        }
        fclose(f);
        return true;
    }
    void chrAndPos(uint64_t ambPos, uint64_t *tp, uint32_t *chr) const {
        uint64_t idx = ambPos >> 18;
        if (idx >= grid.size()) idx = grid.size() - 1;
        uint32_t v = grid[idx];
        while (tr[v].startPos > ambPos) v--;
        *tp = ambPos - tr[v].correction; *chr = tr[v].chrID;
    }
    // decideTargetChr (BGS-IO.cpp:1312-1341): -1 when the read window [pos, pos+readLen) crosses sequences
    int targetChr(uint64_t ambPos, uint32_t readLen) const {
        uint64_t p; uint32_t c0, c1;
        chrAndPos(ambPos, &p, &c0); chrAndPos(ambPos + readLen - 1, &p, &c1);
        return c0 == c1 ? (int)c0 : -1;
    }
};

// ------------------------------------------------------------------------------------------------
// FASTA/FASTQ reader with kseq semantics (name up to the first blank, comment = rest of the header)
struct SeqReader {
    gzFile f = nullptr; std::vector<char> buf; size_t pos = 0, end = 0; bool eof = false; int last = 0;
    bool open(const std::string &path) { f = gzopen(path.c_str(), "rb"); if (!f) return false; gzbuffer(f, 1 << 20); buf.resize(1 << 20); return true; }
    int getc_() { if (pos >= end) { if (eof) return -1; int n = gzread(f, buf.data(), (unsigned)buf.size()); if (n <= 0) { eof = true; return -1; } pos = 0; end = (size_t)n; } return (unsigned char)buf[pos++]; }
    // reads up to the delimiter class: 0 = blank (space/tab/newline), 2 = newline; returns delimiter or -1
    int getuntil(int mode, std::string &s, bool append) {
        if (!append) s.clear();
        for (;;) { int c = getc_(); if (c < 0) return -1; if (mode == 2 ? c == '\n' : isspace(c)) { if (mode == 2 && !s.empty() && s.back() == '\r') s.pop_back(); return c; } s.push_back((char)c); }
    }
    // -> length of the sequence, -1 at end of file
    int read(std::string &name, std::string &comment, std::string &seq, std::string &qual) {
        int c;
        if (last == 0) { while ((c = getc_()) >= 0 && c != '>' && c != '@') {} if (c < 0) return -1; last = c; }
        comment.clear(); seq.clear(); qual.clear();
        if ((c = getuntil(0, name, false)) < 0) return -1;
        if (c != '\n') getuntil(2, comment, false);
        while ((c = getc_()) >= 0 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            seq.push_back((char)c);
            getuntil(2, seq, true);
        }
        if (c == '>' || c == '@') last = c;
        if (c != '+') return (int)seq.size();
        std::string skip;
        getuntil(2, skip, false);                                 // rest of the '+' line
        while (qual.size() < seq.size()) { if (getuntil(2, qual, true) < 0) break; }
        last = 0;
        if (seq.size() != qual.size()) return -2;
        return (int)seq.size();
    }
};

struct ReadBatch {
    uint32_t nReads = 0, wpq = 0, maxReadLength = 0;
    std::vector<uint32_t> queries, lens;
    std::vector<std::string> names, comments, quals;
    std::vector<uint8_t> hasComment;
};

static unsigned char g_charMap[256];
static void fill_char_map() {                                    // INDEXFillCharMap (IndexHandler.cpp:26-45)
    memset(g_charMap, 0, sizeof g_charMap);
    const char *dna = "ACGT";
    for (int i = 0; i < 4; ++i) { g_charMap[(int)dna[i]] = (unsigned char)i; g_charMap[dna[i] - 'A' + 'a'] = (unsigned char)i; }
    g_charMap['U'] = g_charMap['u'] = 3; g_charMap['N'] = g_charMap['n'] = 2;
}

static void append_read(ReadBatch &b, uint32_t id, std::string &name, const std::string &comment, const std::string &seq, const std::string &qual)
{
    uint32_t len = seq.size() > b.maxReadLength - 1 ? b.maxReadLength - 1 : (uint32_t)seq.size();     // QueryParser.cpp:188
    b.lens[id] = len;
    uint32_t *q = b.queries.data() + ((size_t)(id / 32) * 32 * b.wpq + id % 32);
    uint32_t word = 0; int off = 0;
    for (uint32_t i = 0; i < len; ++i) {
        word |= (uint32_t)g_charMap[(unsigned char)seq[i]] << (off * 2);
        if (++off == 16) { *q = word; q += 32; off = 0; word = 0; }
    }
    if (off > 0) *q = word;
    if (name.size() > 2 && name[name.size() - 2] == '/' && isdigit((unsigned char)name.back())) name.resize(name.size() - 2);   // trim_readno
    b.names[id] = name;
    b.hasComment[id] = !comment.empty(); b.comments[id] = comment;
    b.quals[id] = qual.substr(0, len);
}

static uint32_t load_batch(SeqReader &r1, SeqReader &r2, ReadBatch &b, uint32_t maxReads)
{
    size_t words = ((size_t)maxReads + 31) / 32 * 32 * b.wpq;
    b.queries.assign(words, 0); b.lens.assign(maxReads, 0);
    b.names.assign(maxReads, ""); b.comments.assign(maxReads, ""); b.quals.assign(maxReads, ""); b.hasComment.assign(maxReads, 0);
    uint32_t n = 0;
    std::string n1, c1, s1, q1, n2, c2, s2, q2;
    while (n < maxReads) {
        int l1 = r1.read(n1, c1, s1, q1), l2 = r2.read(n2, c2, s2, q2);
        if ((l1 >= 0 && l2 < 0) || (l1 < 0 && l2 >= 0)) { fprintf(stderr, "Error: number of sequences of pair-end files not matched.\n"); exit(1); }
        if (l1 < 0) break;
        append_read(b, n++, n1, c1, s1, q1);
        append_read(b, n++, n2, c2, s2, q2);
    }
    b.nReads = n;
    return n;
}

static uint32_t detect_read_length(const std::vector<uint32_t> &lens, uint32_t numQueries, uint32_t start)   // GetReadLength (QueryParser.cpp:2253-2277)
{
    if (numQueries == 0) return 100;
    uint32_t i = 0, j = 1, mx = lens[start];
    while (i < numQueries && j < 1000000) { if (mx < lens[start + i]) mx = lens[start + i]; j++; i += 2; }
    return mx;
}

// ------------------------------------------------------------------------------------------------
// FASTQ header composition
struct HeaderHit { int score; size_t s, e; };
static int mapping_from_header(const std::string *comment, std::vector<HeaderHit> &v, double top, double scoreT)   // BGS-IO.cpp:1348-1371
{
    v.clear();
    if (!comment || *comment == "IGNORE") return 0;
    const char *c = comment->c_str();
    if (comment->size() < 6) return 0;
    int score = atoi(c + 6);
    if (score < scoreT) return score;
    else if (scoreT < score * top) scoreT = score * top;
    const char *p = strchr(c + 6, ';');
    while (p && *(p + 1) != '\0') {
        HeaderHit m; m.score = atoi(p + 1); m.s = (size_t)(p + 1 - c);
        p = strchr(p + 1, ';');
        if (!p) break;
        m.e = (size_t)(p - c);
        if (score >= scoreT) v.push_back(m);
    }
    return score;
}

static void seq_and_qual(std::string &out, const ReadBatch &b, uint32_t id)
{
    const uint32_t *q = b.queries.data() + ((size_t)(id / 32) * 32 * b.wpq + id % 32);
    uint32_t len = b.lens[id];
    for (uint32_t i = 0; i < len; ++i) out.push_back("ACGT"[(q[(i >> 4) * 32] >> ((i & 15) << 1)) & 3]);
    out += "\n+\n"; out += b.quals[id]; out.push_back('\n');
}

static void header_line(std::string &ret, const ReadBatch &b, uint32_t id, const Annotation &ann, std::vector<std::pair<int, int>> &chrHits,
                        int bestScore, double top, bool ignoreComments)
{
    const std::string *comment = (!ignoreComments && b.hasComment[id]) ? &b.comments[id] : nullptr;
    ret += "@"; ret += b.names[id];
    if (comment && *comment == "IGNORE") { ret += "\tIGNORE\n"; return; }
    std::sort(chrHits.begin(), chrHits.end());
    std::vector<HeaderHit> v;
    int prev = mapping_from_header(comment, v, top, bestScore * top);
    if (prev > bestScore) bestScore = prev;
    ret += "\tSCORE:" + std::to_string((long long)bestScore) + ";";
    if (bestScore > 0)
        for (size_t i = 0; i < chrHits.size(); ++i) {
            if (i > 0 && chrHits[i].first == chrHits[i - 1].first) continue;
            if (-chrHits[i].second > 0 && -chrHits[i].second >= bestScore * top)
                ret += std::to_string(-(long long)chrHits[i].second) + "," + ann.names[chrHits[i].first - 1] + ";";
        }
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i].score >= bestScore * top) { ret += comment->substr(v[i].s, v[i].e - v[i].s); ret += ";"; }
    ret += "\n";
}

// pairDeepDPOutputFastqAPI (BGS-IO.cpp:1966-2091) for the results [first, last) of one pair
static void pair_fastq(std::string &ret, const ReadBatch &b, const Annotation &ann, const mp_pair_result *first, const mp_pair_result *last,
                       int megapathMode, double top, bool ignoreComments)
{
    const uint32_t r1 = first->readID, r2 = r1 + 1;
    int best1 = 0, best2 = 0;
    std::vector<std::pair<int, int>> h1, h2;
    for (const mp_pair_result *p = first; p != last; ++p) {
        int chr1 = ann.targetChr(p->algnmt_1, b.lens[r1]), chr2 = ann.targetChr(p->algnmt_2, b.lens[r2]);
        int s1 = chr1 == -1 ? 0 : p->score_1, s2 = chr2 == -1 ? 0 : p->score_2;
        bool a1 = chr1 != -1, a2 = chr2 != -1;
        if (megapathMode == 2 && (!a1 || !a2)) { a1 = a2 = false; s1 = s2 = 0; }
        if (!a1) s1 = 0;
        if (!a2) s2 = 0;
        if (chr1 == chr2 && a1 && a2) { int sum = s1 + s2; s1 = s2 = sum; }          // normalizeScore
        if (best1 < s1) best1 = s1;
        if (best2 < s2) best2 = s2;
        if (chr1 != -1) h1.push_back(std::make_pair(chr1, -s1));
        if (chr2 != -1) h2.push_back(std::make_pair(chr2, -s2));
    }
    header_line(ret, b, r1, ann, h1, best1, top, ignoreComments); seq_and_qual(ret, b, r1);
    header_line(ret, b, r2, ann, h2, best2, top, ignoreComments); seq_and_qual(ret, b, r2);
}

// unproperlypairDPOutputFastqAPI (BGS-IO.cpp:1384-1446) for one read and its (de-duplicated) single-end hits
static void single_fastq(std::string &ret, const ReadBatch &b, const Annotation &ann, uint32_t id,
                         const std::vector<std::pair<uint64_t, int>> &hits, int megapathMode, double top, bool ignoreComments)
{
    int best = 0;
    std::vector<std::pair<int, int>> h;
    if (megapathMode != 2)
        for (const auto &a : hits) {
            int chr = ann.targetChr(a.first, b.lens[id]);
            int sc = a.second;
            if (chr < 0) sc = 0; else h.push_back(std::make_pair(chr, -sc));
            if (best < sc) best = sc;
        }
    header_line(ret, b, id, ann, h, best, top, ignoreComments); seq_and_qual(ret, b, id);
}

// ------------------------------------------------------------------------------------------------
int main(int argc, char **argv)
{
    Options opt;
    if (!parse_args(argc, argv, opt)) return 1;
    const double t0 = now_s();
    Ini ini;
    std::string iniPath = opt.iniFile.empty() ? std::string(argv[0]) + ".ini" : opt.iniFile;
    if (!ini.load(iniPath)) { fprintf(stderr, "Failed to open config file ... %s\n", iniPath.c_str()); return 1; }
    fprintf(stderr, "\n[Main] soap4 (megapath_b200, B200 hot path)\n");
    fprintf(stderr, "Number of CPU threads: %d\n", opt.numCpuThreads);
    fprintf(stderr, "[Main] Loading read files %s and %s\n", opt.query1.c_str(), opt.query2.c_str());
    if (ini.geti("OtherSettings:SkipSOAP3Alignment", 0) != 1) { fprintf(stderr, "SkipSOAP3Alignment=0 is not supported (the reference asserts, alignment.cpp:66)\n"); return 1; }
    const int maxLen = opt.maxReadLength;
    const int nRounds = maxLen > 120 ? ini.geti("DP:NumberOfRoundOfDeepDPForLongReads", 0) : ini.geti("DP:NumberOfRoundOfDeepDPForShortReads", 0);
    const char *schemeKey = maxLen > 120 ? "SeedingRound1ForLongReads:mmpSeedingScheme" : "SeedingRound1ForShortReads:mmpSeedingScheme";
    if (nRounds != 1 || ini.geti(schemeKey, 0) != 1) { fprintf(stderr, "only one deep-DP round with mmpSeedingScheme=1 is supported (DV-DPForBothUnalign.cpp:262-266)\n"); return 1; }
    mp_align_params P; mp_default_params(&P, 0);
    P.mmp.seedSAsizeThreshold = ini.geti("MMP:mmpSeedSAsizeThreshold", 30); P.mmp.seedMinLength = ini.geti("MMP:mmpSeedMinLength", 17);
    P.mmp.uniqThreshold = ini.geti("MMP:mmpUniqThreshold", 6); P.mmp.indelFuzz = ini.geti("MMP:mmpIndelFuzz", 5);
    P.mmp.goodSeedLen = ini.geti("MMP:mmpGoodSeedLen", 27); P.mmp.reseedLen = ini.geti("MMP:mmpReseedLen", 18);
    P.mmp.reseedRLTratio = ini.getd("MMP:mmpReseedRLTratio", 0.85); P.mmp.reseedAbsDiff = ini.geti("MMP:mmpReseedAbsDiff", 4);
    P.mmp.shortSeedRatio = ini.getd("MMP:mmpShortSeedRatio", 0.5);
    P.matchScore = ini.geti("DP:MatchScore", 1); P.mismatchScore = ini.geti("DP:MismatchScore", -2);
    P.openGapScore = ini.geti("DP:GapOpenScore", -3); P.extendGapScore = ini.geti("DP:GapExtendScore", -1);
    P.softClipLeft = ini.geti("Clipping:MaxFrontLenClipped", 3); P.softClipRight = ini.geti("Clipping:MaxEndLenClipped", 8);
    std::string arr = ini.gets("PairEnd:StrandArrangement", "+/-");
    P.peStrandLeftLeg = (arr == "-/+" || arr == "-/-") ? 2 : 1; P.peStrandRightLeg = (arr == "+/+" || arr == "-/+") ? 1 : 2;
    P.skipDefaultDP = ini.geti("OtherSettings:SkipDefaultDP", 0);
    P.maxReadLength = maxLen; P.insert_high = opt.insert_high;
    const double top = opt.top / 100.0;

    mp_context *gpu = nullptr;
    if (mp_init(opt.device, &gpu)) { fprintf(stderr, "%s\n", mp_last_error()); return 1; }
    fprintf(stderr, "[Main] loading index into device...\n");
    if (mp_index_load(gpu, opt.indexName.c_str())) { fprintf(stderr, "%s\n", mp_last_error()); return 1; }
    Annotation ann;
    if (!ann.load(opt.indexName)) return 1;
    fprintf(stderr, "[Main] Finished loading index into device.\n");
    const double tIndex = now_s();
    fprintf(stderr, "[Main] Loading time : %9.4f seconds\n\n", tIndex - t0);
    fprintf(stderr, "[Main] top_percentage: %f\n", top);
    fprintf(stderr, "[Main] Reference sequence length : %llu\n\n", (unsigned long long)ann.dnaLength);
    if (opt.outputBAM) fprintf(stderr, "[Main] note: BAM output (-b) is not implemented in this build; only the stdout FASTQ is produced\n");

    fill_char_map();
    SeqReader r1, r2;
    if (!r1.open(opt.query1) || !r2.open(opt.query2)) { fprintf(stderr, "Cannot open the read files\n"); return 1; }
    ReadBatch b; b.maxReadLength = (uint32_t)maxLen; b.wpq = ((uint32_t)maxLen + 15) / 16;
    const uint32_t maxNumQueries = 12 * 8192 * 128 / 6;            // SOAP4.cpp:206
    double totalLoad = 0, totalAlign = 0, last = now_s();
    uint64_t totalPairsAligned = 0;
    bool detected = false;
    std::string outbuf;
    while (load_batch(r1, r2, b, maxNumQueries) > 0) {
        const uint32_t numQueries = b.nReads, nPairs = numQueries / 2;
        fprintf(stderr, "[Main] Loaded %u short reads from the query file.\n", numQueries);
        double t = now_s();
        fprintf(stderr, "[Main] Elapsed time on host : %9.4f seconds\n\n", t - last);
        totalLoad += t - last; last = t;
        if (!detected) {
            uint32_t d1 = detect_read_length(b.lens, numQueries, 0), d2 = detect_read_length(b.lens, numQueries, 1);
            if (opt.insert_low < (int)d2) opt.insert_low = (int)d2;
            if (opt.insert_low < (int)d1) opt.insert_low = (int)d1;
            fprintf(stderr, "All reads are directly processed by DP\n");
            detected = true;
        }
        P.insert_low = opt.insert_low;
        mp_results R;
        if (mp_batch_upload(gpu, b.queries.data(), b.lens.data(), numQueries, b.wpq) || mp_align_pairs(gpu, &P, &R)) { fprintf(stderr, "%s\n", mp_last_error()); return 1; }
        fprintf(stderr, "[Main] %u pairs of reads are proceeded to deep DP Round 1.\n", nPairs);
        fprintf(stderr, "[Main] Number of pairs aligned by DP: %llu\n", (unsigned long long)R.numDPAlignedPair);
        fprintf(stderr, "[Main] Number of alignments aligned by DP: %llu\n", (unsigned long long)R.numDPAlignment);
        fprintf(stderr, "[Main] Number of reads aligned by single-end DP: %llu\n", (unsigned long long)R.numSingleDPAligned);
        fprintf(stderr, "[Main] Number of alignments aligned by DP: %llu\n", (unsigned long long)R.numSingleDPAlignment);
        fprintf(stderr, "[Main] Number of pairs aligned by DP: %llu\n", (unsigned long long)R.numRescuedPair);
        fprintf(stderr, "[Main] Number of alignments aligned by DP: %llu\n", (unsigned long long)R.numRescuedAlignment);
        totalPairsAligned += R.numDPAlignedPair + R.numRescuedPair;
        // ---- output ----
        if (opt.megapathMode) {
            std::vector<uint8_t> done(nPairs, 0);
            for (int which = 0; which < 2; ++which) {
                const mp_pair_result *arrp = which == 0 ? R.pairs : R.rescued; uint64_t n = which == 0 ? R.n_pairs : R.n_rescued;
                for (uint64_t i = 0, j; i < n; i = j) {
                    j = i + 1;
                    while (j < n && arrp[j].readID == arrp[i].readID) ++j;
                    pair_fastq(outbuf, b, ann, arrp + i, arrp + j, opt.megapathMode, top, opt.ignoreComments);
                    done[arrp[i].readID >> 1] = 1;
                    if (outbuf.size() > (1u << 22)) { fwrite(outbuf.data(), 1, outbuf.size(), stdout); outbuf.clear(); }
                }
            }
            // pairs neither placed by deep DP nor rescued: per-read single-end hits (alignment.cpp:299-351)
            uint64_t si = 0;
            for (uint32_t p = 0; p < nPairs; ++p) {
                if (done[p]) continue;
                for (uint32_t id = 2 * p; id < 2 * p + 2; ++id) {
                    while (si < R.n_singles && R.singles[si].readID < id) ++si;
                    std::vector<std::pair<uint64_t, int>> hits;
                    uint64_t e = si;
                    while (e < R.n_singles && R.singles[e].readID == id) { hits.push_back(std::make_pair(R.singles[e].algnmt, R.singles[e].score)); ++e; }
                    std::sort(hits.begin(), hits.end());                                     // OutputBuffer::ready: ResultCompare + unique
                    hits.erase(std::unique(hits.begin(), hits.end()), hits.end());
                    single_fastq(outbuf, b, ann, id, hits, opt.megapathMode, top, opt.ignoreComments);
                }
                if (outbuf.size() > (1u << 22)) { fwrite(outbuf.data(), 1, outbuf.size(), stdout); outbuf.clear(); }
            }
            fwrite(outbuf.data(), 1, outbuf.size(), stdout); outbuf.clear();
        }
        mp_results_release(gpu, &R);
        t = now_s();
        fprintf(stderr, "[Main] Elapsed time : %9.4f seconds\n\n", t - last);
        totalAlign += t - last; last = t;
    }
    fflush(stdout);
    fprintf(stderr, "[Main] Overall number of pairs of reads aligned: %llu\n", (unsigned long long)totalPairsAligned);
    fprintf(stderr, "[Main] Overall read load time : %9.4f seconds\n", totalLoad);
    fprintf(stderr, "[Main] Overall alignment time (excl. read loading) : %9.4f seconds\n", totalAlign);
    mp_destroy(gpu);
    fprintf(stderr, "[Main] Overall running time: %f\n", now_s() - t0);
    return 0;
}
