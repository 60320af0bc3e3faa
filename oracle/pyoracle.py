"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the reference
shims under oracle/_ref/.  TEST INFRASTRUCTURE ONLY: importable from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import struct
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class MmpParams(C.Structure):
    _fields_ = [("seedSAsizeThreshold", C.c_int32), ("seedMinLength", C.c_int32),
                ("uniqThreshold", C.c_int32), ("indelFuzz", C.c_int32),
                ("goodSeedLen", C.c_int32), ("reseedLen", C.c_int32),
                ("reseedRLTratio", C.c_double), ("reseedAbsDiff", C.c_int32),
                ("shortSeedRatio", C.c_double)]


def mmp_params(nt2=False):
    """[MMP] section of soap4.ini / soap4-nt2.ini."""
    return MmpParams(30, 17 if nt2 else 22, 6, 5, 27, 18 if nt2 else 23, 0.7, 4, 0.5)


SEEDSA = np.dtype([("query_offset", "<u4"), ("pad0", "<u4"), ("sa_l", "<u8"), ("sa_diff", "<u4"), ("seed_len", "<u4")])
SEEDPOS = np.dtype([("pos", "<u8"), ("strand_readID", "<u4"), ("paired_seedLength", "<u4")])
CAND = np.dtype([("readIDLeft", "<u4"), ("pad", "<u4"), ("pos0", "<u8"), ("pos1", "<u8")])

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_index_load.restype = C.c_void_p
        L.orc_index_load.argtypes = [C.c_char_p]
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_text_length.restype = C.c_uint64
        L.orc_text_length.argtypes = [C.c_void_p]
        L.orc_inverse_sa0.restype = C.c_uint64
        L.orc_inverse_sa0.argtypes = [C.c_void_p]
        L.orc_cum_freq.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_occ_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sa_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_lkt.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_text_base.restype = C.c_uint32
        L.orc_text_base.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_counters.argtypes = [C.c_void_p, C.c_int]
        L.orc_text_window.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_mmp.restype = C.c_int
        L.orc_mmp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_seed_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pair_candidates.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                          C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_dp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Index:
    def __init__(self, prefix):
        self.L = lib()
        self.h = self.L.orc_index_load(prefix.encode())
        self.n = self.L.orc_text_length(self.h)
        self.inverse_sa0 = self.L.orc_inverse_sa0(self.h)

    def occ(self, idx, c):
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.uint32)
        out = np.empty(len(idx), dtype=np.uint64)
        self.L.orc_occ_many(self.h, len(idx), ptr(idx), ptr(c), ptr(out))
        return out

    def sa(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        out = np.empty(len(idx), dtype=np.uint64)
        self.L.orc_sa_many(self.h, len(idx), ptr(idx), ptr(out))
        return out

    def lkt(self, key):
        l, r = C.c_uint64(), C.c_uint64()
        self.L.orc_lkt(self.h, int(key), C.byref(l), C.byref(r))
        return l.value, r.value

    def text(self, start, length):
        out = np.empty(length, dtype=np.uint8)
        self.L.orc_text_window(self.h, int(start), int(length), ptr(out))
        return out

    def counters(self, reset=False):
        out = np.zeros(4, dtype=np.uint64)
        self.L.orc_counters(ptr(out), int(reset))
        return out

    def mmp(self, read, strand, params):
        read = np.ascontiguousarray(read, dtype=np.uint8)
        out = np.zeros(1024, dtype=SEEDSA)
        n = self.L.orc_mmp(self.h, ptr(read), len(read), strand, C.byref(params), ptr(out), len(out))
        return out[:n]

    def seed_pairs(self, reads, lens, params):
        """reads: (2*nPairs, maxLen) uint8 codes; returns (readPos, matePos) SEEDPOS arrays incl. sentinels."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        rp, mp = C.c_void_p(), C.c_void_p()
        nr, nm = C.c_uint64(), C.c_uint64()
        self.L.orc_seed_pairs(self.h, ptr(reads), ptr(lens), reads.shape[1], reads.shape[0] // 2, C.byref(params),
                              C.byref(rp), C.byref(nr), C.byref(mp), C.byref(nm))
        a = np.frombuffer((C.c_char * (nr.value * 16)).from_address(rp.value), dtype=SEEDPOS).copy()
        b = np.frombuffer((C.c_char * (nm.value * 16)).from_address(mp.value), dtype=SEEDPOS).copy()
        self.L.orc_free(rp)
        self.L.orc_free(mp)
        return a, b


def pair_candidates(readPos, matePos, lens, insert_low, insert_high):
    L = lib()
    rp = readPos.copy()
    mp = matePos.copy()
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    out, n = C.c_void_p(), C.c_uint64()
    L.orc_pair_candidates(ptr(rp), len(rp), ptr(mp), len(mp), ptr(lens), insert_low, insert_high, C.byref(out), C.byref(n))
    res = np.frombuffer((C.c_char * (n.value * 24)).from_address(out.value), dtype=CAND).copy() if n.value else np.zeros(0, dtype=CAND)
    L.orc_free(out)
    return res


def dp_cutoff(read_len):
    """max(0.2*L, 30) truncated (definitions.h:166-167; DV-DPfunctions.cpp:2916)."""
    return int(max(0.2 * read_len, 30.0))


def dp(ref, read, clip_lt=130, clip_rt=130, mismatch=-2, gap_open=-3, cutoff=None):
    L = lib()
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    read = np.ascontiguousarray(read, dtype=np.uint8)
    if cutoff is None:
        cutoff = dp_cutoff(len(read))
    score, hit, cnt = C.c_int(), C.c_uint32(), C.c_uint32()
    pat = np.zeros(len(ref) + len(read) + 16, dtype=np.uint8)
    L.orc_dp(ptr(ref), len(ref), ptr(read), len(read), clip_lt, clip_rt, mismatch, gap_open, cutoff,
             C.byref(score), C.byref(hit), C.byref(cnt), ptr(pat))
    return score.value, hit.value, cnt.value, pattern_bytes(pat) if score.value >= cutoff else b""


def pattern_bytes(pat):
    """Pattern up to its terminator; a 'V' is followed by a count byte that may be 0."""
    i = 0
    while pat[i] != 0:
        i += 2 if pat[i] == ord("V") else 1
    return bytes(pat[:i])


# ------------------------------------------------------------------ reference shims
_ref_dp = None
_ref_bwt = None


def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "soap4"))


def ref_dp_lib():
    global _ref_dp
    if _ref_dp is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_dp.so"))
        L.ref_dp_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _ref_dp = L
    return _ref_dp


def ref_dp(refs, dna_lens, reads, read_lens, max_dna, max_read, clip_lt=130, clip_rt=130, mismatch=-2, gap_open=-3):
    """Runs the reference's own callDP on n tasks.  refs: (n, max_dna) codes, reads: (n, max_read)."""
    L = ref_dp_lib()
    n = len(dna_lens)
    refs = np.ascontiguousarray(refs, dtype=np.uint8)
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    dna_lens = np.ascontiguousarray(dna_lens, dtype=np.uint32)
    read_lens = np.ascontiguousarray(read_lens, dtype=np.uint32)
    scores = np.zeros(n, dtype=np.int32)
    hits = np.zeros(n, dtype=np.uint32)
    cnts = np.zeros(n, dtype=np.uint32)
    pats = np.zeros((n, max_dna + max_read), dtype=np.uint8)
    L.ref_dp_run(n, max_dna, max_read, ptr(refs), ptr(dna_lens), ptr(reads), ptr(read_lens),
                 clip_lt, clip_rt, mismatch, gap_open, ptr(scores), ptr(hits), ptr(cnts), ptr(pats))
    return scores, hits, cnts, pats


class RefIndex:
    def __init__(self, prefix):
        global _ref_bwt
        if _ref_bwt is None:
            L = C.CDLL(os.path.join(REF_DIR, "libref_bwt.so"))
            L.ref_index_load.restype = C.c_void_p
            L.ref_index_load.argtypes = [C.c_char_p]
            L.ref_text_length.restype = C.c_uint64
            L.ref_text_length.argtypes = [C.c_void_p]
            L.ref_inverse_sa0.restype = C.c_uint64
            L.ref_inverse_sa0.argtypes = [C.c_void_p]
            L.ref_occ.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_sa.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.ref_lkt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            _ref_bwt = L
        self.L = _ref_bwt
        self.h = self.L.ref_index_load(prefix.encode())
        self.n = self.L.ref_text_length(self.h)

    def occ(self, idx, c):
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.uint32)
        out = np.empty(len(idx), dtype=np.uint64)
        self.L.ref_occ(self.h, len(idx), ptr(idx), ptr(c), ptr(out))
        return out

    def sa(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        out = np.empty(len(idx), dtype=np.uint64)
        self.L.ref_sa(self.h, len(idx), ptr(idx), ptr(out))
        return out

    def lkt(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        l = np.empty(len(keys), dtype=np.uint64)
        r = np.empty(len(keys), dtype=np.uint64)
        self.L.ref_lkt(self.h, len(keys), ptr(keys), ptr(l), ptr(r))
        return l, r


# ------------------------------------------------------------------ dump readers (oracle/ref_hooks.h)
def read_seedpos_dump(path):
    """-> list of (readPos, matePos) per batch."""
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(16)
            if len(h) < 16:
                break
            nr, nm = struct.unpack("<QQ", h)
            a = np.frombuffer(f.read(16 * nr), dtype=SEEDPOS)
            b = np.frombuffer(f.read(16 * nm), dtype=SEEDPOS)
            out.append((a, b))
    return out


def read_cand_dump(path):
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(8)
            if len(h) < 8:
                break
            (n,) = struct.unpack("<Q", h)
            out.append(np.frombuffer(f.read(24 * n), dtype=CAND))
    return out


def read_dp_dump(path, limit=None):
    """-> list of dict(header..., tasks=[(ref, read, cutoff, score, hitLoc, count, pattern)])."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    o = 0
    while o + 32 <= len(data):
        magic, n, maxdna, maxread, cl, cr, mm, go = struct.unpack_from("<8I", data, o)
        assert magic == 0x5044504D
        o += 32
        tasks = []
        for _ in range(n):
            dl, rl, co, sc, hl, mc = struct.unpack_from("<IIiiII", data, o)
            o += 24
            ref = np.frombuffer(data, dtype=np.uint8, count=dl, offset=o)
            o += dl
            rd = np.frombuffer(data, dtype=np.uint8, count=rl, offset=o)
            o += rl
            (pl,) = struct.unpack_from("<H", data, o)
            o += 2
            pat = data[o:o + pl]
            o += pl
            tasks.append((ref, rd, co, sc, hl, mc, pat))
        out.append(dict(n=n, max_dna=maxdna, max_read=maxread, clip_lt=cl, clip_rt=cr,
                        mismatch=C.c_int32(mm).value, gap_open=C.c_int32(go).value, tasks=tasks))
        if limit and len(out) >= limit:
            break
    return out


def read_fastq_codes(path, max_len=None, trunc=None):
    """FASTQ -> (codes (n, maxLen) uint8 with N->G, lens).  `trunc`: reads are truncated to
    trunc bases (QueryParser.cpp:188 truncates to -L minus 1)."""
    seqs = []
    with open(path, "rb") as f:
        for i, line in enumerate(f):
            if i % 4 == 1:
                s = line.strip()
                if trunc is not None:
                    s = s[:trunc]
                seqs.append(s)
    lut = np.zeros(256, dtype=np.uint8)   # unknown -> A(0); N -> G (IndexHandler.cpp:26-45)
    for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut[ch] = v
    lut[ord("U")] = lut[ord("u")] = 3
    lut[ord("N")] = lut[ord("n")] = 2
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    ml = max_len or int(lens.max())
    out = np.zeros((len(seqs), ml), dtype=np.uint8)
    for i, s in enumerate(seqs):
        out[i, :len(s)] = lut[np.frombuffer(s, dtype=np.uint8)]
    return out, lens


# ------------------------------------------------------------------ deep-DP stage (S1) restatement
def margin(read_len):
    """DP2_MARGIN (DV-DPfunctions.cpp:1760)."""
    return 30 if read_len > 100 else 25


def encode_cigar(pattern, gap_open=-3, gap_ext=-1):
    """CigarStringEncoder (DV-DPfunctions.h:344-427) fed as in DP2CPUAlgnThread
    (DV-DPfunctions.cpp:3447-3462).  -> (cigar bytes, counts dict, gapPenalty)"""
    runs = []                      # merged (type, cnt) in pattern (end -> start) order
    last_type = ord("N")
    i = 0
    while i < len(pattern):
        ch = pattern[i]
        if ch == ord("V"):
            t, c = last_type, pattern[i + 1] - 1
            i += 2
        else:
            t, c = ch, 1
            last_type = ch
            i += 1
        if runs and runs[-1][0] == t:
            runs[-1][1] += c
        else:
            runs.append([t, c])
    out = b""
    counts = {}
    gap = 0
    for t, c in reversed(runs):
        if c > 0:
            out += b"%d%c" % (c, t)
            counts[chr(t)] = counts.get(chr(t), 0) + c
            if chr(t) in "ID":
                gap += gap_open + (c - 1) * gap_ext
    return out, counts, gap


def c_div(a, b):
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def read_for_strand(read, strand):
    return read if strand == 1 else (3 - read[::-1]).astype(np.uint8)


def deep_dp(ix, reads, lens, cands, insert_low, insert_high, input_max_read_length,
            clip_l=130, clip_r=130, mm=-2, gap_open=-3):
    """PairEndAlgnBatch::packLeft/packRight + DP2CPUAlgnThread result assembly
    (DV-DPfunctions.cpp:2857-3007, 3391-3540) for StrandArrangement +/-; then per pair
    OutputBuffer::ready (DV-DPfunctions.h:198-243).  -> list of result dicts grouped by pair."""
    n = ix.n
    max_read = (input_max_read_length // 4 + 1) * 4
    max_dna = max_read + 2 * margin(input_max_read_length) + 8
    out = []
    cells = 0
    for cd in cands:
        rid_l = int(cd["readIDLeft"])
        rl = int(lens[rid_l])
        m = margin(rl)
        start_l = (int(cd["pos0"]) - m) & 0xFFFFFFFFFFFFFFFF
        if start_l >= n:
            start_l = 0
        dl = rl + 2 * m
        if start_l + dl > n:
            dl = n - start_l
        cut_l = dp_cutoff(rl)
        sL = dp(ix.text(start_l, dl), read_for_strand(reads[rid_l, :rl], 1), clip_l, clip_r, mm, gap_open, cut_l)
        cells += dl * rl
        if sL[0] < cut_l:
            continue
        rid_r = rid_l ^ 1
        rr = int(lens[rid_r])
        m = margin(rr)
        start_r = (int(cd["pos1"]) - m) & 0xFFFFFFFFFFFFFFFF
        if start_r >= n:
            start_r = 0
        dr = rr + 2 * m
        if start_r + dr > n:
            dr = n - start_r
        hit_left = start_l + sL[1]
        bounded = (hit_left + insert_high - start_r) & 0xFFFFFFFFFFFFFFFF
        if bounded < dr:
            dr = bounded
        cut_r = dp_cutoff(rr)
        sR = dp(ix.text(start_r, dr), read_for_strand(reads[rid_r, :rr], 2), clip_r, clip_l, mm, gap_open, cut_r)
        cells += dr * rr
        if sR[0] < cut_r:
            continue
        side = []
        for (sc, hl, cnt, pat), start, dlen in ((sL, start_l, dl), (sR, start_r, dr)):
            cig, counts, gap = encode_cigar(pat, gap_open, -1)
            L = rr - counts.get("I", 0) - counts.get("S", 0)          # batch->lengths[i] is the right read's
            nmis = c_div(L * 1 + gap - sc, 1 - mm)
            side.append(dict(pos=start + hl, score=sc, cigar=cig, count=cnt, start=start, dlen=dlen,
                             editdist=counts.get("I", 0) + counts.get("D", 0) + nmis,
                             dis=counts.get("D", 0) - counts.get("I", 0) - counts.get("S", 0)))
        read_side = rid_l & 1
        a, b = side[read_side], side[1 - read_side]
        ins = (abs(b["pos"] - a["pos"]) + rr + side[1]["dis"]) & 0xFFFFFFFF
        if ins >= 1 << 31:
            ins -= 1 << 32
        right_anchor = hit_left + insert_low - start_r
        la = [max_dna, max_dna]
        ra = [0, right_anchor if right_anchor > 0 else 0]
        out.append(dict(readID=rid_l - read_side, insertSize=ins,
                        algnmt_1=a["pos"], algnmt_2=b["pos"], score_1=a["score"], score_2=b["score"],
                        editdist_1=a["editdist"], editdist_2=b["editdist"], cigar_1=a["cigar"], cigar_2=b["cigar"],
                        num_sameScore_1=a["count"], num_sameScore_2=b["count"],
                        strand_1=1 if read_side == 0 else 2, strand_2=2 if read_side == 0 else 1,
                        startPos_1=a["start"] & 0xFFFFFFFF, startPos_2=b["start"],
                        refDpLength_1=a["dlen"], refDpLength_2=b["dlen"],
                        peLeftAnchor_1=la[read_side], peLeftAnchor_2=la[1 - read_side],
                        peRightAnchor_1=ra[read_side], peRightAnchor_2=ra[1 - read_side]))
    # per pair: sort by (algnmt_1, algnmt_2, score_1, score_2), drop exact duplicates
    res = []
    i = 0
    while i < len(out):
        j = i
        while j < len(out) and out[j]["readID"] == out[i]["readID"]:
            j += 1
        grp = sorted(out[i:j], key=lambda r: (r["algnmt_1"], r["algnmt_2"], r["score_1"], r["score_2"]))
        keep = [grp[0]]
        for r in grp[1:]:
            k0 = (keep[-1]["algnmt_1"], keep[-1]["algnmt_2"], keep[-1]["score_1"], keep[-1]["score_2"])
            if k0 < (r["algnmt_1"], r["algnmt_2"], r["score_1"], r["score_2"]):
                keep.append(r)
        res.extend(keep)
        i = j
    return res, cells


# ---- stdout contract: the SCORE: header of one read (test infrastructure, restated from the reference) ----
def c_atoi(b):
    """glibc atoi on the bytes b (taken as NUL terminated): (int) strtol(b, NULL, 10) -- blanks, one sign, digits; saturates at the
    long limits before the cast (getMappingFromHeader calls it on whatever follows "SCORE:", BGS-IO.cpp:1353, 1361)"""
    i, n = 0, len(b)
    while i < n and b[i] in b" \t\n\v\f\r":
        i += 1
    neg = False
    if i < n and b[i] in b"+-":
        neg = b[i] == ord("-")
        i += 1
    v = 0
    while i < n and 48 <= b[i] <= 57:
        v = v * 10 + (b[i] - 48)
        i += 1
    v = -v if neg else v
    v = max(-(1 << 63), min((1 << 63) - 1, v))
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def mapping_from_header(comment, top, score_t):
    """getMappingFromHeader (BGS-IO.cpp:1348-1371) -> (previous best score, [(score, entry bytes)]).  comment: bytes after the first
    blank of the header, None when there is none (or with -nc).  Comments the reference cannot read (shorter than "SCORE:", an
    unterminated last entry: it reads past the string / dereferences NULL) are not defined here either."""
    if comment is None or comment == b"IGNORE":
        return 0, []
    score = c_atoi(comment[6:])
    if score < score_t:
        return score, []
    if score_t < score * top:
        score_t = score * top
    out = []
    p = comment.find(b";", 6)
    while 0 <= p and p + 1 < len(comment):
        s = p + 1
        m = c_atoi(comment[s:])
        p = comment.find(b";", s)
        if p < 0:
            break
        if score >= score_t:
            out.append((m, comment[s:p]))
    return score, out


def fastq_header(name, comment, own_best, own_hits, top):
    """The header line pairDeepDPOutputFastqAPI / unproperlypairDPOutputFastqAPI print for one read (BGS-IO.cpp:1384-1446,
    1966-2091): own_best = best score of this run's alignments of the read, own_hits = [(sequence id, score, sequence name)] of the
    alignments that lie within one sequence; comment = the previous chunk's comment (None with -nc)."""
    if comment == b"IGNORE":
        return b"@" + name + b"\tIGNORE"
    hits = sorted((cid, -sc, nm) for cid, sc, nm in own_hits)
    prev, kept = mapping_from_header(comment, top, own_best * top)
    best = max(own_best, prev)
    out = b"@" + name + b"\tSCORE:%d;" % best
    if best > 0:
        last = None
        for cid, nsc, nm in hits:
            if cid == last:
                continue
            last = cid
            if -nsc > 0 and -nsc >= best * top:
                out += b"%d,%s;" % (-nsc, nm)
    for m, entry in kept:
        if m >= best * top:
            out += entry + b";"
    return out
