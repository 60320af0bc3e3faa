"""Host-side logic of the soap4 driver that needs no GPU: the zero-copy FASTQ fast path must see exactly the records the
kseq-style parser sees (the reference reads its input with kseq, soap4/kseq.h via QueryParser.cpp)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "megapath_b200", "bin", "soap4")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="driver binary not built (run __graft_entry__.build())")


def parse(path, generic):
    return subprocess.run([EXE, "__parse", path] + (["generic"] if generic else []), capture_output=True, check=True, timeout=120).stdout


CASES = {
    "plain": b"@r1/1\nACGT\n+\nIIII\n@r2/1 extra words\nGGCA\n+r2\nII#I\n",
    "crlf": b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2 c\r\nAC\r\n+\r\nII\r\n",
    "no_final_newline": b"@r1\nACGT\n+\nIIII\n@r2\nAC\n+\nII",
    "blank_lines": b"\n\n@r1\nACGT\n+\nIIII\n\n@r2\nAC\n+\nII\n\n",
    "multiline": b"@r1\nACGT\nACGT\n+\nIIII\nIIII\n@r2\nAC\n+\nII\n",
    "at_in_quality": b"@r1\nACGT\n+\n@III\n@r2\nACGA\n+\n@@@@\n@r3\nAC\n+\nI@\n",
    "qual_too_long": b"@r1\nACGT\n+\nIIIII\n@r2\nAC\n+\nII\n",
    "qual_short_continues": b"@r1\nACGT\n+\nII\nII\n@r2\nAC\n+\nII\n",
    "fasta": b">s1 desc\nACGTAC\nGT\n>s2\nAAA\n",
    "tabs": b"@r1\tSCORE:10;5,chr1;\nACGT\n+\nIIII\n@r2 \nAC\n+\nII\n",
    "empty": b"",
    "garbage_prefix": b"junk\n@r1\nAC\n+\nII\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_fast_path_equals_generic(tmp_path, name):
    p = tmp_path / (name + ".fq")
    p.write_bytes(CASES[name])
    assert parse(str(p), False) == parse(str(p), True)


def test_fast_path_large_and_gz(tmp_path):
    """records straddling the 4 MiB buffer boundary, plain and gzip"""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(60000):
        L = int(rng.integers(30, 152))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L).tolist())
        qual = bytes(rng.integers(33, 74, L).astype(np.uint8).tolist())
        recs.append(b"@read%d/1%s\n%s\n+\n%s\n" % (i, b" c:%d" % i if i % 7 == 0 else b"", seq, qual))
    data = b"".join(recs)
    p = tmp_path / "big.fq"
    p.write_bytes(data)
    want = parse(str(p), True)
    assert want.count(b"\n") == 60000
    assert parse(str(p), False) == want
    g = tmp_path / "big.fq.gz"
    with gzip.open(g, "wb", compresslevel=1) as f:
        f.write(data)
    assert parse(str(g), False) == want


@pytest.mark.parametrize("size", [0, 5, 70000, 1 << 20])
def test_bgzf_chunks_concatenate_to_a_valid_stream(tmp_path, size):
    """BAM records are compressed into BGZF blocks by the formatting threads, chunk by chunk; the concatenation (after the
    buffered header bytes, before the EOF block) must decompress to the original bytes and every block must be a proper BGZF member"""
    rng = np.random.default_rng(size)
    data = bytes(rng.integers(0, 7, size).astype(np.uint8).tolist())
    src, dst = tmp_path / "in.bin", tmp_path / "out.bgzf"
    src.write_bytes(data)
    subprocess.run([EXE, "__bgzf", str(src), str(dst)], check=True, timeout=120)
    raw = dst.read_bytes()
    assert gzip.decompress(raw) == data
    # walk the members: gzip magic, FEXTRA with the BC subfield, BSIZE consistent, ISIZE <= 0xff00; the last one is the 28-byte EOF marker
    at, members = 0, 0
    while at < len(raw):
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04" and raw[at + 12:at + 16] == b"BC\x02\x00"
        bsize = int.from_bytes(raw[at + 16:at + 18], "little") + 1
        isize = int.from_bytes(raw[at + bsize - 4:at + bsize], "little")
        assert isize <= 0xff00
        at += bsize
        members += 1
    assert at == len(raw) and raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    assert members >= 2 if size else members == 1
