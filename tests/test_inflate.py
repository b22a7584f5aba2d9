"""The gzip decoder of the host reader (host/gt_inflate.{h,cpp}: gt::GzipStream, through bin/gt_gunzip) against zlib: every block type
(stored, fixed, dynamic), compression levels and strategies, window sizes, header flags, several members, bytes behind the last
member, output and input that straddle the decoder's chunk (1 MB) and buffer (2 MB) boundaries, read sizes from 1 byte up -- and files
cut at random places, where the decoder hands out what zlib hands out: everything that inflates before the end of the file.  Damaged
files must end the stream without a crash.  No GPU needed."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import support

TOOL = os.path.join(support.ROOT, "ibm-cbc-genomic-tools_b200", "bin", "gt_gunzip")
DUMP = os.path.join(support.ROOT, "ibm-cbc-genomic-tools_b200", "bin", "gt_regdump")
pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="bin/gt_gunzip not built")


def gunzip(tmp_path, data, read_size=None):
    p = tmp_path / "x.gz"
    p.write_bytes(data)
    r = subprocess.run([TOOL, str(p)] + ([str(read_size)] if read_size else []), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode in (0, 1), (r.returncode, r.stderr[-500:])
    return r.returncode, r.stdout


def member(raw, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15, memlevel=8, flags=0):
    c = zlib.compressobj(level, zlib.DEFLATED, -wbits, memlevel, strategy)
    body = c.compress(raw) + c.flush()
    hdr = b"\x1f\x8b\x08" + bytes([flags]) + b"\0\0\0\0\0\xff"
    if flags & 4:
        hdr += struct.pack("<H", 7) + b"AB\x03\x00xyz"
    if flags & 8:
        hdr += b"file name.txt\0"
    if flags & 16:
        hdr += b"a comment\0"
    if flags & 2:
        hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    return hdr + body + struct.pack("<II", zlib.crc32(raw) & 0xFFFFFFFF, len(raw) & 0xFFFFFFFF)


def zlib_prefix(data):
    """what zlib hands out for a (possibly truncated) series of members"""
    out, rest = b"", data
    while rest[:2] == b"\x1f\x8b":
        d = zlib.decompressobj(31)
        out += d.decompress(rest)
        if not d.eof:
            break
        rest = d.unused_data
    return out


def payload(rng, kind, n):
    if kind == "rand":
        return rng.bytes(n)
    if kind == "text":
        return b"".join(b"chr%d\t%d\t%d\tr%d\t0\t%s\n" % (rng.integers(1, 23), rng.integers(10 ** 8), rng.integers(10 ** 8), i, b"+-"[i % 2:i % 2 + 1])
                        for i in range(n // 30 + 1))[:n]
    if kind == "rle":
        return bytes([int(rng.integers(4))]) * (n // 2) + rng.bytes(3) * (n // 6 + 1)
    parts, have = [], 0
    while have < n:
        k = int(rng.integers(4))
        parts.append([rng.bytes(int(rng.integers(1, 70000))), b"ab" * int(rng.integers(1, 40000)), payload(rng, "text", int(rng.integers(1, 90000))),
                      bytes(int(rng.integers(1, 300000)))][k])
        have += len(parts[-1])
    return b"".join(parts)[:n]


CODINGS = [(0, 0, 15), (1, 0, 15), (6, 0, 15), (9, 0, 15), (6, zlib.Z_FIXED, 15), (6, zlib.Z_HUFFMAN_ONLY, 15), (6, zlib.Z_RLE, 15), (9, zlib.Z_FILTERED, 9), (4, 0, 12)]


@pytest.mark.parametrize("kind", ["rand", "text", "rle", "mixed"])
def test_members_whole_cut_and_damaged(tmp_path, kind):
    rng = np.random.default_rng({"rand": 1, "text": 2, "rle": 3, "mixed": 4}[kind])
    for n in (0, 1, 2, 100, 32767, 32769, 65536, (1 << 20) - 258, (1 << 20) + 1, 2_500_017):
        raw = payload(rng, kind, n)
        for level, strategy, wbits in CODINGS if n < 100_000 else [CODINGS[i] for i in rng.choice(len(CODINGS), 3, replace=False)]:
            data = member(raw, level, strategy, wbits, int(rng.choice([1, 8, 9])), int(rng.choice([0, 0, 8, 4 | 8 | 16, 2, 4 | 2])))
            rc, out = gunzip(tmp_path, data, int(rng.choice([0, 1, 7, 4096, 65536, (1 << 20) + 3])) if n < 100_000 else None)
            assert rc == 0 and out == raw, ("whole", n, level, strategy, wbits)
            for _ in range(2):
                cut = int(rng.integers(0, len(data)))
                rc, out = gunzip(tmp_path, data[:cut])
                assert out == zlib_prefix(data[:cut]), ("cut", n, level, strategy, wbits, cut)
            if len(data) > 30:
                bad = bytearray(data)
                for _ in range(int(rng.integers(1, 4))):
                    bad[int(rng.integers(10, len(bad)))] ^= 1 << int(rng.integers(8))
                gunzip(tmp_path, bytes(bad))                                  # (ends without a crash, whatever it hands out)


def test_several_members_and_what_follows_them(tmp_path):
    rng = np.random.default_rng(9)
    for trial in range(12):
        raws = [payload(rng, ["rand", "text", "rle", "mixed"][int(rng.integers(4))], int(rng.choice([0, 1, 1000, 70000, 1 << 20, 1_500_000]))) for _ in range(int(rng.integers(1, 5)))]
        data = b"".join(member(r, int(rng.choice([0, 1, 6, 9])), flags=int(rng.choice([0, 8, 4]))) for r in raws)
        rc, out = gunzip(tmp_path, data)
        assert rc == 0 and out == b"".join(raws)
        rc, out = gunzip(tmp_path, data + b"\0\0\0bytes that are no gzip header" * int(rng.integers(1, 5)))
        assert out == b"".join(raws)                                          # ignored, as zlib ignores them
        cut = int(rng.integers(0, len(data) + 1))
        rc, out = gunzip(tmp_path, data[:cut])
        assert out == zlib_prefix(data[:cut]), (trial, cut)
    # a wrong CRC or length is a fault
    raw = payload(rng, "text", 50_000)
    data = bytearray(member(raw))
    data[-6] ^= 0x10
    assert gunzip(tmp_path, bytes(data))[0] == 1
    data = bytearray(member(raw))
    data[-3] ^= 0x01
    assert gunzip(tmp_path, bytes(data))[0] == 1


def test_reader_on_gzip_files(tmp_path):
    """the line reader on a gzip file -- this build's decoder, zlib's gzread (GT_ZLIB=1) and the plain file give the same regions"""
    import gzip
    rng = np.random.default_rng(10)
    text = payload(rng, "text", 9_000_000)
    text = text[:text.rindex(b"\n") + 1]
    (tmp_path / "r.bed").write_bytes(text)
    with gzip.open(tmp_path / "r.bed.gz", "wb", compresslevel=5) as f:
        f.write(text)

    def dump(path, env):
        p = subprocess.run([DUMP, str(path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **env))
        return p.returncode, p.stdout
    want = dump(tmp_path / "r.bed", {})
    assert want[0] == 0 and want[1].count(b"\n") > 200_000
    assert dump(tmp_path / "r.bed.gz", {}) == want
    assert dump(tmp_path / "r.bed.gz", {"GT_ZLIB": "1"}) == want
    whole = (tmp_path / "r.bed.gz").read_bytes()
    (tmp_path / "cut.bed.gz").write_bytes(whole[:len(whole) // 2 + 3])
    a, b = dump(tmp_path / "cut.bed.gz", {}), dump(tmp_path / "cut.bed.gz", {"GT_ZLIB": "1"})
    assert a == b and a[1].count(b"\n") > 50_000
