"""Shared test/bench infrastructure: synthetic hg19-shaped inputs, BED writers, the oracle binding
and the reference-binary runner.  TEST INFRASTRUCTURE ONLY -- the product never imports this."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

# UCSC hg19 chromosome sizes; ids are assigned in strcmp order of the names,
# which is the order the reference iterates its std::map (genomic_intervals.cpp:5058, :5099).
HG19 = {
    "chr1": 249250621, "chr2": 243199373, "chr3": 198022430, "chr4": 191154276, "chr5": 180915260,
    "chr6": 171115067, "chr7": 159138663, "chr8": 146364022, "chr9": 141213431, "chr10": 135534747,
    "chr11": 135006516, "chr12": 133851895, "chr13": 115169878, "chr14": 107349540,
    "chr15": 102531392, "chr16": 90354753, "chr17": 81195210, "chr18": 78077248, "chr19": 59128983,
    "chr20": 63025520, "chr21": 48129895, "chr22": 51304566, "chrX": 155270560, "chrY": 59373566,
    "chrM": 16571,
}
HG19_NAMES = sorted(HG19.keys())                     # byte-lexicographic == strcmp for ASCII
HG19_LENS = np.array([HG19[n] for n in HG19_NAMES], dtype=np.int64)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    """Vectorised splitmix64 finaliser over uint64 (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synth_reads(n, seed, read_len=50, chrom_lens=HG19_LENS, first=0, p_range=None):
    """Counter-based synthetic reads: read i depends only on (seed, i) so the CUDA generator
    (gtb_synth_reads) and this function produce identical streams.
      a = splitmix64(seed * 0x9E3779B97F4A7C15 + i); b = splitmix64(a)
      p = a mod sum_c(len_c - read_len + 1); chrom = interval of p; start = offset + 1
      stop = start + read_len - 1; strand = '-' if (b & 1) else '+'."""
    lens = np.asarray(chrom_lens, dtype=np.int64)
    eff = np.maximum(lens - read_len + 1, 0).astype(np.uint64)
    cum = np.concatenate([[0], np.cumsum(eff)]).astype(np.uint64)
    i = np.arange(first, first + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        a = splitmix64((np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + i) & _M64)
    b = splitmix64(a)
    if p_range is None:
        p = a % cum[-1]
    else:                                             # one genome shard: gtb_synth_reads_range
        p_lo, p_hi = np.uint64(p_range[0]), np.uint64(min(p_range[1], int(cum[-1])))
        p = p_lo + a % (p_hi - p_lo)
    chrom = (np.searchsorted(cum, p, side="right") - 1).astype(np.int32)
    start = (p - cum[chrom] + np.uint64(1)).astype(np.int32)
    stop = (start + np.int32(read_len - 1)).astype(np.int32)
    strand = np.where((b & np.uint64(1)) != 0, ord("-"), ord("+")).astype(np.int8)
    return {"chrom": chrom, "start": start, "stop": stop, "strand": strand}


def synth_regions(m, seed, min_len=500, max_len=500_000, chrom_lens=HG19_LENS):
    """Gene-like regions: chromosome ~ length, start uniform, length log-uniform, clipped to the
    chromosome; overlapping allowed; strand uniform (host-side only, numpy Generator)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.asarray(chrom_lens, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(lens)])
    p = rng.integers(0, cum[-1], size=m, dtype=np.int64)
    chrom = (np.searchsorted(cum, p, side="right") - 1).astype(np.int32)
    start = (p - cum[chrom] + 1).astype(np.int64)
    length = np.floor(np.exp(rng.uniform(np.log(min_len), np.log(max_len), size=m))).astype(np.int64)
    stop = np.minimum(start + length - 1, lens[chrom])
    strand = np.where(rng.integers(0, 2, size=m) == 1, ord("-"), ord("+")).astype(np.int8)
    return {"chrom": chrom, "start": start.astype(np.int32), "stop": stop.astype(np.int32), "strand": strand}


# ---------------------------------------------------------------------------------------------
# text writers (the reference binaries read files)
# ---------------------------------------------------------------------------------------------

def write_bed(path, s, names, labels=None, sep="\t", trailing_newline=True):
    """Single-interval set -> BED6 (0-based half-open on disk: start-1, stop)."""
    n = len(s["chrom"])
    with open(path, "w") as f:
        lines = []
        for k in range(n):
            lab = labels[k] if labels is not None else "r%d" % k
            lines.append(sep.join([names[s["chrom"][k]], str(int(s["start"][k]) - 1), str(int(s["stop"][k])),
                                   str(lab), "0", chr(int(s["strand"][k]))]))
        f.write("\n".join(lines))
        if trailing_newline and lines:
            f.write("\n")


def write_reg(path, s, names, labels=None, offsets=None):
    """Region set -> REG (LABEL <TAB> chrom strand start stop [chrom strand start stop]...)."""
    nreg = len(offsets) - 1 if offsets is not None else len(s["chrom"])
    with open(path, "w") as f:
        for k in range(nreg):
            lo, hi = (offsets[k], offsets[k + 1]) if offsets is not None else (k, k + 1)
            lab = labels[k] if labels is not None else "r%d" % k
            toks = []
            for i in range(lo, hi):
                toks += [names[s["chrom"][i]], chr(int(s["strand"][i])), str(int(s["start"][i])), str(int(s["stop"][i]))]
            f.write("%s\t%s\n" % (lab, " ".join(toks)))


def run_ref(tool, args, stdin=None, check=True):
    """Run a reference binary from oracle/_ref; returns (returncode, stdout bytes, stderr bytes)."""
    exe = os.path.join(REF_DIR, tool)
    p = subprocess.run([exe] + [str(a) for a in args], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if check and p.returncode != 0:
        raise RuntimeError("%s %s failed (%d): %s" % (tool, args, p.returncode, p.stderr.decode()[-500:]))
    return p.returncode, p.stdout, p.stderr


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, t)) for t in ("genomic_overlaps", "genomic_scans"))


def parse_label_values(stdout):
    """'LABEL\\tvalue' lines -> (labels, values as str)."""
    labs, vals = [], []
    for line in stdout.decode().splitlines():
        a, b = line.rsplit("\t", 1)
        labs.append(a)
        vals.append(b)
    return labs, vals


# ---------------------------------------------------------------------------------------------
# oracle binding (oracle/oracle.h)
# ---------------------------------------------------------------------------------------------

class _OrcSet(ctypes.Structure):
    _fields_ = [("n_regions", ctypes.c_int64), ("n_intervals", ctypes.c_int64),
                ("chrom", ctypes.c_void_p), ("start", ctypes.c_void_p), ("stop", ctypes.c_void_p),
                ("strand", ctypes.c_void_p), ("weight", ctypes.c_void_p), ("region_offset", ctypes.c_void_p)]


MATCH_GAPS = 1
IGNORE_STRAND = 2


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pack_set(s, weight=None, offsets=None):
    """dict of arrays -> (struct, keepalive).  Arrays are coerced to the ABI dtypes."""
    chrom = np.ascontiguousarray(s["chrom"], dtype=np.int32)
    start = np.ascontiguousarray(s["start"], dtype=np.int32)
    stop = np.ascontiguousarray(s["stop"], dtype=np.int32)
    strand = np.ascontiguousarray(s["strand"], dtype=np.int8)
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.int32)
    off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
    nreg = len(off) - 1 if off is not None else len(chrom)
    st = _OrcSet(nreg, len(chrom), _ptr(chrom), _ptr(start), _ptr(stop), _ptr(strand), _ptr(w), _ptr(off))
    return st, (chrom, start, stop, strand, w, off)


class Oracle:
    """ctypes view of oracle/_ref/liboracle.so (built by `make -C oracle port`)."""

    def __init__(self):
        path = os.path.join(REF_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
        self.lib = ctypes.CDLL(path)
        for fn in ("orc_overlap_count", "orc_overlap_coverage"):
            getattr(self.lib, fn).restype = ctypes.c_int
            getattr(self.lib, fn).argtypes = [ctypes.POINTER(_OrcSet), ctypes.POINTER(_OrcSet), ctypes.c_uint,
                                              ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
        self.lib.orc_scan_counts.restype = ctypes.c_int64
        self.lib.orc_scan_counts.argtypes = [ctypes.POINTER(_OrcSet), ctypes.c_int32, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_int64, ctypes.c_char, ctypes.c_int,
                                             ctypes.c_int64, ctypes.c_int, ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    def _overlap(self, fn, q, idx, flags, qw=None, qoff=None, ioff=None):
        qs, k1 = pack_set(q, qw, qoff)
        is_, k2 = pack_set(idx, None, ioff)
        out = np.zeros(is_.n_regions, dtype=np.uint64)
        err = ctypes.c_int64(-1)
        rc = fn(ctypes.byref(qs), ctypes.byref(is_), flags, _ptr(out), ctypes.byref(err))
        return rc, out, err.value

    def count(self, q, idx, flags=0, **kw):
        return self._overlap(self.lib.orc_overlap_count, q, idx, flags, **kw)

    def coverage(self, q, idx, flags=0, **kw):
        return self._overlap(self.lib.orc_overlap_coverage, q, idx, flags, **kw)

    def scan_counts(self, reads, bound, win_step, win_size, op="1", ignore_strand=False, min_reads=10,
                    emulate_sorted=False, weight=None, offsets=None):
        rs, keep = pack_set(reads, weight, offsets)
        bound = np.ascontiguousarray(bound, dtype=np.int64)
        args = (ctypes.byref(rs), len(bound), _ptr(bound), win_step, win_size, op.encode(), int(ignore_strand),
                min_reads, int(emulate_sorted))
        n = self.lib.orc_scan_counts(*args, 0, None, None, None, None)
        if n < 0:
            return int(n), None
        oc = np.zeros(n, dtype=np.int32); os_ = np.zeros(n, dtype=np.int8)
        ow = np.zeros(n, dtype=np.int64); ov = np.zeros(n, dtype=np.int64)
        n2 = self.lib.orc_scan_counts(*args, n, _ptr(oc), _ptr(os_), _ptr(ow), _ptr(ov))
        assert n2 == n
        return int(n), {"chrom": oc, "strand": os_, "win": ow, "value": ov}


# ---------------------------------------------------------------------------------------------------------------------------------
# A minimal BAM writer (BGZF blocks + BAM records) for the input tests: the reference reads BAM through its vendored samtools,
# the drivers here through host/gt_host.cpp's decoder.
# ---------------------------------------------------------------------------------------------------------------------------------
def write_bam(path, header_text, refs, records, block_bytes=4000):
    """refs: [(name, length)]; records: dicts with qname, flag, tid, pos (0-based), mapq, cigar [(len, op char)], mtid, mpos,
    isize, seq (str or ''), qual (bytes of phred values or None), aux (already encoded bytes)."""
    import struct
    import zlib
    raw = bytearray(b"BAM\x01")
    text = header_text.encode()
    raw += struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, length in refs:
        nm = name.encode() + b"\x00"
        raw += struct.pack("<i", len(nm)) + nm + struct.pack("<i", length)
    ops = "MIDNSHP=X"
    codes = "=ACMGRSVTWYHKDBN"
    for r in records:
        qn = r["qname"].encode() + b"\x00"
        cig = b"".join(struct.pack("<I", (n << 4) | ops.index(op)) for n, op in r["cigar"])
        seq = r.get("seq", "")
        packed = bytearray((len(seq) + 1) // 2)
        for i, ch in enumerate(seq):
            packed[i >> 1] |= codes.index(ch) << (4 if i % 2 == 0 else 0)
        qual = r.get("qual")
        qual = bytes([0xFF] * len(seq)) if qual is None else bytes(qual)
        aux = r.get("aux", b"")
        core = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(qn), r.get("mapq", 0), 4680, len(r["cigar"]), r["flag"], len(seq),
                           r.get("mtid", -1), r.get("mpos", -1), r.get("isize", 0))
        body = core + qn + cig + bytes(packed) + qual + aux
        raw += struct.pack("<i", len(body)) + body

    def bgzf_block(data):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(data) + c.flush()
        bsize = len(comp) + 25
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
                struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))
    with open(path, "wb") as f:
        for lo in range(0, len(raw), block_bytes):
            f.write(bgzf_block(bytes(raw[lo:lo + block_bytes])))
        f.write(bgzf_block(b""))                                        # the end-of-file marker block
