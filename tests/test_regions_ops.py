"""`genomic_regions inv`, `link` and `union` (GenomicRegionSet::RunGlobalInvert / RunGlobalLink / RunUnion, genomic_intervals.cpp:4576-4644, :4397) of this
repo's driver against the reference binary: stdout byte for byte, the fatal cases with the reference's message after the output
it had printed by then.  `inv` is a function of adjacent pairs computed on the host (no GPU needed); `link` gets the linked regions
from the device (gtb_link_regions: prefix-maximum scan) and is a GPU test, as is the ABI-level check against the sequential loop."""
import os
import subprocess

import numpy as np
import pytest

import support

BIN = os.path.join(support.ROOT, "ibm-cbc-genomic-tools_b200", "bin", "genomic_regions")


def run_both(args, stdin=None):
    if not support.have_ref():
        pytest.skip("reference binaries not built (oracle/_ref)")
    want = support.run_ref("genomic_regions", args, stdin=stdin, check=False)
    p = subprocess.run([BIN] + [str(a) for a in args], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return want, (p.returncode, p.stdout, p.stderr)


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("regops")
    rng = np.random.default_rng(5)
    lens = {"chr1": 1000000, "chr10": 600000, "chr2": 800000}
    (d / "genome.bed").write_text("".join("%s\t0\t%d\n" % kv for kv in lens.items()))
    rows = []
    for k in range(3000):
        c = list(lens)[rng.integers(3)]
        s = int(rng.integers(0, lens[c] - 500))
        rows.append((c, "+-"[rng.integers(2)], s, s + int(rng.integers(1, 400)), k))
    by_strand = sorted(rows, key=lambda r: (r[0], r[1], r[2]))
    by_start = sorted(rows, key=lambda r: (r[0], r[2]))
    for name, rs in (("ss", by_strand), ("s", by_start)):
        (d / (name + ".bed")).write_text("".join("%s\t%d\t%d\t%d\t%d\t%s\n" % (c, s, e, k % 13, k % 9, st) for c, st, s, e, k in rs))
        (d / (name + ".reg")).write_text("".join("r%d\t%s %s %d %d\n" % (k, c, st, s + 1, e) for c, st, s, e, k in rs))
        (d / (name + ".gff")).write_text("".join("%s\tsrc\tfeat\t%d\t%d\t0.%d\t%s\t.\tid%d\tnote %d\n" % (c, s + 1, e, k % 7, st, k, k) for c, st, s, e, k in rs))
        (d / (name + ".sam")).write_text("@HD\tVN:1.0\n" + "".join("q%d\t%d\t%s\t%d\t60\t%dM\t=\t7\t-3\t%s\t*\tNM:i:1\n" % (k, 16 if st == "-" else 0, c, s + 1, e - s, "A" * (e - s))
                                                                    for c, st, s, e, k in rs))
    swapped = list(by_strand)
    swapped[1500], swapped[1501] = swapped[1501], swapped[1500]
    if swapped[1500][:2] != swapped[1501][:2]:                              # keep the swap inside one (chromosome, strand) run
        swapped[1502], swapped[1503] = swapped[1503], swapped[1502]
    (d / "unsorted.bed").write_text("".join("%s\t%d\t%d\tr%d\t%d\t%s\n" % (c, s, e, k, k % 9, st) for c, st, s, e, k in swapped))
    (d / "multi.reg").write_text("".join("r%d\t%s %s %d %d\n" % (k, c, st, s + 1, e) for c, st, s, e, k in by_strand[:700]) + "m\tchr1 + 5 9 chr1 + 20 30\n" +
                                 "".join("r%d\t%s %s %d %d\n" % (k, c, st, s + 1, e) for c, st, s, e, k in by_strand[700:]))
    (d / "nochrom.bed").write_text("chr1\t5\t10\ta\t0\t+\nchrZ\t5\t10\tb\t0\t+\n")
    (d / "empty.bed").write_text("")
    return d


@pytest.mark.parametrize("name", ["ss.bed", "ss.reg", "ss.gff", "ss.sam", "unsorted.bed", "multi.reg", "empty.bed"])
def test_inv(files, name):
    want, got = run_both(["inv", "-g", files / "genome.bed", files / name])
    assert got[0] == want[0] and got[1] == want[1], (name, got[2][-200:], want[2][-200:])
    assert got[2] == want[2] or want[0] == 0, (got[2], want[2])
    if name.startswith("ss."):
        assert len(want[1]) > 00


def test_inv_errors(files):
    for args in (["inv", files / "ss.bed"], ["inv", "-g", files / "genome.bed", files / "nochrom.bed"]):
        want, got = run_both(args)
        assert got[0] == want[0] != 0 and got[1] == want[1], (args, got, want)


@pytest.fixture(scope="module")
def union_files(tmp_path_factory):
    """regions of one to six intervals in any order, overlapping, touching, nested, with equal starts -- as REG, BED12 (sorted
    non-overlapping blocks, as the format demands) and SAM (spliced reads)"""
    d = tmp_path_factory.mktemp("union")
    rng = np.random.default_rng(9)
    reg, bed, sam = [], [], []
    for k in range(2000):
        c = "chr%d" % rng.integers(1, 4)
        st = "+-"[rng.integers(2)]
        m = int(rng.integers(1, 7))
        base = int(rng.integers(1, 900000))
        starts = base + rng.integers(0, 400, m)
        if rng.random() < 0.3:
            starts[rng.integers(m)] = starts[0]                             # equal starts
        stops = starts + rng.integers(0, 120, m)
        if rng.random() < 0.3 and m > 1:
            starts[1] = stops[0] + 1                                        # touching: stays apart
            stops[1] = max(stops[1], starts[1])
        reg.append("u%d\t%s\n" % (k, " ".join("%s %s %d %d" % (c, st, a, b) for a, b in zip(starts, stops))))
        sizes = rng.integers(5, 60, m)
        gaps = rng.integers(1, 300, m)
        rel = np.concatenate([[0], np.cumsum(sizes[:-1] + gaps[:-1])])
        bed.append("%s\t%d\t%d\tb%d\t%d\t%s\t%d\t%d\t0,0,255\t%d\t%s\t%s\n" % (c, base, base + rel[-1] + sizes[-1], k, k % 10, st, base, base + 3, m,
                   ",".join(map(str, sizes)), ",".join(map(str, rel))))
        cigar = "".join("%dM%s" % (sizes[i], "%dN" % gaps[i] if i + 1 < m else "") for i in range(m))
        sam.append("q%d\t%d\t%s\t%d\t60\t%s\t=\t7\t-3\t%s\t*\tNM:i:1\n" % (k, 16 if st == "-" else 0, c, base, cigar, "A" * int(sizes.sum())))
    (d / "u.reg").write_text("".join(reg))
    (d / "u.bed").write_text("track name=x\n" + "".join(bed))
    (d / "u.sam").write_text("@HD\tVN:1.0\n" + "".join(sam))
    (d / "bad_strand.reg").write_text("".join(reg[:50]) + "x\tchr1 + 5 9 chr1 - 7 30\n" + "".join(reg[50:60]))
    (d / "bad_chrom.reg").write_text("".join(reg[:5]) + "x\tchr1 + 5 9 chr2 + 7 30 chr1 + 1 2\n" + "".join(reg[5:9]))
    (d / "bad_line.reg").write_text("".join(reg[:7]) + "x\tchr1 + 5\n" + "".join(reg[7:9]))
    (d / "empty.bed").write_text("")
    return d


@pytest.mark.parametrize("name", ["u.reg", "u.bed", "u.sam", "bad_strand.reg", "bad_chrom.reg", "bad_line.reg", "empty.bed"])
def test_union(union_files, files, name):
    want, got = run_both(["union", union_files / name])
    assert got[0] == want[0] and got[1] == want[1], (name, got[1][:300], want[1][:300], got[2][-200:], want[2][-200:])
    assert got[2] == want[2] or want[0] == 0, (got[2], want[2])
    if name.startswith("u."):
        assert want[0] == 0 and len(want[1]) > 0
    for single in ("ss.bed", "ss.gff", "ss.reg") if name == "u.reg" else ():
        want, got = run_both(["union", files / single])
        assert got[0] == want[0] == 0 and got[1] == want[1], single
    if name == "u.reg":                                                    # standard input, and the usage text
        want, got = run_both(["union"], stdin=(union_files / name).read_bytes())
        assert got[0] == want[0] == 0 and got[1] == want[1]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [[], ["-d", "100"], ["-s"], ["-s", "-d", "250"], ["-d", "-20"], ["-s", "-d", "-5"], ["--label-func", "+"], ["-s", "--label-func", "sum"],
                                   ["-d", "50", "--label-func", "min"], ["--label-func", "max", "-d", "1000"]])
def test_link(files, flags):
    stem = "ss" if "-s" in flags else "s"
    # every format under the plain, the -s and the label-function runs; BED and one other format elsewhere (each run is a process
    # with a CUDA context of its own)
    every = flags in ([], ["-s"], ["--label-func", "+"], ["-s", "--label-func", "sum"])
    for ext in ("bed", "reg", "gff", "sam") if every else ("bed", ("reg", "gff", "sam")[len(flags) % 3]):
        want, got = run_both(["link"] + flags + [files / ("%s.%s" % (stem, ext))])
        assert got[0] == want[0] == 0 and got[1] == want[1], (flags, ext, got[1][:200], want[1][:200], got[2][-200:])
        assert len(want[1]) > 0


@pytest.mark.gpu
def test_link_errors(files):
    for args in (["link", "-s", files / "unsorted.bed"], ["link", "-s", files / "multi.reg"], ["link", files / "ss.bed"], ["link", files / "empty.bed"]):
        want, got = run_both(args)
        assert got[0] == want[0] and got[1] == want[1], (args, got[2][-200:], want[2][-200:])
        assert got[2] == want[2] or want[0] == 0, (got[2], want[2])


@pytest.mark.gpu
@pytest.mark.parametrize("n,d", [(1, 0), (2049, 0), (300_000, 0), (2_000_003, 75), (2_000_003, 10_000)])
def test_link_regions_abi(n, d):
    """gtb_link_regions against the sequential loop of RunGlobalLink on a sorted stream of n regions in a few groups"""
    import gtb200
    rng = np.random.default_rng(n + d)
    group = np.sort(rng.integers(0, 5, n)).astype(np.int32)
    start = rng.integers(-1000, 3_000_000, n).astype(np.int64)
    order = np.lexsort((start, group))
    group, start = group[order], start[order]
    stop = start + rng.integers(0, 400, n) * (rng.random(n) < 0.98) + rng.integers(0, 50_000, n) * (rng.random(n) < 0.02)
    ctx = gtb200.Context(0)
    head, lstop = ctx.link_regions(group, start, stop, d)
    ctx.close()
    want_head, want_stop = [], []
    for k in range(n):
        if want_head and group[k] == group[want_head[-1]] and start[k] - want_stop[-1] <= d:
            want_stop[-1] = max(want_stop[-1], int(stop[k]))
        else:
            want_head.append(k); want_stop.append(int(stop[k]))
    assert np.array_equal(head, np.array(want_head, dtype=np.int64)) and np.array_equal(lstop, np.array(want_stop, dtype=np.int32))
