"""Pins the C restatement (oracle/oracle.c) against the unmodified reference binaries built into
oracle/_ref/ by oracle/Makefile.  The reference has no tests of its own for this path
(SURVEY.md section 4), so the binaries are the ground truth: default (bin-index) engine, and the
sorted sweep engine as an independent second opinion where its preconditions hold."""
import os

import numpy as np
import pytest

import randcases
import support
from support import IGNORE_STRAND, MATCH_GAPS

pytestmark = pytest.mark.skipif(not support.have_ref(), reason="reference binaries not built (make -C oracle ref)")

FLAG_SETS = [(0, []), (IGNORE_STRAND, ["-i"]), (MATCH_GAPS, ["-gaps"]), (MATCH_GAPS | IGNORE_STRAND, ["-gaps", "-i"])]


@pytest.fixture(scope="module")
def oracle():
    return support.Oracle()


def _ref_values(op, args, ref_file, q_file):
    _, out, _ = support.run_ref("genomic_overlaps", [op] + args + [ref_file, q_file])
    labs, vals = support.parse_label_values(out)
    return labs, np.array([int(v) for v in vals], dtype=np.uint64)


def test_kat_hand_derived(tmp_path, oracle):
    """SURVEY.md section 8c known-answer vectors, checked against binary AND restatement."""
    ref = tmp_path / "ref.bed"; q = tmp_path / "q.bed"
    ref.write_text("chr1\t100\t200\tgA\t0\t+\nchr1\t150\t400\tgB\t0\t-\nchr2\t10\t20\tgC\t0\t.\n"
                   "chr1\t1000\t2000\tgD\t0\t+\t1000\t2000\t0\t2\t100,100\t0,900\n")
    q.write_text("chr1\t190\t210\t5\t0\t+\nchr1\t199\t200\t7\t0\t-\nchr1\t200\t201\t2\t0\t+\nchr3\t1\t5\tx\t0\t+\n"
                 "chr2\t19\t25\t3\t0\t+\nchr1\t1200\t1300\t4\t0\t+\n"
                 "chr1\t1050\t1950\t9\t0\t+\t1050\t1950\t0\t2\t10,10\t0,890\nchr1\t100\t120\t1\t0\t+")  # last line: no newline -> dropped
    idx = {"chrom": [0, 0, 1, 0, 0], "start": [101, 151, 11, 1001, 1901], "stop": [200, 400, 20, 1100, 2000],
           "strand": [ord(c) for c in "+-+++"]}
    ioff = [0, 1, 2, 3, 5]
    qq = {"chrom": [0, 0, 0, 2, 1, 0, 0, 0], "start": [191, 200, 201, 2, 20, 1201, 1051, 1941],
          "stop": [210, 200, 201, 5, 25, 1300, 1060, 1950], "strand": [ord(c) for c in "+-++++++"]}
    qoff = [0, 1, 2, 3, 4, 5, 6, 8]
    labels = [5, 7, 2, 0, 3, 4, 9]          # atol("x") == 0
    expect = {
        ("count", 0, 1): [1, 1, 1, 1], ("count", IGNORE_STRAND, 1): [2, 3, 1, 1], ("count", MATCH_GAPS, 1): [1, 1, 1, 2],
        ("count", 0, 6): [5, 6, 3, 6], ("coverage", 0, 1): [10, 1, 1, 20], ("coverage", MATCH_GAPS | IGNORE_STRAND, 1): [11, 22, 1, 1000],
    }
    for (op, flags, maxlab), want in expect.items():
        args = []
        if flags & IGNORE_STRAND: args.append("-i")
        if flags & MATCH_GAPS: args.append("-gaps")
        if maxlab > 1: args += ["--max-label-value", str(maxlab)]
        _, got_ref = _ref_values(op, args, str(ref), str(q))
        assert got_ref.tolist() == want, (op, flags, maxlab)
        w = None if maxlab <= 1 else np.minimum(maxlab, labels)
        fn = oracle.count if op == "count" else oracle.coverage
        rc, got, _ = fn(qq, idx, flags, qw=w, qoff=qoff, ioff=ioff)
        assert rc == 0 and got.tolist() == want, (op, flags, maxlab)


@pytest.mark.parametrize("seed", range(6))
def test_random_single_interval(tmp_path, oracle, seed):
    rng = np.random.default_rng(100 + seed)
    gen = randcases.rand_grid if seed % 2 else randcases.rand_single
    idx = gen(rng, 60); q = gen(rng, 400)
    q["chrom"][rng.random(len(q["chrom"])) < 0.05] = 4      # chromosome absent from the index
    labels = rng.integers(-3, 12, size=len(q["chrom"]))
    rf = str(tmp_path / "ref.bed"); qf = str(tmp_path / "q.bed")
    support.write_bed(rf, idx, randcases.NAMES); support.write_bed(qf, q, randcases.NAMES, labels=labels)
    for flags, args in FLAG_SETS:
        for op, fn in (("count", oracle.count), ("coverage", oracle.coverage)):
            _, want = _ref_values(op, args, rf, qf)
            rc, got, _ = fn(q, idx, flags)
            assert rc == 0 and np.array_equal(got, want), (op, args)
            _, want = _ref_values(op, args + ["--max-label-value", "7"], rf, qf)
            rc, got, _ = fn(q, idx, flags, qw=np.minimum(7, labels))
            assert rc == 0 and np.array_equal(got, want), (op, args, "weights")


@pytest.mark.parametrize("seed", range(4))
def test_random_multi_interval(tmp_path, oracle, seed):
    rng = np.random.default_rng(200 + seed)
    idx, ioff = randcases.rand_multi(rng, 40); q, qoff = randcases.rand_multi(rng, 300)
    rf = str(tmp_path / "ref.reg"); qf = str(tmp_path / "q.reg")
    support.write_reg(rf, idx, randcases.NAMES, offsets=ioff); support.write_reg(qf, q, randcases.NAMES, offsets=qoff)
    for flags, args in FLAG_SETS:
        for op, fn in (("count", oracle.count), ("coverage", oracle.coverage)):
            _, want = _ref_values(op, args, rf, qf)
            rc, got, _ = fn(q, idx, flags, qoff=qoff, ioff=ioff)
            assert rc == 0 and np.array_equal(got, want), (op, args)


def test_sorted_engine_agrees(tmp_path, oracle):
    """-S (sweep engine) is an independent reference implementation of the same numbers."""
    rng = np.random.default_rng(7)
    idx = randcases.rand_single(rng, 80); q = randcases.rand_single(rng, 500)
    for s in (idx, q):
        order = np.lexsort((s["start"], s["chrom"]))      # NAMES are in strcmp order already
        for k in s: s[k] = s[k][order]
    rf = str(tmp_path / "ref.bed"); qf = str(tmp_path / "q.bed")
    support.write_bed(rf, idx, randcases.NAMES); support.write_bed(qf, q, randcases.NAMES)
    for op, fn in (("count", oracle.count), ("coverage", oracle.coverage)):
        for flags, args in FLAG_SETS[:2]:
            _, a = _ref_values(op, args, rf, qf); _, b = _ref_values(op, ["-S"] + args, rf, qf)
            rc, got, _ = fn(q, idx, flags)
            assert np.array_equal(a, b) and rc == 0 and np.array_equal(got, a)


def test_invalid_query_is_fatal(tmp_path, oracle):
    """start > stop / stop <= 0 on an indexed chromosome aborts with nothing on stdout; the same
    defect on a chromosome the index does not know is ignored (genomic_intervals.cpp:5731,5740-5741)."""
    rf = tmp_path / "ref.bed"; rf.write_text("chr1\t100\t200\tg\t0\t+\n")
    idx = {"chrom": [0], "start": [101], "stop": [200], "strand": [ord("+")]}
    cases = [("chr1\t150\t160\ta\t0\t+\nchr1\t300\t250\tb\t0\t+\n", [0, 0], [151, 301], [160, 250], 3, 1),
             ("chr1\t-10\t0\ta\t0\t+\n", [0], [-9], [0], 2, 0),
             ("chr2\t300\t250\tb\t0\t+\nchr1\t150\t160\ta\t0\t+\n", [1, 0], [301, 151], [250, 160], 0, -1)]
    for text, chrom, start, stop, want_rc, want_idx in cases:
        qf = tmp_path / "q.bed"; qf.write_text(text)
        code, out, err = support.run_ref("genomic_overlaps", ["count", str(rf), str(qf)], check=False)
        q = {"chrom": chrom, "start": start, "stop": stop, "strand": [ord("+")] * len(chrom)}
        rc, got, ei = oracle.count(q, idx, 0)
        assert rc == want_rc and ei == want_idx
        if want_rc:
            assert code == 1 and out == b"" and (b"Line %d:" % (want_idx + 1)) in err
        else:
            assert code == 0 and out == b"g\t1\n" and got.tolist() == [1]


def test_invalid_index_region_scores_zero(tmp_path, oracle):
    rf = tmp_path / "ref.bed"; rf.write_text("chr1\t300\t250\tbad\t0\t+\nchr1\t-50\t0\tneg\t0\t+\nchr1\t-5\t20\tok\t0\t+\n")
    qf = tmp_path / "q.bed"; qf.write_text("chr1\t0\t400\ta\t0\t+\nchr1\t-3\t2\tb\t0\t+\n")
    idx = {"chrom": [0, 0, 0], "start": [301, -49, -4], "stop": [250, 0, 20], "strand": [43] * 3}
    q = {"chrom": [0, 0], "start": [1, -2], "stop": [400, 2], "strand": [43] * 2}
    for op, fn in (("count", oracle.count), ("coverage", oracle.coverage)):
        _, want = _ref_values(op, [], str(rf), str(qf))
        rc, got, _ = fn(q, idx, 0)
        assert rc == 0 and np.array_equal(got, want), op


SCAN_CASES = [dict(w=200, d=50, op="1", i=False, mn=0), dict(w=200, d=50, op="c", i=False, mn=1),
              dict(w=100, d=100, op="1", i=True, mn=0), dict(w=500, d=25, op="1", i=False, mn=10),
              dict(w=150, d=50, op="c", i=True, mn=2)]


@pytest.mark.parametrize("case", SCAN_CASES)
def test_scan_counts(tmp_path, oracle, case):
    rng = np.random.default_rng(case["w"] + case["d"])
    bounds = np.array([5000, 120, 30, 3000, -1], dtype=np.int64)       # chr10 shorter than a window, chr2 tiny, chrX absent
    gf = tmp_path / "genome.bed"
    gf.write_text("".join("%s\t0\t%d\n" % (randcases.NAMES[c], bounds[c]) for c in range(len(bounds)) if bounds[c] >= 0))
    reads = randcases.rand_single(rng, 3000, n_chrom=5, span=5200, max_len=120)
    labels = rng.integers(0, 5, size=3000)
    qf = str(tmp_path / "reads.bed"); support.write_bed(qf, reads, randcases.NAMES, labels=labels)
    for maxlab in (1, 3):
        args = ["counts", "-g", str(gf), "-w", case["w"], "-d", case["d"], "-op", case["op"], "-min", case["mn"]]
        if case["i"]: args.append("-i")
        if maxlab > 1: args += ["--max-label-value", maxlab]
        _, out, _ = support.run_ref("genomic_scans", args + [qf])
        n, got = oracle.scan_counts(reads, bounds, case["d"], case["w"], case["op"], case["i"], case["mn"],
                                    weight=None if maxlab <= 1 else np.minimum(maxlab, labels))
        text = "".join("%d\t%s %s %d %d\n" % (got["value"][k], randcases.NAMES[got["chrom"][k]], chr(got["strand"][k]),
                                               case["d"] * (got["win"][k] - 1) + 1, case["d"] * (got["win"][k] - 1) + case["w"])
                       for k in range(n))
        assert text.encode() == out
