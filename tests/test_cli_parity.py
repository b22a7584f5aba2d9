"""Drop-in parity of the host drivers: `genomic_overlaps count|coverage|density|rpkm` and `genomic_scans counts`
of this repo (C++ host + CUDA engine) against the reference binaries built into oracle/_ref/ -- stdout
byte-for-byte and exit codes, over the flag combinations and input formats of SURVEY.md section 8a'.
Needs a GPU (the drivers have no CPU path) and the reference binaries (they travel with the snapshot)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import randcases
import support

pytestmark = pytest.mark.gpu

BIN = os.path.join(support.ROOT, "ibm-cbc-genomic-tools_b200", "bin")
NAMES = randcases.NAMES


def run_new(tool, args, stdin=None):
    p = subprocess.run([os.path.join(BIN, tool)] + [str(a) for a in args], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout, p.stderr


def both(tool, args, stdin=None):
    if not support.have_ref():
        pytest.skip("reference binaries not built (oracle/_ref)")
    want = support.run_ref(tool, args, stdin=stdin, check=False)
    got = run_new(tool, args, stdin=stdin)
    return want, got


def assert_same(tool, args, stdin=None, nonempty=False):
    want, got = both(tool, args, stdin)
    assert got[0] == want[0], (args, got[2][-300:], want[2][-300:])
    assert got[1] == want[1], (args, got[1][:300], want[1][:300])
    if nonempty:
        assert len(want[1]) > 0, args
    return want


# ------------------------------------------------------------------------------------------------
# the hand-derived known-answer test of SURVEY.md section 8c
# ------------------------------------------------------------------------------------------------
KAT_REF = ("chr1\t100\t200\tgA\t0\t+\n"
           "chr1\t150\t400\tgB\t0\t-\n"
           "chr1\t1000\t2000\tgD\t0\t+\t1000\t2000\t0\t2\t100,100\t0,900\n")
KAT_REF_GFF = "chr2\tsrc\tgene\t11\t20\t.\t.\t.\tgC\n"
KAT_TEST = ("chr1\t190\t210\t5\t0\t+\n"
            "chr1\t199\t200\t7\t0\t-\n"
            "chr1\t200\t201\t2\t0\t+\n"
            "chr3\t1\t5\tx\t0\t+\n"
            "chr2\t19\t25\t3\t0\t+\n"
            "chr1\t1200\t1300\t4\t0\t+\n"
            "chr1\t1050\t1950\t9\t0\t+\t1050\t1950\t0\t2\t10,10\t0,890\n"
            "chr1\t100\t200\tdropped\t0\t+")            # no trailing newline: dropped by the reference


@pytest.fixture(scope="module")
def kat(tmp_path_factory):
    d = tmp_path_factory.mktemp("kat")
    (d / "ref.bed").write_text(KAT_REF)
    (d / "ref.gff").write_text(KAT_REF_GFF)
    (d / "test.bed").write_text(KAT_TEST)
    return d


@pytest.mark.parametrize("op", ["count", "coverage", "density", "rpkm"])
@pytest.mark.parametrize("flags", [[], ["-i"], ["-gaps"], ["-gaps", "-i"], ["--max-label-value", "6"], ["-min", "1"]])
def test_kat(kat, op, flags):
    if op in ("rpkm", "density") and flags == ["-min", "1"]:
        flags = ["-min", "0.05"]
    want = assert_same("genomic_overlaps", [op] + flags + [kat / "ref.bed", kat / "test.bed"], nonempty=True)
    if op == "count" and flags == []:
        assert want[1] == b"gA\t1\ngB\t1\ngD\t1\n"
    if op == "coverage" and flags == []:
        assert want[1] == b"gA\t10\ngB\t1\ngD\t20\n"


def test_kat_gff_strand(kat):
    assert_same("genomic_overlaps", ["count", kat / "ref.gff", kat / "test.bed"])
    assert_same("genomic_overlaps", ["count", "-i", kat / "ref.gff", kat / "test.bed"])
    assert_same("genomic_overlaps", ["density", "-i", kat / "ref.gff", kat / "test.bed"])


def test_stdin_and_legacy_dash(kat):
    data = (kat / "test.bed").read_bytes()
    assert_same("genomic_overlaps", ["-count", kat / "ref.bed"], stdin=data, nonempty=True)


# ------------------------------------------------------------------------------------------------
# randomised files in every supported format
# ------------------------------------------------------------------------------------------------
def write_gff(path, s, names, labels):
    with open(path, "w") as f:
        for k in range(len(s["chrom"])):
            f.write("\t".join([names[s["chrom"][k]], "src", "feat", str(int(s["start"][k])), str(int(s["stop"][k])), ".",
                               chr(int(s["strand"][k])), ".", labels[k]]) + "\n")


def write_bed12(path, s, off, names, labels):
    with open(path, "w") as f:
        for k in range(len(off) - 1):
            lo, hi = off[k], off[k + 1]
            st, en = int(s["start"][lo]) - 1, int(s["stop"][hi - 1])
            sizes = ",".join(str(int(s["stop"][i]) - int(s["start"][i]) + 1) for i in range(lo, hi))
            starts = ",".join(str(int(s["start"][i]) - 1 - st) for i in range(lo, hi))
            f.write("\t".join([names[s["chrom"][lo]], str(st), str(en), labels[k], "0", chr(int(s["strand"][lo])), str(st), str(en), "0",
                               str(hi - lo), sizes, starts]) + "\n")


def gz(path):
    with open(path, "rb") as f, gzip.open(str(path) + ".gz", "wb") as g:
        g.write(f.read())
    return str(path) + ".gz"


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("rand")
    rng = np.random.default_rng(77)
    out = {}
    idx = randcases.rand_single(rng, 400)
    q = randcases.rand_single(rng, 30000, n_chrom=5)
    ilab = ["g%d" % k for k in range(400)]
    qlab = [str(int(v)) for v in rng.integers(-3, 12, size=30000)]
    support.write_bed(d / "idx.bed", idx, NAMES, ilab)
    support.write_bed(d / "idx_space.bed", idx, NAMES, ilab, sep=" ")
    support.write_reg(d / "idx.reg", idx, NAMES, ilab)
    write_gff(d / "idx.gff", idx, NAMES, ilab)
    support.write_bed(d / "q.bed", q, NAMES, qlab)
    support.write_reg(d / "q.reg", q, NAMES, qlab)
    write_gff(d / "q.gff", q, NAMES, qlab)
    out["qgz"] = gz(d / "q.bed")
    out["igz"] = gz(d / "idx.bed")
    midx, moff = randcases.rand_multi(rng, 300)
    mq, mqoff = randcases.rand_multi(rng, 20000)
    write_bed12(d / "midx.bed", midx, moff, NAMES, ["m%d" % k for k in range(300)])
    support.write_reg(d / "midx.reg", midx, NAMES, ["m%d" % k for k in range(300)], offsets=moff)
    write_bed12(d / "mq.bed", mq, mqoff, NAMES, [str(k % 7) for k in range(20000)])
    support.write_reg(d / "mq.reg", mq, NAMES, [str(k % 7) for k in range(20000)], offsets=mqoff)
    # sorted copies for -S (LC_ALL=C sort -k1,1 -k2,2n == strcmp(chrom), start)
    for name in ("idx.bed", "q.bed"):
        env = dict(os.environ, LC_ALL="C")
        with open(d / ("s_" + name), "wb") as f:
            subprocess.check_call(["sort", "-k1,1", "-k2,2n", str(d / name)], stdout=f, env=env)
        with open(d / ("ss_" + name), "wb") as f:
            subprocess.check_call(["sort", "-k1,1", "-k6,6", "-k2,2n", str(d / name)], stdout=f, env=env)
    out["dir"] = d
    return out


# Every invocation is a process that creates a CUDA context (a second or so on the box): the three pairs that differ in what the
# engine sees (single intervals, multi-interval on both sides, multi-interval queries only) run under every operation and flag
# set, the pairs that differ only in how the file is spelt under every other combination.
RANDOM_PAIRS = [("idx.bed", "q.bed"), ("idx_space.bed", "q.reg"), ("idx.reg", "q.gff"), ("idx.gff", "q.bed"),
                ("idx.bed.gz", "q.bed.gz"), ("midx.bed", "mq.bed"), ("midx.reg", "mq.reg"), ("midx.bed", "q.bed"),
                ("idx.bed", "mq.reg")]
RANDOM_FLAGS = [[], ["-i"], ["-gaps"], ["--max-label-value", "5", "-i"]]
RANDOM_CASES = [(op, pair, flags) for o, op in enumerate(["count", "coverage", "density"]) for p, pair in enumerate(RANDOM_PAIRS)
                for f, flags in enumerate(RANDOM_FLAGS) if p in (0, 5, 8) or (o + p + f) % 2 == 0]


@pytest.mark.parametrize("op,pair,flags", RANDOM_CASES)
def test_random_files(files, op, pair, flags):
    d = files["dir"]
    assert_same("genomic_overlaps", [op] + flags + [d / pair[0], d / pair[1]], nonempty=True)


@pytest.mark.parametrize("op", ["count", "coverage"])
def test_sorted_flags(files, op):
    d = files["dir"]
    assert_same("genomic_overlaps", [op, "-S", d / "s_idx.bed", d / "s_q.bed"], nonempty=True)
    assert_same("genomic_overlaps", [op, "-S", "-s", d / "ss_idx.bed", d / "ss_q.bed"], nonempty=True)
    assert_same("genomic_overlaps", [op, "-S", "-i", d / "s_idx.bed", d / "s_q.bed"], nonempty=True)
    assert_same("genomic_overlaps", [op, "-min", "40", d / "idx.bed", d / "q.bed"])
    # toggling semantics of boolean flags (core.cpp:2212): -i -i == no -i
    assert_same("genomic_overlaps", [op, "-i", "-i", d / "idx.bed", d / "q.bed"], nonempty=True)


def test_sorted_engine_admission(tmp_path):
    """-S selects the reference's SortedGenomicRegionSetOverlaps, whose admission differs from the default engine's
    (genomic_intervals.cpp:5807-5937): zero-length queries (BED `chr 5 5`) and queries with stop <= 0 are not fatal but matched by
    the raw predicate, index regions with stop <= 0 are not skipped, and the index file's order and well-formedness are only
    looked at as far as the queries reach."""
    (tmp_path / "idx.bed").write_text("chr1\t-30\t-5\tgneg\t0\t+\nchr1\t0\t100\tg1\t0\t+\nchr1\t60\t200\tg2\t0\t+\nchr2\t10\t20\tg3\t0\t+\n")
    (tmp_path / "q.bed").write_text("chr1\t-10\t-3\tqn\t0\t+\nchr1\t5\t5\tq0\t0\t+\nchr1\t10\t60\tq1\t0\t+\nchr1\t60\t60\tq2\t0\t+\n"
                                    "chr1\t70\t80\tq3\t0\t+\nchr1\t100\t100\tq4\t0\t-\nchr2\t15\t15\tq5\t0\t+\n")
    for op in ("count", "coverage", "density"):
        for flags in ([], ["-i"], ["-gaps"]):
            assert_same("genomic_overlaps", [op, "-S"] + flags + [tmp_path / "idx.bed", tmp_path / "q.bed"], nonempty=True)
    # the default engine dies on the first of these queries, and so do we
    want, got = both("genomic_overlaps", ["count", tmp_path / "idx.bed", tmp_path / "q.bed"])
    assert got[0] == want[0] == 1 and got[1] == want[1] == b"" and got[2] == want[2]
    # an index line out of order BEHIND the last query's reach is never read; one the queries reach is fatal
    (tmp_path / "late.bed").write_text("chr1\t0\t100\tg1\t0\t+\nchr1\t60\t200\tg2\t0\t+\nchr1\t500\t600\tg3\t0\t+\nchr1\t300\t400\tg4\t0\t+\n")
    (tmp_path / "q2.bed").write_text("chr1\t10\t60\tq1\t0\t+\nchr1\t70\t80\tq3\t0\t+\n")
    (tmp_path / "q3.bed").write_text("chr1\t10\t60\tq1\t0\t+\nchr1\t70\t80\tq3\t0\t+\nchr1\t550\t560\tq5\t0\t+\n")
    assert_same("genomic_overlaps", ["count", "-S", tmp_path / "late.bed", tmp_path / "q2.bed"], nonempty=True)
    want, got = both("genomic_overlaps", ["count", "-S", tmp_path / "late.bed", tmp_path / "q3.bed"])
    assert got[0] == want[0] == 1 and got[1] == want[1] == b"" and got[2] == want[2]
    # the same for a malformed (multi-interval, overlapping blocks) index line: read as soon as it becomes the current region, i.e.
    # one region ahead of the queries (and named by its 0-based place in the file: CountIndexOverlaps renumbers the index, :5309)
    (tmp_path / "late.reg").write_text("g1\tchr1 + 1 100\ng2\tchr1 + 61 200\ng3\tchr1 + 501 600\ng4\tchr1 + 801 900 chr1 + 850 950\n")
    assert_same("genomic_overlaps", ["count", "-S", tmp_path / "late.reg", tmp_path / "q2.bed"], nonempty=True)
    want, got = both("genomic_overlaps", ["count", "-S", tmp_path / "late.reg", tmp_path / "q3.bed"])
    assert got[0] == want[0] == 1 and got[1] == want[1] == b"" and got[2] == want[2]


# ------------------------------------------------------------------------------------------------
# error paths: same exit code, nothing on stdout
# ------------------------------------------------------------------------------------------------
def test_errors(files, tmp_path):
    d = files["dir"]
    (tmp_path / "bad_stop.bed").write_text("chr1\t5\t20\ta\t0\t+\nchr1\t-10\t-3\tb\t0\t+\nchr1\t7\t9\tc\t0\t+\n")
    (tmp_path / "bad_order.bed").write_text("chr1\t5\t20\ta\t0\t+\nchr1\t30\t10\tb\t0\t+\n")
    (tmp_path / "bad_strand.bed").write_text("chr1\t5\t20\ta\t0\t?\n")
    (tmp_path / "other_chrom_bad.bed").write_text("chrZ\t30\t10\tb\t0\t+\nchr1\t5\t20\ta\t0\t+\n")
    (tmp_path / "empty.bed").write_text("")
    for bad in ("bad_stop.bed", "bad_order.bed", "bad_strand.bed", "other_chrom_bad.bed", "empty.bed"):
        for op in ("count", "coverage"):
            want, got = both("genomic_overlaps", [op, d / "idx.bed", tmp_path / bad])
            assert got[0] == want[0], (bad, op, got, want)
            assert got[1] == want[1], (bad, op)
    # unsorted input under -S is fatal
    want, got = both("genomic_overlaps", ["count", "-S", d / "s_idx.bed", d / "q.bed"])
    assert got[0] == want[0] == 1 and got[1] == want[1] == b""
    want, got = both("genomic_overlaps", ["count", "-S", "-s", "-i", d / "s_idx.bed", d / "s_q.bed"])
    assert got[0] == want[0] == 1 and got[1] == want[1]
    want, got = both("genomic_overlaps", ["count", "-nonsense", d / "idx.bed", d / "q.bed"])
    assert got[0] == want[0] and got[1] == want[1]
    # invalid index regions are skipped silently: value 0 for count, omitted for density
    (tmp_path / "idx_bad.bed").write_text("chr1\t5\t20\ta\t0\t+\nchr1\t30\t10\tb\t0\t+\nchr1\t-9\t-2\tc\t0\t+\nchr1\t-5\t50\td\t0\t+\n")
    for op in ("count", "coverage", "density"):
        assert_same("genomic_overlaps", [op, tmp_path / "idx_bad.bed", d / "q.bed"], nonempty=True)


def test_usage_text():
    for args in ([], ["count"], ["count", "-h"], ["density", "--help"], ["rpkm"], ["bogus"]):
        want, got = both("genomic_overlaps", args)
        assert got[0] == want[0] and got[1] == want[1], args
    for args in ([], ["counts", "-h"], ["bogus"]):
        want, got = both("genomic_scans", args)
        assert got[0] == want[0] and got[1] == want[1], args


# ------------------------------------------------------------------------------------------------
# genomic_scans counts
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def scanfiles(tmp_path_factory):
    d = tmp_path_factory.mktemp("scan")
    rng = np.random.default_rng(5)
    lens = {"chr1": 30000, "chr10": 21000, "chr2": 777, "chrM": 120, "chrX": 9000}
    with open(d / "genome.bed", "w") as f:
        for n in ("chrX", "chr1", "chr2", "chr10", "chrM"):                    # deliberately not in sorted order
            f.write("%s\t0\t%d\n" % (n, lens[n]))
    n = 40000
    chrom = rng.integers(0, 5, size=n).astype(np.int32)
    L = np.array([lens[x] for x in NAMES])[chrom]
    start = (rng.integers(0, 1 << 30, size=n) % (L + 200) - 50).astype(np.int32)   # some reads start < 1 or beyond the end
    stop = (start + rng.integers(0, 90, size=n)).astype(np.int32)
    ok = stop > 0
    reads = {"chrom": chrom[ok], "start": np.maximum(start[ok], -20), "stop": stop[ok], "strand": np.where(rng.integers(0, 2, size=ok.sum()) == 1, ord("-"), ord("+")).astype(np.int8)}
    names6 = NAMES + ["chrUn"]
    reads["chrom"][::97] = 5                                                    # a chromosome the genome file lacks
    labels = [str(int(v)) for v in rng.integers(0, 9, size=len(reads["chrom"]))]
    support.write_bed(d / "reads.bed", reads, names6, labels)
    gz(d / "reads.bed")
    ref = randcases.rand_single(rng, 60, n_chrom=5, span=20000, max_len=900)
    support.write_bed(d / "ref.bed", ref, NAMES, ["r%d" % k for k in range(60)])
    env = dict(os.environ, LC_ALL="C")
    valid = (reads["start"] >= 1)
    vr = {k: v[valid] for k, v in reads.items()}
    vr["chrom"] = np.where(vr["chrom"] == 5, 0, vr["chrom"]).astype(np.int32)
    support.write_bed(d / "valid.bed", vr, NAMES, [labels[i] for i in np.nonzero(valid)[0]])
    with open(d / "s_valid.bed", "wb") as f:
        subprocess.check_call(["sort", "-k1,1", "-k6,6", "-k2,2n", str(d / "valid.bed")], stdout=f, env=env)
    with open(d / "si_valid.bed", "wb") as f:
        subprocess.check_call(["sort", "-k1,1", "-k2,2n", str(d / "valid.bed")], stdout=f, env=env)
    return d


@pytest.mark.parametrize("flags", [["-w", "200", "-d", "50", "-min", "3"], ["-w", "200", "-d", "50", "-min", "1", "-i"],
                                   ["-min", "5"], ["-w", "100", "-d", "100", "-min", "0"], ["-w", "300", "-d", "100", "-op", "c", "-min", "2"],
                                   ["-w", "200", "-d", "50", "-min", "4", "--max-label-value", "6"],
                                   ["-w", "1000", "-d", "250", "-min", "1"]])
def test_scans_counts(scanfiles, flags):
    d = scanfiles
    assert_same("genomic_scans", ["counts", "-g", d / "genome.bed"] + flags + [d / "reads.bed"], nonempty=True)


def test_scans_variants(scanfiles):
    d = scanfiles
    base = ["counts", "-g", d / "genome.bed", "-w", "200", "-d", "50", "-min", "2"]
    assert_same("genomic_scans", base + [d / "reads.bed.gz"], nonempty=True)
    assert_same("genomic_scans", base, stdin=(d / "reads.bed").read_bytes(), nonempty=True)
    assert_same("genomic_scans", base + ["-r", d / "ref.bed", d / "reads.bed"], nonempty=True)
    assert_same("genomic_scans", base + ["-i", "-r", d / "ref.bed", d / "reads.bed"], nonempty=True)
    assert_same("genomic_scans", base + ["-S", d / "s_valid.bed"], nonempty=True)
    assert_same("genomic_scans", base + ["-S", "-i", d / "si_valid.bed"], nonempty=True)
    # fatal conditions
    for args in (["counts", "-w", "200", "-d", "30", "-g", d / "genome.bed", d / "reads.bed"], ["counts", d / "reads.bed"],
                 ["counts", "-g", d / "genome.bed", "-S", d / "valid.bed"]):
        want, got = both("genomic_scans", args)
        assert got[0] == want[0] and got[1] == want[1], args


# ------------------------------------------------------------------------------------------------
# SAM input (GenomicRegionSAM::Read, genomic_intervals.cpp:2771-2813): POS + CIGAR -> blocks, strand from FLAG 0x10
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def samfiles(files):
    d = files["dir"]
    rng = np.random.default_rng(99)
    lines = ["@HD\tVN:1.0\tSO:unsorted", "@SQ\tSN:chr1\tLN:5000"]
    for k in range(20000):
        c = NAMES[rng.integers(0, 5)]
        pos = int(rng.integers(1, 2000))
        ops, frag = [], 0
        for j in range(int(rng.integers(1, 5))):
            op = "M" if j == 0 else "MMMNNIDSX"[rng.integers(9)]
            ln = int(rng.integers(1, 120))
            ops.append("%d%s" % (ln, op))
            frag += ln if op in "MISX" else 0
        cigar, seq = "".join(ops), "A" * frag
        pick = rng.integers(20)
        if pick == 0:
            cigar, seq = "*", "ACGT" * 9                                 # no CIGAR: <length of SEQ>M
        elif pick == 1:
            seq = "*"
        flag = int(rng.integers(0, 2048))
        if pick == 2:
            c, pos, cigar, flag = "*", 0, "*", 4                         # unmapped
        lines.append("\t".join([str(int(rng.integers(0, 9))), str(flag), c, str(pos), "60", cigar, "*", "0", "0", seq, "*"] +
                               (["NM:i:1"] if k % 3 == 0 else [])))
    (d / "q.sam").write_text("\n".join(lines) + "\n")
    gz(d / "q.sam")
    return d


@pytest.mark.parametrize("op", ["count", "coverage", "density"])
@pytest.mark.parametrize("flags", [[], ["-i"], ["-gaps"], ["--max-label-value", "5"]])
def test_sam_queries(samfiles, op, flags):
    assert_same("genomic_overlaps", [op] + flags + [samfiles / "idx.bed", samfiles / "q.sam"], nonempty=True)


def test_sam_variants(samfiles, scanfiles):
    d = samfiles
    assert_same("genomic_overlaps", ["count", d / "midx.bed", d / "q.sam.gz"], nonempty=True)
    assert_same("genomic_overlaps", ["count", d / "idx.gff"], stdin=(d / "q.sam").read_bytes(), nonempty=True)
    assert_same("genomic_scans", ["counts", "-g", scanfiles / "genome.bed", "-w", "200", "-d", "50", "-min", "2", d / "q.sam"], nonempty=True)
    bad = (d / "q.sam").read_text().splitlines()
    bad[5000] = "r\t0\tchr1\t5\t60\t3M2I\t*\t0\t0\tACGTACGT\t*"          # SEQ length != CIGAR's: fatal with the reference's words
    (d / "bad.sam").write_text("\n".join(bad) + "\n")
    want, got = both("genomic_overlaps", ["count", d / "idx.bed", d / "bad.sam"])
    assert got[0] == want[0] != 0 and got[1] == want[1] and got[2] == want[2]


def test_bam_queries(samfiles, scanfiles):
    """BAM input (FileBufferBAM, core.cpp:371-430): the SAM fixture's alignments written as BAM (BGZF blocks of 4 000 bytes, so
    records straddle blocks), through count / coverage / subset / window counts"""
    import re
    d = samfiles
    refs = [(n, 5000) for n in NAMES[:5]]
    recs = []
    for line in (d / "q.sam").read_text().splitlines():
        if line.startswith("@"):
            continue
        f = line.split("\t")
        cigar = [] if f[5] == "*" else [(int(n), op) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", f[5])]
        recs.append({"qname": f[0], "flag": int(f[1]), "tid": -1 if f[2] == "*" else NAMES.index(f[2]), "pos": int(f[3]) - 1, "mapq": int(f[4]), "cigar": cigar,
                     "mtid": -1, "mpos": -1, "isize": 0, "seq": "" if f[9] == "*" else f[9], "qual": None, "aux": b"NMC\x01" if len(f) > 11 else b""})
    support.write_bam(d / "q.bam", "@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:chr1\tLN:5000\n", refs, recs)
    for args in (["count"], ["coverage", "-i"], ["density", "-gaps"], ["subset"], ["subset", "-inv"]):
        assert_same("genomic_overlaps", args + [d / "idx.bed", d / "q.bam"], nonempty=True)
    assert_same("genomic_scans", ["counts", "-g", scanfiles / "genome.bed", "-w", "200", "-d", "50", "-min", "2", d / "q.bam"], nonempty=True)


# ------------------------------------------------------------------------------------------------
# the per-query operations: subset / overlap (genomic_overlaps.cpp:782-800, :706-739), and genomic_regions gsort
# ------------------------------------------------------------------------------------------------
SUBSET_PAIRS = [("idx.bed", "q.bed"), ("idx_space.bed", "q.reg"), ("idx.reg", "q.gff"), ("idx.gff", "q.bed.gz"),
                ("midx.bed", "mq.bed"), ("midx.reg", "mq.reg"), ("midx.bed", "q.bed"), ("idx.bed", "mq.reg")]
SUBSET_ARGS = [["subset"], ["subset", "-inv"], ["subset", "-i"], ["subset", "-gaps", "-inv"], ["overlap"], ["overlap", "-gaps", "-i"]]
SUBSET_CASES = [(pair, args) for p, pair in enumerate(SUBSET_PAIRS) for a, args in enumerate(SUBSET_ARGS) if p in (0, 4, 7) or (p + a) % 2 == 0]


@pytest.mark.parametrize("pair,args", SUBSET_CASES)
def test_subset_overlap(files, pair, args):
    d = files["dir"]
    assert_same("genomic_overlaps", args + [d / pair[0], d / pair[1]], nonempty=True)


def test_subset_overlap_sam_sorted_and_errors(files, samfiles, tmp_path):
    d = files["dir"]
    # SAM queries: the header lines are echoed first (the test set is opened with hide_header == false), then the records
    for args in (["subset"], ["subset", "-inv"], ["overlap"], ["overlap", "-gaps"]):
        want = assert_same("genomic_overlaps", args + [d / "idx.bed", d / "q.sam"], nonempty=True)
        assert want[1].startswith(b"@HD")
    assert_same("genomic_overlaps", ["subset", d / "midx.bed"], stdin=(d / "q.sam").read_bytes(), nonempty=True)
    # -S
    assert_same("genomic_overlaps", ["subset", "-S", d / "s_idx.bed", d / "s_q.bed"], nonempty=True)
    assert_same("genomic_overlaps", ["overlap", "-S", "-s", d / "ss_idx.bed", d / "ss_q.bed"], nonempty=True)
    assert_same("genomic_overlaps", ["subset", "-S", "-inv", "-i", d / "s_idx.bed", d / "s_q.bed"], nonempty=True)
    # a fatal query in the middle of the file: what was printed before it stays printed, then the message and exit code 1
    lines = (d / "q.bed").read_text().splitlines()
    t = lines[2000].split("\t"); t[1], t[2] = "900", "100"; lines[2000] = "\t".join(t)
    chrom_of_idx = (d / "idx.bed").read_text().split("\t", 1)[0]
    t = lines[2000].split("\t"); t[0] = chrom_of_idx; lines[2000] = "\t".join(t)
    (tmp_path / "bad.bed").write_text("\n".join(lines) + "\n")
    for args in (["subset"], ["overlap"], ["subset", "-inv"]):
        want, got = both("genomic_overlaps", args + [d / "idx.bed", tmp_path / "bad.bed"])
        assert got[0] == want[0] == 1 and got[1] == want[1] and len(want[1]) > 0 and got[2] == want[2], args
    # a malformed line further down: the same
    lines[2000] = "chr1\tnot-a-number"
    (tmp_path / "bad2.bed").write_text("\n".join(lines) + "\n")
    want, got = both("genomic_overlaps", ["subset", d / "idx.bed", tmp_path / "bad2.bed"])
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]


@pytest.fixture(scope="module")
def labelfiles(tmp_path_factory):
    """index regions of every size from 10 bp to 2 Mbp (so that they land in every level of the reference's bin index and
    share bins), queries long enough to walk several bins of several levels; plain, multi-interval and sorted copies"""
    d = tmp_path_factory.mktemp("label")
    rng = np.random.default_rng(4242)
    def regions(n, max_len_log2, tag):
        chrom = rng.integers(0, 3, n)
        length = (2.0 ** rng.uniform(3.3, max_len_log2, n)).astype(np.int64)
        start = rng.integers(0, 6_000_000, n)
        strand = rng.integers(0, 2, n)
        return ["chr%d\t%d\t%d\t%s%d\t0\t%s\n" % (chrom[k] + 1, start[k], start[k] + length[k], tag, k, "+-"[strand[k]]) for k in range(n)]
    idx, q = regions(3000, 21, "g"), regions(6000, 18.5, "q")
    (d / "idx.bed").write_text("".join(idx))
    (d / "q.bed").write_text("".join(q))
    (d / "q.gff").write_text("".join("%s\tsrc\tfeat\t%d\t%s\t.\t%s\t.\t%s\tnote\n" % (t[0], int(t[1]) + 1, t[2], t[5].strip(), t[3]) for t in (l.split("\t") for l in q)))
    (d / "q.sam").write_text("@HD\tVN:1.0\n" + "".join("%s\t%d\t%s\t%d\t60\t%dM\t=\t7\t-3\t*\t*\n" % (t[3], 16 if t[5].strip() == "-" else 0, t[0], int(t[1]) + 1, int(t[2]) - int(t[1]))
                                                        for t in (l.split("\t") for l in q[:1500])))
    for name in ("idx.bed", "q.bed"):
        env = dict(os.environ, LC_ALL="C")
        with open(d / ("s_" + name), "wb") as f:
            subprocess.check_call(["sort", "-s", "-k1,1", "-k2,2n", str(d / name)], stdout=f, env=env)
        with open(d / ("ss_" + name), "wb") as f:
            subprocess.check_call(["sort", "-s", "-k1,1", "-k6,6", "-k2,2n", str(d / name)], stdout=f, env=env)
    return d


@pytest.mark.parametrize("flags", [[], ["-i"], ["-gaps"], ["-B", "8,12"], ["-B", "19"], ["-B", "10,11,12,13,14,15,16"], ["-S"], ["-S", "-i"], ["-S", "-s"]])
def test_overlap_label(labelfiles, files, flags):
    """overlap -label: one line per (query, matching reference region) labelled "query:reference", in the order the reference's
    engine walks the matches -- bin levels, bins, LIFO chains for the default engine (any -B), file order under -S"""
    d = labelfiles
    stem = "ss_" if "-s" in flags else "s_" if "-S" in flags else ""
    want = assert_same("genomic_overlaps", ["overlap", "-label"] + flags + [d / (stem + "idx.bed"), d / (stem + "q.bed")], nonempty=True)
    assert want[1].count(b"\n") > 10000
    if not flags:
        # the walk does tell bin layouts apart on this input
        other = support.run_ref("genomic_overlaps", ["overlap", "-label", "-B", "8,12", d / "idx.bed", d / "q.bed"], check=False)
        assert other[1] != want[1] and sorted(other[1].splitlines()) == sorted(want[1].splitlines())
        for name in ("q.gff", "q.sam"):
            assert_same("genomic_overlaps", ["overlap", "-label", d / "idx.bed", d / name], nonempty=True)
        r = files["dir"]
        for pair in (("midx.bed", "mq.bed"), ("midx.reg", "mq.reg"), ("idx.bed", "mq.reg"), ("midx.bed", "q.bed")):
            assert_same("genomic_overlaps", ["overlap", "-label", r / pair[0], r / pair[1]], nonempty=True)
            assert_same("genomic_overlaps", ["overlap", "-label", "-gaps", "-i", r / pair[0], r / pair[1]], nonempty=True)


@pytest.mark.parametrize("name", ["q.bed", "q.reg", "q.gff", "mq.bed", "mq.reg", "q.sam", "idx_space.bed", "q.bed.gz"])
@pytest.mark.parametrize("flags", [[], ["-s"], ["-b", "5"]])
def test_gsort(files, samfiles, name, flags):
    assert_same("genomic_regions", ["gsort"] + flags + [files["dir"] / name], nonempty=True)


def test_gsort_details(files, tmp_path):
    # intervals inside a region are put in order first (r->Sort()); equal keys keep their input order; stop descending at equal starts
    (tmp_path / "u.reg").write_text("r1\tchr1 + 500 600 chr1 + 100 200\nr2\tchr1 - 150 160\nr3\tchr1 + 100 900\nr4\tchr1 + 100 900\nr5\tchr10 + 1 2\nr6\tchr2 - 7 9\n")
    assert_same("genomic_regions", ["gsort", tmp_path / "u.reg"], nonempty=True)
    assert_same("genomic_regions", ["gsort", "-s", tmp_path / "u.reg"], nonempty=True)
    # (the reference cannot gsort standard input -- "region set must be loaded in memory" -- this driver can: the same bytes as from the file)
    d = files["dir"]
    assert run_new("genomic_regions", ["gsort"], stdin=(d / "q.bed").read_bytes())[1] == run_new("genomic_regions", ["gsort", d / "q.bed"])[1]


# ------------------------------------------------------------------------------------------------
# genomic_scans peaks (PeakFinder, genomic_scans.cpp:215-380): the same windows in the same order, p-values to the printed digits.
# The reference runs here against the oracle build's stand-ins for GSL's tail probabilities (oracle/gsl_stub: defining sums in long
# double); the driver computes them on the device by continued fractions -- two formulations, so the last printed digit may differ.
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def peakfiles(tmp_path_factory):
    d = tmp_path_factory.mktemp("peaks")
    rng = np.random.default_rng(3)
    lens = {"chr1": 200000, "chr10": 90000, "chr2": 150000, "chrS": 130}          # chrS: shorter than a window
    (d / "genome.bed").write_text("".join("%s\t0\t%d\n" % kv for kv in lens.items()))
    names = list(lens)[:3]

    def reads(n, peaks, fn, sort=False):
        rows = []
        for k in range(n):
            c = names[rng.integers(3)]
            if peaks and rng.random() < 0.25:
                s = int(rng.normal([50000, 70000, 30000][rng.integers(3)], 150))
            else:
                s = int(rng.integers(1, lens[c] - 60))
            s = max(1, min(s, lens[c] - 60))
            rows.append((c, "+-"[rng.integers(2)], s, k))
        if sort:
            rows.sort(key=lambda r: (r[0], r[1], r[2]))
        (d / fn).write_text("".join("%s\t%d\t%d\t%d\t0\t%s\n" % (c, s, s + 50, k % 7, st) for c, st, s, k in rows))
    reads(60000, True, "signal.bed"); reads(50000, False, "control.bed")
    reads(30000, True, "s_signal.bed", sort=True); reads(30000, False, "s_control.bed", sort=True)
    return d


def parse_peaks(out):
    rows = []
    for line in out.decode().splitlines():
        p, iv = line.split("\t")
        rows.append((float(p), iv))
    return rows


@pytest.mark.parametrize("flags", [[], ["-cmp"], ["-M", "poisson"], ["-cmp", "-M", "poisson", "-pval", "1e-3"], ["-cmp", "-M", "binomial2"],
                                   ["-cmp", "-M", "cbinomial", "-qval", "0.2"], ["-cmp", "-M", "normal"], ["-norm", "-cmp"], ["-i", "-min", "5"],
                                   ["-w", "500", "-d", "25", "-qval", "0.01"], ["--max-label-value", "4", "-cmp"], ["-pval", "1e-5", "-qval", "0.5"]])
def test_peaks(peakfiles, flags):
    d = peakfiles
    args = ["peaks", "-g", d / "genome.bed"] + (flags if "-w" in flags else ["-w", "200", "-d", "50"] + flags) + [d / "signal.bed", d / "control.bed"]
    want, got = both("genomic_scans", args)
    assert got[0] == want[0] == 0, (got[2][-300:], want[2][-300:])
    w, g = parse_peaks(want[1]), parse_peaks(got[1])
    assert len(w) > 0 and [iv for _, iv in g] == [iv for _, iv in w], (len(g), len(w), flags)
    for (pg, _), (pw, iv) in zip(g, w):
        assert abs(pg - pw) <= 2e-4 * pw + 1e-300, (iv, pg, pw)


def test_peaks_sorted_and_errors(peakfiles):
    d = peakfiles
    base = ["peaks", "-g", d / "genome.bed", "-w", "200", "-d", "50"]
    want, got = both("genomic_scans", base + ["-S", d / "s_signal.bed", d / "s_control.bed"])
    assert got[0] == want[0] == 0
    w, g = parse_peaks(want[1]), parse_peaks(got[1])
    assert len(w) > 0 and [iv for _, iv in g] == [iv for _, iv in w]
    for args in (base + ["-M", "nosuch", d / "signal.bed", d / "control.bed"], base + ["-M", "cbinomial", d / "signal.bed", d / "control.bed"],
                 base + ["-S", d / "signal.bed", d / "control.bed"], ["peaks", "-g", d / "genome.bed", "-w", "200", "-d", "30", d / "signal.bed", d / "control.bed"],
                 ["peaks", "-g", d / "genome.bed"]):
        want, got = both("genomic_scans", args)
        assert got[0] == want[0] != 0 and got[1] == want[1], (args, got, want)
    # without a control file the reference draws Poisson variates from a time-seeded generator: not reproducible; here a seed pins it
    env = dict(os.environ, GT_SEED="7")
    run = lambda: subprocess.run([os.path.join(BIN, "genomic_scans")] + [str(a) for a in base + [d / "signal.bed"]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    a, b = run(), run()
    assert a.returncode == 0 and a.stdout == b.stdout and len(a.stdout) > 0
