"""INTEGRATION.md section 2, proven: the reference's own `genomic_overlaps` driver, compiled unmodified with
GPUGenomicRegionSetOverlaps (oracle/shim/genomic_intervals.h) in place of its two engines and linked against libgtb200
(oracle/Makefile, target `shim`), prints the same bytes as the stock binary.  The reference's parsers, region objects and
printing loops are all still in play; only CountIndexOverlaps / CalcIndexCoverage come from the GPU."""
import os
import subprocess

import numpy as np
import pytest

import randcases
import support

pytestmark = pytest.mark.gpu

SHIM = os.path.join(support.REF_DIR, "genomic_overlaps_gpu")
GENES = os.path.join(support.REF_DIR, "examples", "genes.bed.gz")


def run(exe, args):
    p = subprocess.run([exe] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout, p.stderr


def same(args, expect_rc=0):
    if not (support.have_ref() and os.path.exists(SHIM)):
        pytest.skip("reference binaries / shim not built into oracle/_ref")
    want = run(os.path.join(support.REF_DIR, "genomic_overlaps"), args)
    got = run(SHIM, args)
    assert got[0] == want[0] == expect_rc, (args, got[2][-300:], want[2][-300:])
    assert got[1] == want[1], (args, got[1][:200], want[1][:200])
    if expect_rc == 0:
        assert len(want[1]) > 0
    else:
        assert got[2] == want[2], (got[2], want[2])


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("shim")
    rng = np.random.default_rng(77)
    idx = randcases.rand_single(rng, 500)
    q = randcases.rand_single(rng, 50_000)
    support.write_bed(str(d / "ref.bed"), idx, randcases.NAMES, labels=["g%d" % k for k in range(500)])
    support.write_bed(str(d / "test.bed"), q, randcases.NAMES, labels=[str(1 + k % 9) for k in range(50_000)])
    midx, ioff = randcases.rand_multi(rng, 300)
    mq, qoff = randcases.rand_multi(rng, 20_000)
    support.write_reg(str(d / "ref.reg"), midx, randcases.NAMES, offsets=ioff)
    support.write_reg(str(d / "test.reg"), mq, randcases.NAMES, offsets=qoff)
    if os.path.exists(GENES):
        reads = support.synth_reads(300_000, seed=1, read_len=50, chrom_lens=np.array([197195432]), p_range=(2_999_999, 197195332))
        support.write_bed(str(d / "mm9.bed"), reads, ["chr1"])
    return d


@pytest.mark.parametrize("op", ["count", "coverage", "density", "rpkm"])
@pytest.mark.parametrize("flags", [[], ["-i"], ["-gaps"], ["--max-label-value", "5"], ["-min", "2"]])
def test_shim_single_interval(files, op, flags):
    if op == "rpkm" and flags == ["-min", "2"]:
        flags = []
    same([op] + flags + [files / "ref.bed", files / "test.bed"])


@pytest.mark.parametrize("op", ["count", "coverage", "density"])
@pytest.mark.parametrize("flags", [[], ["-i"], ["-gaps"], ["-gaps", "-i"]])
def test_shim_multi_interval(files, op, flags):
    same([op] + flags + [files / "ref.reg", files / "test.reg"])


def test_shim_shipped_example(files):
    if not os.path.exists(GENES):
        pytest.skip("examples not copied into oracle/_ref")
    for op in ("count", "coverage", "density"):
        same([op, GENES, files / "mm9.bed"])


def test_shim_fatal_query(files, tmp_path):
    # a query with start > stop on an indexed chromosome: the reference dies at that line, and so does the shim
    lines = open(files / "test.bed").read().splitlines()
    t = lines[1234].split("\t")
    t[1], t[2] = "500", "400"
    lines[1234] = "\t".join(t)
    (tmp_path / "bad.bed").write_text("\n".join(lines) + "\n")
    same(["count", files / "ref.bed", tmp_path / "bad.bed"], expect_rc=1)


def test_shim_other_operations_still_the_references(files):
    # operations outside the accelerated path run on the reference's own engine inside the same binary
    same(["overlap", files / "ref.bed", files / "test.bed"])
    same(["subset", files / "ref.bed", files / "test.bed"])
