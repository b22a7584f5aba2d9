"""Several GPUs behind one index in one process (gtb_mgpu_*, include/gtb200.h): query batches cut into one slice per device, the
devices' values added up.  With one GPU in the box the "devices" are several contexts on that GPU -- the slicing, the stream-order
error indices and the summation are the same code; with more GPUs the real devices are used too."""
import numpy as np
import pytest

import randcases
import support

pytestmark = pytest.mark.gpu


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.fixture(scope="module")
def gtb():
    import gtb200
    return gtb200


@pytest.fixture(scope="module")
def oracle():
    return support.Oracle()


@pytest.mark.parametrize("op", ["count", "coverage"])
def test_mgpu_matches_oracle(gtb, oracle, op):
    reads = support.synth_reads(3_000_007, seed=71)
    regions = support.synth_regions(4_000, seed=72)
    fn = oracle.count if op == "count" else oracle.coverage
    gop = gtb.OP_COUNT if op == "count" else gtb.OP_COVERAGE
    for devices in device_lists():
        mg = gtb.MultiGpu(devices)
        for flags in (0, gtb.IGNORE_STRAND):
            rc, want, _ = fn(reads, regions, flags)
            assert rc == 0
            ix = gtb.MultiIndex(mg, regions, gop, flags)
            for rep in range(2):                                          # the second round: reset, then three unequal batches
                ix.reset()
                cuts = [0, len(reads["chrom"])] if rep == 0 else [0, 1_000_001, 1_000_002, len(reads["chrom"])]
                for lo, hi in zip(cuts[:-1], cuts[1:]):
                    ix.add_host({k: v[lo:hi] for k, v in reads.items()})
                assert np.array_equal(ix.finish(), want), (devices, flags, rep)
            ix.close()
        mg.close()


def test_mgpu_multi_interval_weights_and_errors(gtb, oracle):
    rng = np.random.default_rng(73)
    idx, ioff = randcases.rand_multi(rng, 300)
    q, qoff = randcases.rand_multi(rng, 50_000)
    w = rng.integers(1, 9, len(qoff) - 1).astype(np.int32)
    for devices in device_lists():
        mg = gtb.MultiGpu(devices)
        for flags in (0, gtb.MATCH_GAPS):
            rc, want, _ = oracle.coverage(q, idx, flags, qoff=qoff, ioff=ioff, qw=w)
            assert rc == 0
            ix = gtb.MultiIndex(mg, idx, gtb.OP_COVERAGE, flags, roffsets=ioff)
            ix.add_host(q, weight=w, offsets=qoff)
            assert np.array_equal(ix.finish(), want), (devices, flags)
            ix.close()
        # the first offending query in stream order, wherever its slice went
        reads = support.synth_reads(400_000, seed=74)
        regions = support.synth_regions(500, seed=75)
        bad = {k: v.copy() for k, v in reads.items()}
        where = [123_456, 300_000, 399_999]
        for k in where:
            bad["chrom"][k] = regions["chrom"][0]
            bad["stop"][k] = bad["start"][k] - 3
        rc, _, ei = oracle.count(bad, regions, 0)
        assert rc != 0 and ei == where[0]
        ix = gtb.MultiIndex(mg, regions, gtb.OP_COUNT, 0)
        ix.add_host({k: v[:200_000] for k, v in bad.items()})
        ix.add_host({k: v[200_000:] for k, v in bad.items()})
        with pytest.raises(gtb.GtbError) as e:
            ix.finish()
        assert e.value.index == where[0] and e.value.code == rc, (devices, e.value.index, e.value.code)
        ix.close()
        mg.close()


def test_mgpu_pairs_without_offsets(gtb, oracle):
    """regions of two intervals each, handed over without offsets: every device takes a slice of whole pairs"""
    import test_baseline_configs as bc
    pairs, off = bc.synth_pairs(700_001, seed=79)
    regions = support.synth_regions(5_000, seed=80)
    for devices in device_lists():
        mg = gtb.MultiGpu(devices)
        for flags in (0, gtb.MATCH_GAPS):
            rc, want, _ = oracle.coverage(pairs, regions, flags, qoff=off)
            assert rc == 0
            ix = gtb.MultiIndex(mg, regions, gtb.OP_COVERAGE, flags)
            st, keep = gtb.host_set(pairs, per_region=2)
            ix.add_set(st)
            assert np.array_equal(ix.finish(), want), (devices, flags)
            ix.close()
        mg.close()


def test_mgpu_packed_reads(gtb, oracle):
    n = 2_000_003
    reads = support.synth_reads(n, seed=76, read_len=36)
    regions = support.synth_regions(3_000, seed=77)
    meta = (reads["chrom"].astype(np.uint8) | np.where(reads["strand"] == ord("-"), 0x80, 0).astype(np.uint8))
    start = np.ascontiguousarray(reads["start"])
    rc, want, _ = oracle.count(reads, regions, 0)
    assert rc == 0
    for devices in device_lists():
        mg = gtb.MultiGpu(devices)
        ix = gtb.MultiIndex(mg, regions, gtb.OP_COUNT, 0)
        ix.add_packed_ptr(n, start.ctypes.data, meta.ctypes.data, 36)
        assert np.array_equal(ix.finish(), want), devices
        ix.close()
        mg.close()


def test_cli_with_several_devices(tmp_path):
    """GTB_GPUS in the environment of bin/genomic_overlaps: stdout and exit code are those of the single-device run"""
    import os
    import subprocess
    import torch
    rng = np.random.default_rng(78)
    names = ["chr1", "chr2", "chrX"]
    (tmp_path / "ref.bed").write_text("".join("%s\t%d\t%d\tg%d\t0\t%s\n" % (names[rng.integers(3)], s, s + rng.integers(50, 5000), k, "+-"[rng.integers(2)])
                                              for k, s in enumerate(rng.integers(0, 900_000, 800))))
    with open(tmp_path / "reads.bed", "w") as f:
        for k, s in enumerate(rng.integers(0, 1_000_000, 300_000)):
            f.write("%s\t%d\t%d\tr\t0\t%s\n" % (names[rng.integers(3)], s, s + 50, "+-"[rng.integers(2)]))
    exe = os.path.join(support.ROOT, "ibm-cbc-genomic-tools_b200", "bin", "genomic_overlaps")
    lists = ["0,0,0"] + ([str(min(torch.cuda.device_count(), 8))] if torch.cuda.device_count() >= 2 else [])
    for op in ("count", "coverage", "density", "rpkm"):
        base = subprocess.run([exe, op, str(tmp_path / "ref.bed"), str(tmp_path / "reads.bed")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert base.returncode == 0 and len(base.stdout) > 0
        for gpus in lists:
            got = subprocess.run([exe, op, str(tmp_path / "ref.bed"), str(tmp_path / "reads.bed")], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                 env=dict(os.environ, GTB_GPUS=gpus))
            assert got.returncode == 0 and got.stdout == base.stdout, (op, gpus, got.stderr[-300:])
