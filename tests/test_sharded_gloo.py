"""N > 1 host logic of the genome-sharded driver (gtb200.sharded) on CPU: world_size 2 and 3 over gloo.

The planning / ownership / routing / gather code is the product's; the per-shard ENGINE is replaced by the oracle
(test infrastructure) because this container has no GPU -- on the GPU box the same driver runs with the CUDA
engine (tests/test_gpu_parity.py::test_sharded_single_process and bench.py --gpus N).
The sharded result must equal the unsharded oracle run, including reads that straddle every cut point, regions
that span several shards, multi-interval regions, weights and the first-fatal-query behaviour."""
import os
import socket
import sys
import traceback
import zlib

import numpy as np
import pytest

import randcases
import support

ROOT = support.ROOT
sys.path.insert(0, os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python"))


class OracleEngine:
    """Stand-in engine with the CudaShardEngine interface."""

    def __init__(self, regions, roffsets, op, flags):
        self.orc = support.Oracle()
        self.regions, self.roffsets, self.op, self.flags = regions, roffsets, op, flags
        self.batches = []

    def add(self, queries, weight, offsets):
        self.batches.append((queries, weight, offsets))

    def finish(self):
        n_reg = len(self.roffsets) - 1 if self.roffsets is not None else len(self.regions["chrom"])
        total = np.zeros(n_reg, dtype=np.uint64)
        base = 0
        fn = self.orc.count if self.op == 0 else self.orc.coverage
        for q, w, off in self.batches:
            rc, vals, err = fn(q, self.regions, self.flags, qw=w, qoff=off, ioff=self.roffsets)
            if rc != 0:
                return rc, base + err, total
            total += vals
            base += len(off) - 1 if off is not None else len(q["chrom"])
        return 0, -1, total

    def close(self):
        pass


def make_case(name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))      # hash() is salted per process
    if name == "single":
        idx = randcases.rand_single(rng, 500, n_chrom=4, span=4000, max_len=900)
        q = randcases.rand_single(rng, 30000, n_chrom=5, span=4200, max_len=120)
        return dict(idx=idx, ioff=None, q=q, qoff=None, qw=None, op=0, flags=0)
    if name == "single_cov_weighted":
        idx = randcases.rand_grid(rng, 400)
        q = randcases.rand_grid(rng, 20000)
        return dict(idx=idx, ioff=None, q=q, qoff=None, qw=rng.integers(-2, 7, size=20000).astype(np.int32), op=1, flags=2)
    if name == "multi":
        idx, ioff = randcases.rand_multi(rng, 300)
        q, qoff = randcases.rand_multi(rng, 9000)
        return dict(idx=idx, ioff=ioff, q=q, qoff=qoff, qw=None, op=0, flags=0)
    if name == "multi_gaps_cov":
        idx, ioff = randcases.rand_multi(rng, 300)
        q, qoff = randcases.rand_multi(rng, 9000)
        return dict(idx=idx, ioff=ioff, q=q, qoff=qoff, qw=None, op=1, flags=1)
    if name == "fatal":
        idx = randcases.rand_single(rng, 200, n_chrom=3)
        q = randcases.rand_single(rng, 5000, n_chrom=4)
        q["stop"][4100] = q["start"][4100] - 5          # start > stop on an indexed chromosome
        q["chrom"][4100] = 1
        q["stop"][1234] = -3; q["start"][1234] = -9      # stop <= 0, earlier in the stream: this one must be reported
        q["chrom"][1234] = 2
        q["stop"][77] = -3; q["start"][77] = -9          # ... but not this one: chromosome absent from the index
        q["chrom"][77] = 3
        return dict(idx=idx, ioff=None, q=q, qoff=None, qw=None, op=0, flags=0)
    raise KeyError(name)


CASES = ["single", "single_cov_weighted", "multi", "multi_gaps_cov", "fatal"]


def _worker(rank, world, port, out_dir):
    try:
        import torch.distributed as dist
        from gtb200 import GtbError, sharded
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        for name in CASES:
            c = make_case(name)
            for balanced in (False, True):
                hist = None
                if balanced:            # coarse histogram of read starts, as an ingest pass would produce
                    width = 256
                    qc = c["q"]["chrom"] if c["qoff"] is None else c["q"]["chrom"][c["qoff"][:-1]]
                    qs = c["q"]["start"] if c["qoff"] is None else c["q"]["start"][c["qoff"][:-1]]
                    hist = (width, {int(ch): np.bincount(np.clip(qs[qc == ch], 0, None) // width) for ch in np.unique(qc)})
                so = sharded.ShardedOverlap(c["idx"], op=c["op"], flags=c["flags"], roffsets=c["ioff"], read_hist=hist,
                                            engine_factory=lambda r, o, op, fl: OracleEngine(r, o, op, fl))
                # two batches, to exercise the global stream index bookkeeping
                nq = len(c["qoff"]) - 1 if c["qoff"] is not None else len(c["q"]["chrom"])
                cut = nq // 3
                for lo, hi in ((0, cut), (cut, nq)):
                    ids = np.arange(lo, hi)
                    sub, w, off = sharded._take_queries(c["q"], ids, c["qw"], c["qoff"])
                    so.add(sub, w, off)
                try:
                    vals = so.finish()
                    res = ("ok", vals)
                except GtbError as e:
                    res = ("err", np.array([e.code, e.index], dtype=np.int64))
                loads = np.array([sum(len(g) for g in so.global_ids)], dtype=np.int64)
                np.save(os.path.join(out_dir, "%s_%d_%d_%d.npy" % (name, int(balanced), world, rank)), res[1])
                np.save(os.path.join(out_dir, "%s_%d_%d_%d.load.npy" % (name, int(balanced), world, rank)), loads)
                with open(os.path.join(out_dir, "%s_%d_%d_%d.kind" % (name, int(balanced), world, rank)), "w") as f:
                    f.write(res[0])
                so.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        with open(os.path.join(out_dir, "fail_%d_%d.txt" % (world, rank)), "w") as f:
            f.write(traceback.format_exc())
        raise


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(world, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    fails = [f for f in os.listdir(tmp_path) if f.startswith("fail_")]
    assert not fails, open(os.path.join(tmp_path, fails[0])).read()
    orc = support.Oracle()
    for name in CASES:
        c = make_case(name)
        fn = orc.count if c["op"] == 0 else orc.coverage
        rc, want, err = fn(c["q"], c["idx"], c["flags"], qw=c["qw"], qoff=c["qoff"], ioff=c["ioff"])
        nq = len(c["qoff"]) - 1 if c["qoff"] is not None else len(c["q"]["chrom"])
        for balanced in (0, 1):
            loads = []
            for rank in range(world):
                stem = os.path.join(tmp_path, "%s_%d_%d_%d" % (name, balanced, world, rank))
                kind = open(stem + ".kind").read()
                got = np.load(stem + ".npy")
                loads.append(int(np.load(stem + ".load.npy")[0]))
                if rc == 0:
                    assert kind == "ok" and np.array_equal(got, want), (name, balanced, rank)
                else:
                    assert kind == "err" and int(got[0]) == rc and int(got[1]) == err, (name, balanced, rank, got, rc, err)
            # every query is processed at least once overall, and replication stays marginal
            assert sum(loads) >= nq * 0.5 and sum(loads) <= nq * 1.8, (name, loads, nq)
            if balanced and name in ("single", "multi"):
                assert max(loads) <= 1.6 * sum(loads) / world, (name, loads)


def test_plan_properties():
    """Ownership is a partition of the regions; routing covers every overlapping (query, region) pair."""
    from gtb200 import sharded
    rng = np.random.default_rng(3)
    idx = randcases.rand_single(rng, 800, n_chrom=5, span=20000, max_len=5000)
    q = randcases.rand_single(rng, 5000, n_chrom=6, span=21000, max_len=300)
    for g in (1, 2, 4, 8):
        plan = sharded.ShardPlan(idx, g)
        owned = np.concatenate(plan.owned)
        assert np.array_equal(np.sort(owned), np.arange(800))
        routed = [set(plan.route(q, s).tolist()) for s in range(g)]
        ov = (q["chrom"][:, None] == idx["chrom"][None, :]) & (q["start"][:, None] <= idx["stop"][None, :]) & (q["stop"][:, None] >= idx["start"][None, :])
        qi, ri = np.nonzero(ov)
        for a, b in zip(qi.tolist(), ri.tolist()):
            assert a in routed[int(plan.owner[b])]
    # cut points are monotone and the outer shards are open-ended
    plan = sharded.ShardPlan(idx, 4)
    assert np.all(np.diff(plan.lo) >= 0) and plan.lo[0] < 0
    assert len(plan.cut_positions()) == 3


# ------------------------------------------------------------------------------------------------
# window counts over several ranks (gtb200.sharded.ShardedScan): range cuts inside chromosomes, halo of combine - 1 micro-windows
# ------------------------------------------------------------------------------------------------
class OracleScanEngine:
    def __init__(self, bound, win_step, win_size, op, ignore_strand, min_reads):
        self.orc = support.Oracle()
        self.args = (bound, win_step, win_size, op, ignore_strand, min_reads)
        self.batches = []

    def add(self, reads):
        self.batches.append(reads)

    def finish(self):
        reads = {k: np.concatenate([b[k] for b in self.batches]) if self.batches else np.zeros(0, np.int8 if k == "strand" else np.int32)
                 for k in ("chrom", "start", "stop", "strand")}
        n, out = self.orc.scan_counts(reads, *self.args)
        return out

    def close(self):
        pass


SCAN_CASES = [dict(step=25, size=100, op="1", ign=False, mn=2), dict(step=50, size=50, op="c", ign=True, mn=1), dict(step=10, size=70, op="1", ign=False, mn=3)]


def _scan_inputs():
    rng = np.random.default_rng(4242)
    bound = np.array([5000, -1, 1730, 12007, 90], dtype=np.int64)          # a chromosome absent from the genome file, one shorter than a window
    n = 40000
    reads = {"chrom": rng.integers(0, 6, n).astype(np.int32), "start": rng.integers(-20, 12500, n).astype(np.int32),
             "strand": rng.choice(np.array([43, 45], np.int8), n)}
    reads["stop"] = (reads["start"] + rng.integers(-2, 120, n)).astype(np.int32)
    return bound, reads


def _scan_worker(rank, world, port, out_dir):
    try:
        import torch.distributed as dist
        from gtb200 import sharded
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        bound, reads = _scan_inputs()
        for i, c in enumerate(SCAN_CASES):
            ss = sharded.ShardedScan(bound, c["step"], c["size"], c["op"], c["ign"], c["mn"], engine_factory=lambda *a: OracleScanEngine(*a))
            half = len(reads["chrom"]) // 2
            ss.add({k: v[:half] for k, v in reads.items()})
            ss.add({k: v[half:] for k, v in reads.items()})
            got = ss.finish()
            np.savez(os.path.join(out_dir, "scan_%d_%d_%d.npz" % (i, world, rank)), **got)
            ss.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        with open(os.path.join(out_dir, "fail_%d_%d.txt" % (world, rank)), "w") as f:
            f.write(traceback.format_exc())
        raise


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_equals_unsharded(world, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_scan_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    fails = [f for f in os.listdir(tmp_path) if f.startswith("fail_")]
    assert not fails, open(os.path.join(tmp_path, fails[0])).read()
    orc = support.Oracle()
    bound, reads = _scan_inputs()
    for i, c in enumerate(SCAN_CASES):
        n, want = orc.scan_counts(reads, bound, c["step"], c["size"], c["op"], c["ign"], c["mn"])   # the Unsorted scanner, its spurious windows included
        assert n > 50
        for rank in range(world):
            got = np.load(os.path.join(tmp_path, "scan_%d_%d_%d.npz" % (i, world, rank)))
            for k in ("chrom", "strand", "win", "value"):
                assert np.array_equal(got[k], want[k]), (i, world, rank, k)
