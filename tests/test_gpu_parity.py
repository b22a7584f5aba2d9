"""Parity of the CUDA path (through the C ABI) with the oracle and the committed golden vectors.
Bit-exact: all of this is integer arithmetic."""
import numpy as np
import pytest

import goldens
import randcases
import support

pytestmark = pytest.mark.gpu

ENGINES = {"auto": 0, "rank": 1 << 17, "enumerate": 1 << 16, "bucket": 1 << 19, "direct": 1 << 20}
CELL_KS = ("3", "6", "10")      # three settings of the bucket / direct engines' knobs (cell and bucket widths, kernel forms)


@pytest.fixture(params=CELL_KS)
def cell_k(request, monkeypatch):
    # the bucket engine's knobs ride along: cell width and bucket width (12..16 bits -> several buckets on the toy genomes)
    monkeypatch.setenv("GTB_BUCKET_K", {"3": "2", "6": "5", "10": "8"}[request.param])
    monkeypatch.setenv("GTB_BUCKET_BITS", {"3": "9", "6": "12", "10": "16"}[request.param])
    monkeypatch.setenv("GTB_DIRECT_CELL_BITS", {"3": "4", "6": "7", "10": "12"}[request.param])   # the direct engine's cell width
    if request.param == "6":
        monkeypatch.setenv("GTB_BUCKET_PAGED", "1")       # the paged form of pass 1
    if request.param == "3":
        monkeypatch.setenv("GTB_BUCKET_WC", "1")          # the write-combining form even with few buckets (rings overflow -> general step)
    return request.param



@pytest.fixture(scope="module")
def gtb():
    import gtb200
    return gtb200


@pytest.fixture(scope="module")
def ctx(gtb):
    c = gtb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return support.Oracle()


@pytest.mark.parametrize("engine", list(ENGINES))
@pytest.mark.parametrize("case", goldens.overlap_cases(), ids=lambda c: c["name"])
def test_overlap_golden(ctx, case, engine, cell_k):
    if engine not in ("bucket", "direct") and cell_k != CELL_KS[0]:
        pytest.skip("cell width only matters to the bucket / direct engines")
    multi = case["ioff"] is not None or case["qoff"] is not None
    for (op, flags), want in case["expect"].items():
        if engine == "rank" and multi and op == "count" and not (flags & 1):
            continue        # ranks cannot express "any block pair overlaps"; auto routes these to enumerate
        fn = ctx.overlap_count if op == "count" else ctx.overlap_coverage
        got = fn(case["queries"], case["index"], flags | ENGINES[engine], qweight=case["qw"],
                 qoffsets=case["qoff"], roffsets=case["ioff"])
        assert np.array_equal(got, want), (case["name"], op, flags, engine)


@pytest.mark.parametrize("case", goldens.scan_cases(), ids=lambda c: c["name"])
def test_scan_golden(gtb, ctx, case):
    sc = gtb.Scan(ctx, case["bounds"], case["win_step"], case["win_size"], case["op"], case["ignore_strand"], case["min_reads"])
    sc.add_host(case["reads"], weight=case["rw"])
    n = sc.finish()
    assert n == len(case["expect"]["value"])
    got = sc.fetch(0, n)
    for k in ("chrom", "strand", "win", "value"):
        assert np.array_equal(got[k], case["expect"][k]), (case["name"], k)
    sc.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_vs_oracle(ctx, oracle, seed, cell_k):
    rng = np.random.default_rng(1000 + seed)
    if seed % 2 == 0:
        idx = (randcases.rand_grid if seed % 4 else randcases.rand_single)(rng, 300)
        q = (randcases.rand_grid if seed % 4 else randcases.rand_single)(rng, 20000)
        ioff = qoff = None
    else:
        idx, ioff = randcases.rand_multi(rng, 200)
        q, qoff = randcases.rand_multi(rng, 8000)
    w = rng.integers(-3, 9, size=(len(qoff) - 1) if qoff is not None else len(q["chrom"])).astype(np.int32)
    for flags in range(4):
        for weights in (None, w):
            rc, want, _ = oracle.count(q, idx, flags, qw=weights, qoff=qoff, ioff=ioff)
            got = ctx.overlap_count(q, idx, flags, qweight=weights, qoffsets=qoff, roffsets=ioff)
            assert rc == 0 and np.array_equal(got, want), ("count", flags, seed)
            rc, want, _ = oracle.coverage(q, idx, flags, qw=weights, qoff=qoff, ioff=ioff)
            got = ctx.overlap_coverage(q, idx, flags, qweight=weights, qoffsets=qoff, roffsets=ioff)
            assert rc == 0 and np.array_equal(got, want), ("coverage", flags, seed)
            got = ctx.overlap_coverage(q, idx, flags | ENGINES["enumerate"], qweight=weights, qoffsets=qoff, roffsets=ioff)
            assert np.array_equal(got, want), ("coverage/enumerate", flags, seed)


@pytest.mark.parametrize("seed", range(4))
def test_direct_engine_vs_oracle(gtb, ctx, oracle, seed, cell_k):
    """The one-pass DIRECT engine on batches large enough for its tiled path (tens of thousands of queries on toy genomes, so
    that cells hold 0, 1..3 and more than three evaluation points at the three cell widths), with everything its fast path
    must hand to the general one mixed in: strands other than +/-, chromosomes the index has never seen, reads reaching past
    the last evaluation point, long reads spanning many cells -- and, in a second run, invalid queries whose error must be
    the first in stream order."""
    rng = np.random.default_rng(4200 + seed)
    gen = randcases.rand_grid if seed % 2 else randcases.rand_single
    idx = gen(rng, 400 if seed < 2 else 40)
    n = 4096 * 9 + 1234
    q = gen(rng, n, strands="+-+-+-." if seed % 2 else "+-")
    q["chrom"][rng.random(n) < 0.02] = 7                                 # unknown chromosome
    far = rng.random(n) < 0.05
    q["start"][far] += 5000; q["stop"][far] += 5000                      # beyond every evaluation point
    wide = rng.random(n) < 0.05
    q["stop"][wide] += rng.integers(1, 6000, size=int(wide.sum())).astype(np.int32)
    for flags in range(4):
        rc, want, _ = oracle.count(q, idx, flags)
        assert rc == 0
        got = ctx.overlap_count(q, idx, flags | ENGINES["direct"])
        assert np.array_equal(got, want), (flags, seed)
        assert np.array_equal(ctx.overlap_count(q, idx, flags | ENGINES["bucket"]), want)
        # coverage: reads of every length (each one off the batch's common length is a reduction of its own) ...
        rc, want, _ = oracle.coverage(q, idx, flags)
        assert rc == 0
        assert np.array_equal(ctx.overlap_coverage(q, idx, flags | ENGINES["direct"]), want), ("coverage", flags, seed)
        # ... and mostly one length (the byte counters count, the commit multiplies)
        fixed = {k: v.copy() for k, v in q.items()}
        keep = rng.random(n) < 0.1
        fixed["stop"] = np.where(keep, fixed["stop"], fixed["start"] + 36).astype(np.int32)
        rc, want, _ = oracle.coverage(fixed, idx, flags)
        assert rc == 0
        assert np.array_equal(ctx.overlap_coverage(fixed, idx, flags | ENGINES["direct"]), want), ("coverage/fixed", flags, seed)
    # position-sorted input (by chromosome and start with the strands mixed, and by chromosome, strand and start): whole warps
    # between two evaluation points leave as one reduction per strand before any lookup
    for keys in ((q["start"], q["chrom"]), (q["start"], q["strand"], q["chrom"])):
        order = np.lexsort(keys)
        srt = {k: np.ascontiguousarray(v[order]) for k, v in q.items()}
        fsrt = {k: v.copy() for k, v in srt.items()}
        fsrt["stop"] = (fsrt["start"] + 36).astype(np.int32)
        for flags in (0, 1):
            rc, want, _ = oracle.count(srt, idx, flags)
            assert rc == 0 and np.array_equal(ctx.overlap_count(srt, idx, flags | ENGINES["direct"]), want), ("sorted", flags, seed)
            rc, want, _ = oracle.coverage(fsrt, idx, flags)
            assert rc == 0 and np.array_equal(ctx.overlap_coverage(fsrt, idx, flags | ENGINES["direct"]), want), ("sorted coverage", flags, seed)
    bad = {k: v.copy() for k, v in q.items()}
    where = sorted(rng.choice(n, size=3, replace=False).tolist())
    bad["stop"][where[0]] = bad["start"][where[0]] - 1                   # start > stop
    bad["stop"][where[1]] = 0; bad["start"][where[1]] = -3               # stop <= 0
    bad["stop"][where[2]] = bad["start"][where[2]] - 5
    bad["chrom"][where] = idx["chrom"][0]                                # on an indexed chromosome: fatal (:5731-5741)
    rc, _, ei = oracle.count(bad, idx, 0)
    assert rc != 0
    with pytest.raises(gtb.GtbError) as e:
        ctx.overlap_count(bad, idx, ENGINES["direct"])
    assert (e.value.code, e.value.index) == (rc, ei)


@pytest.mark.parametrize("seed", [0, 1])
def test_direct_engine_weighted(gtb, ctx, oracle, seed, cell_k):
    """--max-label-value weights through the DIRECT engine: weights of 1..127 ride the byte counters (a counter passes 128 after
    two or three of the larger ones: spills all the time), everything else -- 0, negative, 128 and up, near 2^31 -- leaves as a
    reduction of its own; count and coverage (one common read length, and every length), against the oracle."""
    rng = np.random.default_rng(5200 + seed)
    gen = randcases.rand_grid if seed % 2 else randcases.rand_single
    idx = gen(rng, 300)
    n = 4096 * 7 + 77
    q = gen(rng, n, strands="+-+-+-." if seed % 2 else "+-")
    w = rng.integers(1, 128, n).astype(np.int32)
    odd = rng.random(n) < 0.1
    w[odd] = rng.choice(np.array([0, -3, 128, 255, 70000, 2_000_000_000], dtype=np.int64), size=int(odd.sum())).astype(np.int32)
    fixed = {k: v.copy() for k, v in q.items()}
    fixed["stop"] = np.where(rng.random(n) < 0.1, fixed["stop"], fixed["start"] + 36).astype(np.int32)
    for flags in range(4):
        rc, want, _ = oracle.count(q, idx, flags, qw=w)
        assert rc == 0
        assert np.array_equal(ctx.overlap_count(q, idx, flags | ENGINES["direct"], qweight=w), want), (flags, seed)
        for reads in (q, fixed):
            rc, want, _ = oracle.coverage(reads, idx, flags, qw=w)
            assert rc == 0
            assert np.array_equal(ctx.overlap_coverage(reads, idx, flags | ENGINES["direct"], qweight=w), want), ("coverage", flags, seed)
    # piled-up weighted reads: byte counters overflow before their spills land -> the batch is replayed with its weights
    piled = {k: v.copy() for k, v in q.items()}
    piled["chrom"][:] = q["chrom"][0]; piled["strand"][:] = q["strand"][0]
    piled["start"][:] = q["start"][0]; piled["stop"][:] = q["stop"][0]
    rc, want, _ = oracle.count(piled, idx, 0, qw=w)
    assert rc == 0 and np.array_equal(ctx.overlap_count(piled, idx, ENGINES["direct"], qweight=w), want)


def test_direct_engine_counter_overflow_is_replayed(gtb, ctx, oracle):
    """Byte counters in shared memory: heavy skew (most reads on a handful of loci, in random order) makes some counter take
    more than 127 adds before its spill lands -- or it does not, depending on timing.  Either way the result is exact: an
    overflow discards the batch and replays it through the general step, and the index then leaves the engine."""
    import torch
    n = 3_000_000
    reads = support.synth_reads(n, seed=31)
    regions = support.synth_regions(3_000, seed=32)
    rng = np.random.default_rng(7)
    hot = rng.random(n) < 0.9
    loci = rng.integers(0, 4, size=n)
    reads["chrom"][hot] = 2
    reads["start"][hot] = (1_000_000 + loci[hot] * 37).astype(np.int32)
    reads["stop"] = (reads["start"] + 49).astype(np.int32)
    dev = {k: torch.from_numpy(v).cuda() for k, v in reads.items()}
    for flags in (0, gtb.IGNORE_STRAND):
        for op, fn in ((gtb.OP_COUNT, oracle.count), (gtb.OP_COVERAGE, oracle.coverage)):
            rc, want, _ = fn(reads, regions, flags)
            assert rc == 0
            ix = gtb.Index(ctx, regions, op, flags | ENGINES["direct"])
            for rep in range(3):
                ix.reset()
                ix.add_device(dev)
                assert np.array_equal(ix.finish(), want), (op, flags, rep)
            ix.close()


def test_negative_and_degenerate_coordinates(ctx, oracle, cell_k):
    """Index regions with start<=0, invalid index regions (value 0), queries starting at <= 0."""
    idx = {"chrom": [0, 0, 0, 0, 1], "start": [301, -49, -4, 5, 1], "stop": [250, 0, 20, 5, 2147483647], "strand": [43] * 5}
    q = {"chrom": [0, 0, 0, 1, 1], "start": [1, -2, 5, 2147483000, 1], "stop": [400, 2, 5, 2147483647, 1], "strand": [43] * 5}
    for flags in (0, 1, 2, 3):
        for fn_o, fn_g in ((oracle.count, ctx.overlap_count), (oracle.coverage, ctx.overlap_coverage)):
            rc, want, _ = fn_o(q, idx, flags)
            assert rc == 0 and np.array_equal(fn_g(q, idx, flags), want)
            assert np.array_equal(fn_g(q, idx, flags | ENGINES["enumerate"]), want)
            assert np.array_equal(fn_g(q, idx, flags | ENGINES["rank"]), want)


def test_fatal_conditions(gtb, ctx, oracle):
    idx = {"chrom": [0], "start": [101], "stop": [200], "strand": [43]}
    cases = [({"chrom": [0, 0, 0], "start": [151, 301, 7], "stop": [160, 250, 3], "strand": [43] * 3}, 3, 1),
             ({"chrom": [0, 0], "start": [150, -9], "stop": [160, 0], "strand": [43] * 2}, 2, 1),
             ({"chrom": [1, 0], "start": [301, 151], "stop": [250, 160], "strand": [43] * 2}, 0, -1)]
    for q, code, where in cases:
        rc, _, ei = oracle.count(q, idx, 0)
        assert (rc, ei) == (code, where)
        for eng in ENGINES.values():
            if code == 0:
                assert ctx.overlap_count(q, idx, eng).tolist() == [1]
            else:
                with pytest.raises(gtb.GtbError) as e:
                    ctx.overlap_count(q, idx, eng)
                assert (e.value.code, e.value.index) == (code, where)
    # malformed regions: blocks out of order / overlapping / on different strands
    bad = {"chrom": [0, 0], "start": [50, 10], "stop": [60, 20], "strand": [43, 43]}
    off = [0, 2]
    rc, _, ei = oracle.count(bad, idx, 0, qoff=off)
    assert (rc, ei) == (4, 0)
    with pytest.raises(gtb.GtbError) as e:
        ctx.overlap_count(bad, idx, 0, qoffsets=off)
    assert (e.value.code, e.value.index) == (4, 0)
    rc, _, ei = oracle.count(idx, bad, 0, ioff=off)
    assert (rc, ei) == (5, 0)
    with pytest.raises(gtb.GtbError) as e:
        ctx.overlap_count(idx, bad, 0, roffsets=off)
    assert (e.value.code, e.value.index) == (5, 0)


def test_set_shape_arguments(gtb, ctx):
    """a query set without offsets is one interval per region, or k >= 2 per region: anything else is an argument error"""
    import ctypes
    idx = {"chrom": [0], "start": [1], "stop": [10], "strand": [43]}
    ix = gtb.Index(ctx, idx, gtb.OP_COUNT, 0)
    q = {"chrom": [0] * 6, "start": [1, 2, 3, 4, 5, 6], "stop": [3, 4, 5, 6, 7, 8], "strand": [43] * 6}
    st, keep = gtb.host_set(q)
    for n_regions, ok in ((6, True), (3, True), (2, True), (4, False), (5, False), (7, False)):
        st.n_regions = n_regions
        rc = gtb.lib().gtb_index_add_queries(ix._h, ctypes.byref(st), gtb.MEM_HOST)
        assert (rc == 0) == ok, (n_regions, rc)
    ix.close()


def test_empty_inputs(ctx):
    empty = {"chrom": [], "start": [], "stop": [], "strand": []}
    idx = {"chrom": [0], "start": [1], "stop": [10], "strand": [43]}
    assert ctx.overlap_count(empty, idx).tolist() == [0]
    assert ctx.overlap_count(idx, empty).tolist() == []
    assert ctx.overlap_coverage(empty, empty).tolist() == []


def test_streaming_batches_equal_one_shot(gtb, ctx, oracle):
    reads = support.synth_reads(300_000, seed=21)
    regions = support.synth_regions(3_000, seed=22)
    rc, want, _ = oracle.coverage(reads, regions, 0)
    ix = gtb.Index(ctx, regions, gtb.OP_COVERAGE, 0)
    for lo in range(0, 300_000, 70_001):
        ix.add_host({k: v[lo:lo + 70_001] for k, v in reads.items()})
    assert np.array_equal(ix.finish(), want)
    assert np.array_equal(ix.finish(), want), "finish must be repeatable"
    ix.reset()
    ix.add_host(reads)
    assert np.array_equal(ix.finish(), want)
    ix.close()


def test_hg19_shaped_medium(gtb, ctx, oracle):
    """2 M hg19-shaped 50-bp reads vs 5 000 gene-like regions, every flag combination, plus
    device-resident inputs generated by the library's own generator (== the numpy generator)."""
    import torch
    n = 2_000_000
    reads = support.synth_reads(n, seed=2)
    regions = support.synth_regions(5_000, seed=3)
    dev = {"chrom": torch.empty(n, dtype=torch.int32, device="cuda"), "start": torch.empty(n, dtype=torch.int32, device="cuda"),
           "stop": torch.empty(n, dtype=torch.int32, device="cuda"), "strand": torch.empty(n, dtype=torch.int8, device="cuda")}
    ctx.synth_reads(2, 0, n, 50, support.HG19_LENS, dev)
    for k in dev:
        assert np.array_equal(dev[k].cpu().numpy(), reads[k]), k
    for flags in (0, gtb.IGNORE_STRAND):
        for op, fn in ((gtb.OP_COUNT, oracle.count), (gtb.OP_COVERAGE, oracle.coverage)):
            rc, want, _ = fn(reads, regions, flags)
            assert rc == 0
            ix = gtb.Index(ctx, regions, op, flags)
            ix.add_device(dev)
            assert np.array_equal(ix.finish(), want), (op, flags)
            ix.close()


def test_sorted_and_clustered_input(gtb, ctx, oracle):
    """Position-sorted reads (whole warps in one bucket: the direct-block path of pass 1, the one-atomic-per-warp path of
    pass 2) and reads piled onto a few loci (rings overflow -> general step, then the paged form)."""
    import torch
    n = 1_500_000
    reads = support.synth_reads(n, seed=5)
    regions = support.synth_regions(6_000, seed=6)
    order = np.lexsort((reads["start"], reads["strand"], reads["chrom"]))
    srt = {k: np.ascontiguousarray(v[order]) for k, v in reads.items()}
    piled = {k: v.copy() for k, v in reads.items()}
    piled["chrom"][:] = 3
    piled["start"] = (piled["start"] % 300_000 + 1_000_000).astype(np.int32)
    piled["stop"] = (piled["start"] + 49).astype(np.int32)
    for q in (srt, piled):
        dev = {k: torch.from_numpy(v).cuda() for k, v in q.items()}
        for flags in (0, gtb.IGNORE_STRAND):
            for op, fn in ((gtb.OP_COUNT, oracle.count), (gtb.OP_COVERAGE, oracle.coverage)):
                rc, want, _ = fn(q, regions, flags)
                assert rc == 0
                ix = gtb.Index(ctx, regions, op, flags)
                for rep in range(2):                      # the second pass runs after the skew watchdog has had its say
                    ix.reset()
                    ix.add_device(dev)
                    assert np.array_equal(ix.finish(), want), (op, flags, rep)
                ix.close()


def test_dense_index(gtb, ctx, oracle):
    """An index dense enough (hundreds of thousands of short regions) that one bucket's points fill most of an SM's shared
    memory: the directory stops refining, pass 2 runs one wide CTA per SM.  Count and coverage, default engine."""
    import torch
    n = 1_000_000
    reads = support.synth_reads(n, seed=51)
    regions = support.synth_regions(250_000, seed=52, min_len=200, max_len=5_000)
    dev = {k: torch.from_numpy(v).cuda() for k, v in reads.items()}
    for flags in (0, gtb.IGNORE_STRAND):
        for op, fn in ((gtb.OP_COUNT, oracle.count), (gtb.OP_COVERAGE, oracle.coverage)):
            rc, want, _ = fn(reads, regions, flags)
            assert rc == 0
            ix = gtb.Index(ctx, regions, op, flags)
            ix.add_device(dev)
            assert np.array_equal(ix.finish(), want), (op, flags)
            ix.close()


@pytest.fixture(params=("bucketed", "direct"))
def scan_form(request, monkeypatch):
    """Both ways of building the micro-window histogram: the write-combining partition + shared-memory counters (forced for
    batches of any size), and one global reduction per read."""
    if request.param == "bucketed":
        monkeypatch.setenv("GTB_SCAN_BUCKET_MIN", "1")
    else:
        monkeypatch.setenv("GTB_SCAN_DIRECT", "1")
    return request.param


def test_scan_hg19_vs_oracle(gtb, ctx, oracle, scan_form):
    reads = support.synth_reads(1_000_000, seed=4)
    bound = support.HG19_LENS.copy()
    bound[5] = -1                                    # one chromosome missing from the genome file
    for (w, d, op, ign, mn) in ((200, 50, "1", False, 2), (500, 25, "c", True, 3), (1000, 1000, "1", False, 1)):
        n, want = oracle.scan_counts(reads, bound, d, w, op, ign, mn)
        sc = gtb.Scan(ctx, bound, d, w, op, ign, mn)
        for lo in range(0, 1_000_000, 400_000):
            sc.add_host({k: v[lo:lo + 400_000] for k, v in reads.items()})
        assert sc.finish() == n
        got = sc.fetch(0, n)
        for k in want:
            assert np.array_equal(got[k], want[k]), (w, d, op, k)
        sc.close()


def _scan_check(gtb, ctx, oracle, reads, bound, d, w, op, ign, mn, batches=1, offsets=None, weight=None, device=False):
    import torch
    n, want = oracle.scan_counts(reads, bound, d, w, op, ign, mn, weight=weight, offsets=offsets)
    sc = gtb.Scan(ctx, bound, d, w, op, ign, mn)
    for rep in range(2):                              # the second round checks reset + accumulation into a used table
        if rep:
            sc.reset()
        if device:
            dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in reads.items()}
            sc.add_device(dev)
        elif batches == 1:
            sc.add_host(reads, weight=weight, offsets=offsets)
        else:
            total = len(reads["chrom"])
            step = (total + batches - 1) // batches
            for lo in range(0, total, step):
                sc.add_host({k: v[lo:lo + step] for k, v in reads.items()})
        assert sc.finish() == n, (d, w, op, ign, mn, rep)
        got = sc.fetch(0, n)
        for k in want:
            assert np.array_equal(got[k], want[k]), (d, w, op, ign, mn, k, rep)
    sc.close()


def test_scan_bucketed_forms(gtb, ctx, oracle, scan_form):
    """The bucketed histogram on everything that bends it: buckets walked by 1, 4 and 8 CTAs, position-sorted reads (whole
    warps in one bucket, runs of equal elements), reads piled onto one micro-window (16-bit counters spill many times, rings
    overflow), reads that count nowhere, odd strand bytes, negative starts with the centre operator, device-resident input."""
    n = 1_200_000
    reads = support.synth_reads(n, seed=41)
    bound = support.HG19_LENS.copy()
    bound[3] = -1
    # sprinkle reads the scanner must ignore, and strands other than '+'/'-'
    rng = np.random.default_rng(7)
    bad = rng.choice(n, 5_000, replace=False)
    reads["stop"][bad[:1000]] = reads["start"][bad[:1000]] - 1            # start > stop
    reads["start"][bad[1000:2000]] = -40; reads["stop"][bad[1000:2000]] = -3   # stop <= 0
    reads["chrom"][bad[2000:3000]] = 99                                   # unknown chromosome
    reads["start"][bad[3000:4000]] = -20; reads["stop"][bad[3000:4000]] = 60   # start < 1: counts only under -op c
    reads["strand"][bad[4000:5000]] = ord(".")
    for (w, d, op, ign, mn) in ((200, 50, "1", False, 1), (100, 25, "c", False, 2), (400, 200, "1", True, 1), (5000, 1000, "c", False, 3)):
        _scan_check(gtb, ctx, oracle, reads, bound, d, w, op, ign, mn, batches=3)
    order = np.lexsort((reads["start"], reads["strand"], reads["chrom"]))
    srt = {k: np.ascontiguousarray(v[order]) for k, v in reads.items()}
    _scan_check(gtb, ctx, oracle, srt, bound, 50, 200, "1", False, 2, device=True)
    piled = {k: v.copy() for k, v in support.synth_reads(n, seed=42).items()}
    piled["chrom"][:] = 7
    piled["start"] = (piled["start"] % 180 + 5_000_000).astype(np.int32)   # four micro-windows of 50 bp
    piled["stop"] = (piled["start"] + 49).astype(np.int32)
    _scan_check(gtb, ctx, oracle, piled, bound, 50, 200, "1", False, 10)
    psrt = {k: np.ascontiguousarray(v[np.argsort(piled["start"], kind="stable")]) for k, v in piled.items()}
    _scan_check(gtb, ctx, oracle, psrt, bound, 50, 200, "1", True, 10, device=True)


def test_scan_multi_interval_and_weights(gtb, ctx, oracle, scan_form):
    """Multi-interval reads count once per block (bucketed when unweighted); weighted reads take the direct form."""
    n = 300_000
    reads = support.synth_reads(n, seed=43)
    bound = support.HG19_LENS.copy()
    offsets = np.arange(0, n + 1, 3, dtype=np.int64)                      # regions of three blocks
    _scan_check(gtb, ctx, oracle, reads, bound, 50, 200, "1", False, 2, offsets=offsets)
    weight = (np.arange(len(offsets) - 1) % 5).astype(np.int32)
    _scan_check(gtb, ctx, oracle, reads, bound, 50, 200, "1", False, 2, offsets=offsets, weight=weight)


@pytest.mark.parametrize("shards", [2, 4])
@pytest.mark.parametrize("op", ["count", "coverage"])
def test_sharded_plan_cuda(gtb, ctx, oracle, shards, op):
    """The genome-sharded decomposition with the CUDA engine, the G shards run one after the other on this GPU:
    ownership + routing + per-shard Index + assembly by owner must equal the unsharded oracle."""
    from gtb200 import sharded
    reads = support.synth_reads(400_000, seed=21)
    regions = support.synth_regions(3_000, seed=22)
    plan = sharded.ShardPlan(regions, shards, chrom_extent=support.HG19_LENS)
    fn = oracle.count if op == "count" else oracle.coverage
    rc, want, _ = fn(reads, regions, 0)
    assert rc == 0
    result = np.zeros(len(regions["chrom"]), dtype=np.uint64)
    routed = 0
    for s in range(shards):
        sub, _ = plan.subset(regions, s)
        ids = plan.route(reads, s)
        routed += len(ids)
        q, _, _ = sharded._take_queries(reads, ids)
        ix = gtb.Index(ctx, sub, gtb.OP_COUNT if op == "count" else gtb.OP_COVERAGE, 0)
        ix.add_host(q)
        result[plan.owned[s]] = ix.finish()
        ix.close()
    assert np.array_equal(result, want)
    assert routed <= 1.05 * 400_000           # replication across the cut points stays marginal


def test_sharded_device_world1(gtb, ctx, oracle):
    """ShardedDeviceOverlap degenerates to the single-GPU engine when there is one rank."""
    import torch
    from gtb200 import sharded
    reads = support.synth_reads(300_000, seed=31)
    regions = support.synth_regions(2_000, seed=32)
    plan = sharded.ShardPlan(regions, 1, chrom_extent=support.HG19_LENS)
    dev = {k: torch.from_numpy(v).cuda() for k, v in reads.items()}
    assert bool(sharded.route_mask_torch(plan, dev, 0).cpu().numpy()[plan.route(reads, 0)].all())
    sh = sharded.ShardedDeviceOverlap(ctx, regions, plan, gtb.OP_COUNT, 0)
    dset, keep = gtb.device_set(dev)
    got = sh.step(dset, gtb.MEM_DEVICE).cpu().numpy().view(np.uint64)
    rc, want, _ = oracle.count(reads, regions, 0)
    assert rc == 0 and np.array_equal(got, want)
    # the same without a host wait inside the step (gtb_index_finish_async; streams ordered by events), status read afterwards
    for _ in range(3):
        out = sh.step(dset, gtb.MEM_DEVICE, defer_status=True)
    sh.check()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    # ... and a fatal query is reported by check() exactly as the waiting form raises it
    bad = {k: v.clone() for k, v in dev.items()}
    bad["stop"][1234] = 0
    bad["start"][1234] = -5
    bad["chrom"][1234] = int(regions["chrom"][0])
    bset, keep_bad = gtb.device_set(bad)
    with pytest.raises(gtb.GtbError) as e1:
        sh.step(bset, gtb.MEM_DEVICE)
    sh.step(bset, gtb.MEM_DEVICE, defer_status=True)
    with pytest.raises(gtb.GtbError) as e2:
        sh.check()
    assert (e1.value.code, e1.value.index) == (e2.value.code, e2.value.index) == (2, 1234)
    sh.close()


def test_synth_range_matches_numpy(gtb, ctx):
    import torch
    n = 100_000
    eff = support.HG19_LENS - 50 + 1
    total = int(eff.sum())
    rng = (total // 3, total // 2)
    dev = {"chrom": torch.empty(n, dtype=torch.int32, device="cuda"), "start": torch.empty(n, dtype=torch.int32, device="cuda"),
           "stop": torch.empty(n, dtype=torch.int32, device="cuda"), "strand": torch.empty(n, dtype=torch.int8, device="cuda")}
    ctx.synth_reads(9, 12345, n, 50, support.HG19_LENS, dev, p_range=rng)
    want = support.synth_reads(n, 9, first=12345, p_range=rng)
    for k in want:
        assert np.array_equal(dev[k].cpu().numpy(), want[k]), k


def test_full_size_properties(gtb, ctx, oracle):
    """BASELINE.json configs[1] at full size (100 M reads x 60 000 regions, strand-aware), where the oracle is too slow:
    size-independent properties.  (1) three independent device algorithms agree (bucket, direct, rank engines);
    (2) permutation invariance; (3) additivity over query batches; (4) checksum of checksums: sum_r count[r] equals
    sum_q #regions overlapping q, the latter from an independent per-QUERY formulation (torch.searchsorted over the
    regions' sorted starts / stops); (5) coverage >= count and coverage <= 50 * count for 50-bp reads;
    (6) the first 1 M reads of the same stream agree with the oracle."""
    import torch
    n, m = 100_000_000, 60_000
    regions = support.synth_regions(m, seed=3)
    dev = {"chrom": torch.empty(n, dtype=torch.int32, device="cuda"), "start": torch.empty(n, dtype=torch.int32, device="cuda"),
           "stop": torch.empty(n, dtype=torch.int32, device="cuda"), "strand": torch.empty(n, dtype=torch.int8, device="cuda")}
    ctx.synth_reads(2, 0, n, 50, support.HG19_LENS, dev)

    def run(tensors, op=gtb.OP_COUNT, engine=0, parts=1):
        ix = gtb.Index(ctx, regions, op, engine)
        step = (tensors["chrom"].numel() + parts - 1) // parts
        for lo in range(0, tensors["chrom"].numel(), step):
            ix.add_device({k: v[lo:lo + step] for k, v in tensors.items()})
        out = ix.finish()
        ix.close()
        return out

    base = run(dev)
    assert np.array_equal(run(dev, engine=gtb.ENGINE_BUCKET), base)
    assert np.array_equal(run(dev, engine=gtb.ENGINE_DIRECT), base)
    assert np.array_equal(run(dev, engine=gtb.ENGINE_RANK), base)
    assert np.array_equal(run(dev, parts=3), base)                                  # 33 333 334-read batches (unaligned tails)
    perm = torch.randperm(n, device="cuda")
    shuffled = {k: v[perm] for k, v in dev.items()}
    del perm
    assert np.array_equal(run(shuffled), base)
    del shuffled
    # per-query dual: #overlaps(q) = #{r in group: rs <= qe} - #{r in group: re < qs}
    grp_r = torch.from_numpy(regions["chrom"].astype(np.int64) * 2 + (regions["strand"] == ord("-"))).cuda()
    ks = torch.sort((grp_r << 32) + torch.from_numpy(regions["start"].astype(np.int64)).cuda()).values
    ke = torch.sort((grp_r << 32) + torch.from_numpy(regions["stop"].astype(np.int64)).cuda()).values
    total = 0
    for lo in range(0, n, 25_000_000):
        sl = slice(lo, lo + 25_000_000)
        g = (dev["chrom"][sl].long() * 2 + (dev["strand"][sl] == ord("-")).long()) << 32
        a = torch.searchsorted(ks, g + dev["stop"][sl].long(), right=True) - torch.searchsorted(ks, g, right=False)
        b = torch.searchsorted(ke, g + dev["start"][sl].long(), right=False) - torch.searchsorted(ke, g, right=False)
        total += int((a - b).sum().item())
    assert int(base.sum()) == total
    cov = run(dev, op=gtb.OP_COVERAGE)
    assert np.all(cov >= base) and np.all(cov <= 50 * base)
    sub = {k: v[:1_000_000].cpu().numpy() for k, v in dev.items()}
    rc, want, _ = oracle.count(sub, regions, 0)
    got = run({k: v[:1_000_000] for k, v in dev.items()})
    assert rc == 0 and np.array_equal(got, want)


def test_packed_host_path(gtb, ctx, oracle, monkeypatch):
    """Host-resident chunks travel re-encoded (5, 6 or 8 B/interval) when every query fits a packed form and in the plain
    layout otherwise; both must give the oracle's values, as must a run with the packing pool switched off."""
    reads = support.synth_reads(600_000, seed=41)
    regions = support.synth_regions(4_000, seed=42)
    odd = {k: v.copy() for k, v in reads.items()}
    odd["stop"][123_456] = odd["start"][123_456] + 70_000          # longer than the packed length field
    odd["strand"][400_000] = ord(".")                               # a strand byte the packed form cannot carry
    odd["chrom"][7] = 20_000                                        # chromosome id beyond the packed field (not in the index)
    for q in (reads, odd):
        for op, fn in ((gtb.OP_COUNT, oracle.count), (gtb.OP_COVERAGE, oracle.coverage)):
            rc, want, _ = fn(q, regions, 0)
            assert rc == 0
            ix = gtb.Index(ctx, regions, op, 0)
            ix.add_host(q)
            assert np.array_equal(ix.finish(), want)
            ix.close()
    # the three widths of the packed form: one read length -> 5 B/interval, lengths below 256 -> 6, longer -> 8, and a chunk
    # that fits none of them -> the plain 13; a fresh context each so that the byte counts are the form's own
    rng = np.random.default_rng(43)
    n = len(reads["chrom"])
    short = {k: v.copy() for k, v in reads.items()}
    short["stop"] = (short["start"] + rng.integers(0, 256, size=n)).astype(np.int32)
    longer = {k: v.copy() for k, v in reads.items()}
    longer["stop"] = (longer["start"] + rng.integers(0, 3000, size=n)).astype(np.int32)
    many_chrom = {k: v.copy() for k, v in reads.items()}
    many_chrom["chrom"][rng.random(n) < 0.01] = 150                   # beyond the 7-bit chromosome field of the narrow forms
    for q, bytes_per_interval in ((reads, 5), (short, 6), (longer, 8), (many_chrom, 8), (odd, 13)):
        for op, fn in ((gtb.OP_COUNT, oracle.count), (gtb.OP_COVERAGE, oracle.coverage)):
            rc, want, _ = fn(q, regions, 0)
            assert rc == 0
            c1 = gtb.Context(0)
            ix = gtb.Index(c1, regions, op, 0)
            ix.add_host(q)
            assert np.array_equal(ix.finish(), want), (bytes_per_interval, op)
            assert c1.transfer_stats()["h2d_bytes"] == bytes_per_interval * n, (bytes_per_interval, c1.transfer_stats())
            ix.close()
            c1.close()
    monkeypatch.setenv("GTB_INGEST_THREADS", "0")
    c2 = gtb.Context(0)
    rc, want, _ = oracle.count(reads, regions, 0)
    assert np.array_equal(c2.overlap_count(reads, regions, 0), want)
    c2.close()


def test_packed_reads_entry(gtb, ctx, oracle):
    """gtb_index_add_packed: reads of one length handed over as int32 start + one byte of chromosome | strand (host and device
    memory, count and coverage) give what the same reads give as a gtb_set."""
    import torch
    n = 3_000_001                                                       # not a multiple of anything
    reads = support.synth_reads(n, seed=21, read_len=36)
    regions = support.synth_regions(5_000, seed=22)
    meta = (reads["chrom"].astype(np.uint8) | np.where(reads["strand"] == ord("-"), 0x80, 0).astype(np.uint8))
    start = np.ascontiguousarray(reads["start"])
    for op, fn in ((gtb.OP_COUNT, oracle.count), (gtb.OP_COVERAGE, oracle.coverage)):
        for flags in (0, gtb.IGNORE_STRAND):
            rc, want, _ = fn(reads, regions, flags)
            assert rc == 0
            ix = gtb.Index(ctx, regions, op, flags)
            ix.add_packed_ptr(n, start.ctypes.data, meta.ctypes.data, 36, gtb.MEM_HOST)
            assert np.array_equal(ix.finish(), want), ("host", op, flags)
            ix.reset()
            d_start, d_meta = torch.from_numpy(start).cuda(), torch.from_numpy(meta).cuda()
            ix.add_packed_ptr(n, d_start.data_ptr(), d_meta.data_ptr(), 36, gtb.MEM_DEVICE)
            assert np.array_equal(ix.finish(), want), ("device", op, flags)
            ix.close()


@pytest.mark.parametrize("n", [1, 31, 4097, 1_000_003])
def test_sort_regions_matches_stable_lexsort(ctx, n):
    """gtb_sort_regions (device LSD radix sort) against numpy's stable lexsort on the same key: chromosome rank, ('+' first),
    start ascending, stop descending, ties in input order -- the order of GenomicRegionSet::RunGlobalSort."""
    rng = np.random.default_rng(n)
    chrom = rng.integers(0, 7 if n > 100 else 2, n).astype(np.int32)
    start = rng.integers(-50, 400 if n < 10_000 else 3_000_000, n).astype(np.int32)          # many ties on small inputs, negative starts too
    stop = (start + rng.integers(0, 6, n) * 100).astype(np.int32)
    strand = rng.choice(np.array([43, 45, 46], dtype=np.int8), n)
    if n > 1000:
        start[::5000] = np.int32(2**31 - 1000); stop[::5000] = np.int32(2**31 - 1)            # the far end of the coordinate range
        start[1::5000] = np.int32(-2**31); stop[1::5000] = np.int32(-2**31 + 5)
    for by_strand in (False, True):
        got = ctx.sort_regions(chrom, start, stop, strand, by_strand)
        sclass = (strand != 43).astype(np.int64) if by_strand else np.zeros(n, np.int64)
        want = np.lexsort((-stop.astype(np.int64), start.astype(np.int64), sclass, chrom.astype(np.int64)))    # last key is the primary one; lexsort is stable
        assert np.array_equal(got, want), (n, by_strand)


def test_sharded_scan_cuda_engine(gtb, oracle):
    """gtb200.sharded.ShardedScan with the CUDA engine, every shard of a 3-way plan run in turn on this GPU: the kept windows of
    the shards, put together, are the unsharded result (the N > 1 exchange itself is covered on the CPU over gloo)."""
    from gtb200 import sharded
    lens = support.HG19_LENS[:6]
    reads = support.synth_reads(2_000_000, seed=41, chrom_lens=lens)
    n, want = oracle.scan_counts(reads, lens, 50, 200, "1", False, 1)
    plan = sharded.ScanShardPlan(lens, 50, 200, 3)
    parts = []
    for shard in range(3):
        eng = sharded.CudaScanEngine(lens, 50, 200, "1", False, 1, 0)
        ids = plan.route(reads, shard)
        eng.add({k: np.ascontiguousarray(v[ids]) for k, v in reads.items()})
        got = eng.finish()
        keep = plan.owns(got["chrom"], got["win"], shard)
        parts.append({k: v[keep] for k, v in got.items()})
        eng.close()
    allw = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    order = np.lexsort((allw["win"], allw["strand"] != ord("+"), allw["chrom"]))
    assert len(order) == n
    for k in ("chrom", "strand", "win", "value"):
        assert np.array_equal(allw[k][order], want[k]), k
