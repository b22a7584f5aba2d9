"""Per-QUERY overlap counts (gtb_index_query_counts: the core of `genomic_overlaps subset / overlap`) against a brute-force
restatement of the reference's predicate (GenomicRegionSetOverlaps::GetOverlap, genomic_intervals.cpp:5224-5236: span match,
then any pair of intervals unless -gaps, then strand unless -i), and against the per-REGION engine: both sides of the same
overlap relation must add up to the same number of pairs."""
import numpy as np
import pytest

import randcases
import support

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gtb():
    import gtb200
    return gtb200


@pytest.fixture(scope="module")
def ctx(gtb):
    c = gtb.Context(0)
    yield c
    c.close()


def brute(q, qoff, idx, ioff, gaps, ignore_strand):
    """n(q) by the definition, one (query, region) pair at a time (small inputs only)."""
    qoff = np.arange(len(q["chrom"]) + 1) if qoff is None else qoff
    ioff = np.arange(len(idx["chrom"]) + 1) if ioff is None else ioff
    out = np.zeros(len(qoff) - 1, dtype=np.uint32)
    present = set(int(idx["chrom"][ioff[k]]) for k in range(len(ioff) - 1)
                  if not (idx["start"][ioff[k]] > idx["stop"][ioff[k + 1] - 1] or idx["stop"][ioff[k + 1] - 1] <= 0))
    for a in range(len(qoff) - 1):
        lo, hi = qoff[a], qoff[a + 1]
        c, sb, qs, qe = int(q["chrom"][lo]), int(q["strand"][lo]), int(q["start"][lo]), int(q["stop"][hi - 1])
        if c not in present:
            continue
        n = 0
        for k in range(len(ioff) - 1):
            ilo, ihi = ioff[k], ioff[k + 1]
            rs, re_ = int(idx["start"][ilo]), int(idx["stop"][ihi - 1])
            if rs > re_ or re_ <= 0 or int(idx["chrom"][ilo]) != c:
                continue
            if not (qs <= re_ and qe >= rs):
                continue
            if not ignore_strand and int(idx["strand"][ilo]) != sb:
                continue
            if not gaps:
                hit = False
                for i in range(lo, hi):
                    for j in range(ilo, ihi):
                        if not (q["start"][i] > idx["stop"][j] or q["stop"][i] < idx["start"][j]):
                            hit = True
                if not hit:
                    continue
            n += 1
        out[a] = n
    return out


@pytest.mark.parametrize("seed", range(4))
def test_query_counts_vs_brute_force(gtb, ctx, seed):
    rng = np.random.default_rng(900 + seed)
    if seed % 2 == 0:
        idx, ioff = (randcases.rand_grid if seed else randcases.rand_single)(rng, 150), None
        q, qoff = (randcases.rand_grid if seed else randcases.rand_single)(rng, 1500, strands="+-."), None
        q["chrom"][rng.random(1500) < 0.03] = 9                              # a chromosome the index has never seen
    else:
        idx, ioff = randcases.rand_multi(rng, 120)
        q, qoff = randcases.rand_multi(rng, 800)
    for flags in range(4):
        gaps, ign = bool(flags & gtb.MATCH_GAPS), bool(flags & gtb.IGNORE_STRAND)
        ix = gtb.Index(ctx, idx, gtb.OP_COUNT, flags, roffsets=ioff)
        got = ix.query_counts(q, offsets=qoff)
        want = brute(q, qoff, idx, ioff, gaps, ign)
        assert np.array_equal(got, want), (seed, flags, np.nonzero(got != want)[0][:5])
        # the other side of the same relation: per-region counts of the same (unweighted) queries
        ix.add_host(q, offsets=qoff)
        assert int(ix.finish().sum()) == int(got.sum())
        ix.close()


def test_query_counts_hg19_consistency(gtb, ctx):
    import torch
    n = 3_000_000
    reads = support.synth_reads(n, seed=31)
    regions = support.synth_regions(60_000, seed=3)
    for flags in (0, gtb.IGNORE_STRAND):
        ix = gtb.Index(ctx, regions, gtb.OP_COUNT, flags)
        got = ix.query_counts(reads)
        ix.add_host(reads)
        per_region = ix.finish()
        assert int(got.sum()) == int(per_region.sum()) > 0
        # spot check against numpy on the first 20 000 reads
        m = 20_000
        grp = lambda s: s["chrom"].astype(np.int64) * (1 if flags else 3) + (0 if flags else (s["strand"] == 45))
        want = np.zeros(m, dtype=np.uint32)
        gr = grp(regions)
        for g in np.unique(grp({k: v[:m] for k, v in reads.items()})):
            sel = np.nonzero(grp({k: v[:m] for k, v in reads.items()}) == g)[0]
            rs = np.sort(regions["start"][gr == g]); re_ = np.sort(regions["stop"][gr == g])
            want[sel] = np.searchsorted(rs, reads["stop"][sel], side="right") - np.searchsorted(re_, reads["start"][sel], side="left")
        assert np.array_equal(got[:m], want)
        ix.close()


def test_query_counts_fatal_queries(gtb, ctx):
    idx = {"chrom": np.array([0, 0], np.int32), "start": np.array([10, 500], np.int32), "stop": np.array([100, 900], np.int32), "strand": np.array([43, 43], np.int8)}
    q = {"chrom": np.array([0, 0, 0], np.int32), "start": np.array([5, 300, 50], np.int32), "stop": np.array([20, 200, 60], np.int32), "strand": np.array([43, 43, 43], np.int8)}
    ix = gtb.Index(ctx, idx, gtb.OP_COUNT, 0)
    with pytest.raises(gtb.GtbError) as e:
        ix.query_counts(q)
    assert e.value.code == gtb.ERR_QUERY_START_GT_STOP and e.value.index == 1
    q["start"][1] = 150
    assert list(ix.query_counts(q)) == [1, 0, 1]
    ix.close()


def bin_order_key(idx, ioff, bits):
    """(level, bin, -k) of every admitted index region in the reference's bin index (genomic_intervals.cpp:5653-5672)"""
    ioff = np.arange(len(idx["chrom"]) + 1) if ioff is None else ioff
    bits = list(bits) + [60]
    key = {}
    for k in range(len(ioff) - 1):
        s, e = int(idx["start"][ioff[k]]), int(idx["stop"][ioff[k + 1] - 1])
        if s > e or e <= 0:
            continue
        s = max(s, 1)
        for l, b in enumerate(bits):
            if (s >> b) == (e >> b):
                key[k] = (l, s >> b, -k)
                break
    return key


@pytest.mark.parametrize("seed", range(4))
def test_query_matches_vs_brute_force(gtb, ctx, seed):
    """gtb_index_query_matches: the matching index regions of every query, in the order of the reference's walk (bin levels, bins,
    LIFO chains; file order under GTB_SORTED_RULES), against the pair-by-pair definition"""
    rng = np.random.default_rng(1900 + seed)
    if seed % 2 == 0:
        idx, ioff = randcases.rand_single(rng, 300), None
        q, qoff = randcases.rand_single(rng, 2500, strands="+-."), None
        idx["start"] = (idx["start"].astype(np.int64) * 37).astype(np.int32); idx["stop"] = (idx["stop"].astype(np.int64) * 37 + 30).astype(np.int32)
        q["start"] = (q["start"].astype(np.int64) * 37).astype(np.int32); q["stop"] = (q["stop"].astype(np.int64) * 37 + 30).astype(np.int32)
        bad = (q["start"] > q["stop"]) | (q["stop"] <= 0)
        q["start"][bad] = 5; q["stop"][bad] = 50
    else:
        idx, ioff = randcases.rand_multi(rng, 200)
        q, qoff = randcases.rand_multi(rng, 1200)
    qo = np.arange(len(q["chrom"]) + 1) if qoff is None else qoff
    io = np.arange(len(idx["chrom"]) + 1) if ioff is None else ioff
    for flags, bits in ((0, None), (gtb.IGNORE_STRAND, [3, 6]), (gtb.MATCH_GAPS, [5]), (gtb.MATCH_GAPS | gtb.IGNORE_STRAND, [2, 4, 6, 8, 10, 12, 14]), (gtb.SORTED_RULES, None)):
        odd = np.array([idx["start"][io[k]] > idx["stop"][io[k + 1] - 1] or idx["stop"][io[k + 1] - 1] <= 0 for k in range(len(io) - 1)])
        ix = gtb.Index(ctx, idx, gtb.OP_COUNT, flags, roffsets=ioff)
        if flags & gtb.SORTED_RULES and odd.any():
            with pytest.raises(gtb.GtbError):
                ix.query_matches(q, offsets=qoff)
            ix.close()
            continue
        off, got = ix.query_matches(q, offsets=qoff, bin_bits=bits)
        counts = brute(q, qoff, idx, ioff, bool(flags & gtb.MATCH_GAPS), bool(flags & gtb.IGNORE_STRAND))
        assert np.array_equal(np.diff(off), counts), (seed, flags)
        key = bin_order_key(idx, ioff, bits or [17, 20, 23, 26])
        for a in rng.choice(len(qo) - 1, 300, replace=False):
            m = [int(v) for v in got[off[a]:off[a + 1]]]
            if len(m) < 2:
                continue
            want = sorted(m) if flags & gtb.SORTED_RULES else sorted(m, key=lambda k: key[k])
            assert m == want, (seed, flags, a, m, want)
            one = brute({k: v[qo[a]:qo[a + 1]] for k, v in q.items()}, None if qoff is None else np.array([0, qo[a + 1] - qo[a]]), idx, ioff,
                        bool(flags & gtb.MATCH_GAPS), bool(flags & gtb.IGNORE_STRAND))
            assert int(one[0]) == len(m) == len(set(m))
        ix.close()
    # offsets that do not belong to these queries are an argument error, not a buffer overrun
    ix = gtb.Index(ctx, idx, gtb.OP_COUNT, 0, roffsets=ioff)
    st, keep = gtb.host_set(q, None, qoff)
    import ctypes
    off = np.zeros(st.n_regions + 1, dtype=np.int64)
    matches = np.zeros(4, dtype=np.int32)
    err = ctypes.c_int64(-1)
    rc = gtb.lib().gtb_index_query_matches(ix._h, ctypes.byref(st), gtb.MEM_HOST, None, 0, off.ctypes.data_as(ctypes.c_void_p), matches.ctypes.data_as(ctypes.c_void_p), ctypes.byref(err))
    assert rc == gtb.ERR_ARG
    ix.close()
