"""Parity on the BASELINE.json configurations at the sizes SURVEY.md section 8d names for them ("next" item 1).

  (a) configs[0]: the reference's shipped examples/genes.bed.gz (space-separated BED6, mm9 chr1) x 5 M seed-1 reads:
      count / coverage / density, default and -S, the command-line drivers byte for byte against the reference binaries.
  (b) configs[1]: the first 10 M reads of the 100 M-read stream x 60 000 regions, against the oracle.
  (c) configs[2]: chr21 + chr22 x 20 M reads, `genomic_scans counts -w 200 -d 50`, -min 10 and -min 1.
  (d) configs[3]: 10 M read PAIRS (two-interval regions), coverage with and without -gaps through the C ABI against the
      oracle; density (the drivers' arithmetic on it) byte for byte on 1 M pairs.
  (e) configs[4] shape: 1 M regions x 10 M reads.
  (f) scale limits: one region's count past 2^32 and one chromosome-strand past 2^31 intervals (streamed batches), one region's
      coverage past 2^32, label weights.

examples/genes.bed.gz is copied next to the reference binaries by oracle/Makefile (`examples`): /root/reference does not exist on
the GPU box."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import support

pytestmark = pytest.mark.gpu

BIN = os.path.join(support.ROOT, "ibm-cbc-genomic-tools_b200", "bin")
GENES = os.path.join(support.REF_DIR, "examples", "genes.bed.gz")
MM9_CHR1 = 197195432


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


def run_new(tool, args):
    p = subprocess.run([os.path.join(BIN, tool)] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout, p.stderr


def same_as_reference(tool, args):
    want = support.run_ref(tool, args, check=False)
    got = run_new(tool, args)
    assert got[0] == want[0] == 0, (args, got[2][-300:], want[2][-300:])
    assert len(want[1]) > 0 and got[1] == want[1], (args, got[1][:200], want[1][:200])
    return want[1]


def write_bed_fast(path, reads, names, sep="\t"):
    """BED6 text for millions of reads without a Python loop per line."""
    import pandas as pd
    n = len(reads["chrom"])
    df = pd.DataFrame({"c": np.asarray(names, dtype=object)[reads["chrom"]], "s": reads["start"].astype(np.int64) - 1, "e": reads["stop"].astype(np.int64),
                       "l": np.char.add("r", np.arange(n).astype(str)), "x": np.zeros(n, dtype=np.int8),
                       "t": np.where(reads["strand"] == ord("-"), "-", "+")})
    df.to_csv(path, sep=sep, header=False, index=False)


def read_genes():
    """examples/genes.bed.gz parsed independently of the host reader: whitespace-separated BED6, 0-based half-open -> 1-based closed."""
    chrom, start, stop, strand = [], [], [], []
    with gzip.open(GENES, "rt") as f:
        for line in f:
            t = line.split()
            assert t[0] == "chr1"
            chrom.append(0); start.append(int(t[1]) + 1); stop.append(int(t[2])); strand.append(ord(t[5]))
    return {"chrom": np.array(chrom, np.int32), "start": np.array(start, np.int32), "stop": np.array(stop, np.int32), "strand": np.array(strand, np.int8)}


# ---------------------------------------------------------------------------------------------------------------------------------
# (a) configs[0]: the shipped example
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mm9_reads(tmp_path_factory):
    if not (support.have_ref() and os.path.exists(GENES)):
        pytest.skip("reference binaries / examples not built into oracle/_ref")
    d = tmp_path_factory.mktemp("cfg0")
    # seed 1, 5 M 50-bp reads uniform on mm9 chr1 [3 000 000, 197 195 332]
    reads = support.synth_reads(5_000_000, seed=1, read_len=50, chrom_lens=np.array([MM9_CHR1]), p_range=(2_999_999, MM9_CHR1 - 100))
    write_bed_fast(str(d / "reads.bed"), reads, ["chr1"])
    order = np.argsort(reads["start"], kind="stable")
    srt = {k: v[order] for k, v in reads.items()}
    write_bed_fast(str(d / "reads.sorted.bed"), srt, ["chr1"])
    order = np.lexsort((reads["start"], reads["strand"]))                 # '+' (0x2B) before '-' (0x2D), then start: what -S -s wants
    write_bed_fast(str(d / "reads.strand_sorted.bed"), {k: v[order] for k, v in reads.items()}, ["chr1"])
    # genes.bed.gz is not sorted by start within chr1: the sorted forms for -S
    with gzip.open(GENES, "rt") as f:
        lines = f.readlines()
    lines.sort(key=lambda ln: int(ln.split()[1]))
    (d / "genes.sorted.bed").write_text("".join(lines))
    lines.sort(key=lambda ln: (ln.split()[5], int(ln.split()[1])))
    (d / "genes.strand_sorted.bed").write_text("".join(lines))
    return d, reads


@pytest.mark.parametrize("op", ["count", "coverage", "density"])
def test_config0_shipped_example_cli(mm9_reads, op):
    d, _ = mm9_reads
    same_as_reference("genomic_overlaps", [op, GENES, d / "reads.bed"])
    same_as_reference("genomic_overlaps", [op, "-i", GENES, d / "reads.bed"])
    same_as_reference("genomic_overlaps", [op, "-S", d / "genes.sorted.bed", d / "reads.sorted.bed"])
    same_as_reference("genomic_overlaps", [op, "-S", "-s", d / "genes.strand_sorted.bed", d / "reads.strand_sorted.bed"])


def test_config0_shipped_example_abi(mm9_reads, ctx, oracle, gtb):
    _, reads = mm9_reads
    genes = read_genes()
    assert len(genes["chrom"]) == 4785
    for flags in (0, gtb.IGNORE_STRAND):
        rc, want, _ = oracle.count(reads, genes, flags)
        assert rc == 0 and want.sum() > 0
        assert np.array_equal(ctx.overlap_count(reads, genes, flags), want)
        for eng in (gtb.ENGINE_DIRECT, gtb.ENGINE_BUCKET, gtb.ENGINE_RANK):
            assert np.array_equal(ctx.overlap_count(reads, genes, flags | eng), want), eng
        rc, want, _ = oracle.coverage(reads, genes, flags)
        assert rc == 0
        assert np.array_equal(ctx.overlap_coverage(reads, genes, flags), want)
        assert np.array_equal(ctx.overlap_coverage(reads, genes, flags | gtb.ENGINE_DIRECT), want)


# ---------------------------------------------------------------------------------------------------------------------------------
# (b) configs[1]: first 10 M reads x 60 000 regions     (e) 1 M regions x 10 M reads
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,seed,lens", [(60_000, 3, (500, 500_000)), (1_000_000, 7, (200, 100_000))], ids=["60k_regions", "1M_regions"])
def test_config1_and_4_first_10M_reads(ctx, oracle, gtb, m, seed, lens):
    import torch
    n = 10_000_000
    regions = support.synth_regions(m, seed, lens[0], lens[1])
    reads = support.synth_reads(n, seed=2)
    dev = {"chrom": torch.empty(n, dtype=torch.int32, device="cuda"), "start": torch.empty(n, dtype=torch.int32, device="cuda"),
           "stop": torch.empty(n, dtype=torch.int32, device="cuda"), "strand": torch.empty(n, dtype=torch.int8, device="cuda")}
    ctx.synth_reads(2, 0, n, 50, support.HG19_LENS, dev)
    for flags in (0, gtb.IGNORE_STRAND) if m < 90_000 else (0,):           # (the scalar oracle needs half a minute per pass over 1 M regions)
        rc, want, _ = oracle.count(reads, regions, flags)
        assert rc == 0
        for eng in (0, gtb.ENGINE_BUCKET) + ((gtb.ENGINE_DIRECT,) if m < 90_000 else ()):
            ix = gtb.Index(ctx, regions, gtb.OP_COUNT, flags | eng)
            ix.add_device(dev)
            got = ix.finish()
            ix.close()
            assert np.array_equal(got, want), (m, flags, eng)
    # coverage: all 10 M reads against the 60 000 regions, the first 2 M against the 1 M regions
    n_cov = n if m < 90_000 else 2_000_000
    rc, want, _ = oracle.coverage({k: v[:n_cov] for k, v in reads.items()}, regions, 0)
    assert rc == 0
    ix = gtb.Index(ctx, regions, gtb.OP_COVERAGE, 0)
    ix.add_device({k: v[:n_cov] for k, v in dev.items()})
    assert np.array_equal(ix.finish(), want)
    ix.close()


# ---------------------------------------------------------------------------------------------------------------------------------
# (c) configs[2]: chr21 + chr22 x 20 M reads, window counts
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("min_reads", [10, 1])
def test_config2_scans_chr21_chr22(ctx, oracle, gtb, min_reads):
    lens = np.array([support.HG19["chr21"], support.HG19["chr22"]], dtype=np.int64)
    reads = support.synth_reads(20_000_000, seed=4, chrom_lens=lens)
    n, want = oracle.scan_counts(reads, lens, 50, 200, "1", False, min_reads)
    assert n > 100_000
    sc = gtb.Scan(ctx, lens, 50, 200, "1", False, min_reads)
    for lo in range(0, 20_000_000, 7_000_000):                             # three unequal host batches
        sc.add_host({k: v[lo:lo + 7_000_000] for k, v in reads.items()})
    n_got = sc.finish()
    assert n_got == n
    got = sc.fetch(0, n_got)
    for k in ("chrom", "strand", "win", "value"):
        assert np.array_equal(got[k], want[k]), k
    sc.close()


# ---------------------------------------------------------------------------------------------------------------------------------
# (d) configs[3]: read pairs as two-interval regions
# ---------------------------------------------------------------------------------------------------------------------------------
def synth_pairs(n_pairs, seed):
    """mate 1 = a 50-bp read of the stream, gap uniform 100-400 bp, mate 2 50 bp, same strand; pairs whose
    second mate would leave the chromosome are pulled back (both mates shifted left)."""
    m1 = support.synth_reads(n_pairs, seed=seed)
    gap = (support.splitmix64(np.arange(n_pairs, dtype=np.uint64) * np.uint64(77) + np.uint64(seed)) % np.uint64(301)).astype(np.int64) + 100
    span = 50 + gap + 50
    over = np.maximum(m1["start"].astype(np.int64) + span - 1 - support.HG19_LENS[m1["chrom"]], 0)
    s1 = np.maximum(m1["start"].astype(np.int64) - over, 1)
    s2 = s1 + 50 + gap
    chrom = np.repeat(m1["chrom"], 2)
    strand = np.repeat(m1["strand"], 2)
    start = np.empty(2 * n_pairs, dtype=np.int32); stop = np.empty(2 * n_pairs, dtype=np.int32)
    start[0::2] = s1; stop[0::2] = s1 + 49
    start[1::2] = s2; stop[1::2] = s2 + 49
    off = np.arange(0, 2 * n_pairs + 1, 2, dtype=np.int64)
    return {"chrom": chrom, "start": start, "stop": stop, "strand": strand}, off


def test_config3_paired_coverage_abi(ctx, oracle, gtb):
    pairs, off = synth_pairs(10_000_000, seed=5)
    regions = support.synth_regions(60_000, 3)
    for flags in (0, gtb.MATCH_GAPS, gtb.IGNORE_STRAND):
        rc, want, _ = oracle.coverage(pairs, regions, flags, qoff=off)
        assert rc == 0 and want.sum() > 0
        got = ctx.overlap_coverage(pairs, regions, flags, qoffsets=off)
        assert np.array_equal(got, want), flags
    rc, want, _ = oracle.count(pairs, regions, gtb.MATCH_GAPS, qoff=off)
    assert rc == 0
    assert np.array_equal(ctx.overlap_count(pairs, regions, gtb.MATCH_GAPS, qoffsets=off), want)


def test_config3_pairs_without_offsets(ctx, oracle, gtb):
    """Read pairs handed over as regions of two intervals each WITHOUT offsets (gtb_set.region_offset == NULL, n_intervals ==
    2 * n_regions): coverage takes the intervals as they lie and checks every pair in the engine's registers; the other shapes
    (count, -gaps, a forced engine) get their offsets written out on the device.  Host and device memory, odd batch sizes, and the
    fatal cases with the first offending REGION's stream index."""
    import torch
    pairs, off = synth_pairs(3_000_001, seed=15)
    regions = support.synth_regions(60_000, 3)
    dev = {k: torch.from_numpy(v).cuda() for k, v in pairs.items()}
    for op, fn, flag_sets in ((gtb.OP_COVERAGE, oracle.coverage, (0, gtb.MATCH_GAPS, gtb.IGNORE_STRAND, gtb.ENGINE_RANK)), (gtb.OP_COUNT, oracle.count, (gtb.MATCH_GAPS,))):
        for flags in flag_sets:
            rc, want, _ = fn(pairs, regions, flags & 3, qoff=off)
            assert rc == 0
            ix = gtb.Index(ctx, regions, op, flags)
            ix.add_host(pairs, per_region=2)
            assert np.array_equal(ix.finish(), want), ("host", op, flags)
            ix.reset()
            cut = 2 * 1_234_567                                            # two device batches, the second one not a multiple of a tile
            ix.add_device({k: v[:cut] for k, v in dev.items()}, per_region=2)
            ix.add_device({k: v[cut:] for k, v in dev.items()}, per_region=2)
            assert np.array_equal(ix.finish(), want), ("device", op, flags)
            ix.close()
    # count WITHOUT -gaps: a region counts a pair once if either mate overlaps it.  Against short regions (20-600 bp, many of them
    # inside the 100-400 bp gaps of the pairs) the pairs whose span holds an evaluation point go to the enumeration engine, the rest
    # are counted as spans in the one pass
    short = support.synth_regions(80_000, 33, 20, 600)
    for flags in (0, gtb.IGNORE_STRAND):
        rc, want, _ = oracle.count(pairs, short, flags, qoff=off)
        rc2, with_gaps, _ = oracle.count(pairs, short, flags | gtb.MATCH_GAPS, qoff=off)
        assert rc == 0 and rc2 == 0 and not np.array_equal(want, with_gaps)               # (the two semantics do differ on this input)
        ix = gtb.Index(ctx, short, gtb.OP_COUNT, flags)
        ix.add_host(pairs, per_region=2)
        assert np.array_equal(ix.finish(), want), ("count, host", flags)
        ix.reset()
        ix.add_device({k: v[:cut] for k, v in dev.items()}, per_region=2)
        ix.add_device({k: v[cut:] for k, v in dev.items()}, per_region=2)
        assert np.array_equal(ix.finish(), want), ("count, device", flags)
        ix.close()
    # a malformed pair (mates overlap), a pair on two chromosomes, and a pair whose span ends at or before 0 on an indexed chromosome
    for where, kind in ((2_000_123, "overlap"), (2_999_999, "chrom"), (777, "stop")):
        bad = {k: v.copy() for k, v in pairs.items()}
        i = 2 * where
        if kind == "overlap":
            bad["start"][i + 1] = bad["stop"][i]
        elif kind == "chrom":
            bad["chrom"][i + 1] = (bad["chrom"][i] + 1) % 24
        else:
            bad["chrom"][i] = bad["chrom"][i + 1] = regions["chrom"][0]
            bad["start"][i], bad["stop"][i], bad["start"][i + 1], bad["stop"][i + 1] = -90, -60, -40, -3
        for op, fn, flags in ((gtb.OP_COVERAGE, oracle.coverage, 0), (gtb.OP_COVERAGE, oracle.coverage, gtb.MATCH_GAPS), (gtb.OP_COUNT, oracle.count, 0)):
            rc, _, ei = fn(bad, regions, flags, qoff=off)                  # the pairs checked next to their blocks / the spans formed in the engine
            assert rc != 0 and ei == where
            ix = gtb.Index(ctx, regions, op, flags)
            ix.add_host(bad, per_region=2)
            with pytest.raises(gtb.GtbError) as e:
                ix.finish()
            assert (e.value.code, e.value.index) == (rc, where), (kind, op, flags, e.value.code, e.value.index, rc)
            ix.close()


def synth_spliced(n, seed):
    """reads of one to four blocks (30-80 bp) separated by gaps of 50-2000 bp, CSR offsets; about half of them single-block"""
    rng = np.random.default_rng(seed)
    base = support.synth_reads(n, seed=seed)
    k = np.where(rng.random(n) < 0.5, 1, rng.integers(2, 5, n))
    off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
    m = int(off[-1])
    reg = np.repeat(np.arange(n), k)
    j = np.arange(m) - off[reg]                                           # block number inside its region
    blen = rng.integers(30, 81, m)
    gap = rng.integers(50, 2001, m)
    step = np.where(j > 0, np.roll(blen, 1) + gap, 0)                      # distance from the previous block's start
    rel = np.cumsum(step) - np.repeat(np.cumsum(step)[off[:-1]], k)       # start relative to the region's first block
    start = (base["start"][reg].astype(np.int64) + rel)
    room = support.HG19_LENS[base["chrom"]][reg] - (start + blen)         # keep every block inside its chromosome
    shift = np.repeat(np.minimum.reduceat(np.minimum(room, 0), off[:-1]), k)
    start = np.maximum(start + shift, 1)
    return ({"chrom": base["chrom"][reg].astype(np.int32), "start": start.astype(np.int32), "stop": (start + blen - 1).astype(np.int32),
             "strand": base["strand"][reg].astype(np.int8)}, off)


def test_count_spliced_reads_without_gaps(ctx, oracle, gtb):
    """count without -gaps over regions of one to four blocks (what spliced reads in a SAM file become): the spans go through the
    one-pass engine, the multi-block regions whose span holds an evaluation point through the enumeration engine -- against short
    index regions (many of them inside the gaps between blocks, where the span's count would be wrong), host and device memory"""
    import torch
    q, off = synth_spliced(900_001, seed=91)
    short = support.synth_regions(80_000, 33, 20, 600)
    genes = support.synth_regions(60_000, 3)
    dev = {k: torch.from_numpy(v).cuda() for k, v in q.items()}
    dev_off = torch.from_numpy(off).cuda()
    for idx in (short, genes):
        for flags in (0, gtb.IGNORE_STRAND):
            rc, want, _ = oracle.count(q, idx, flags, qoff=off)
            assert rc == 0 and want.sum() > 0
            ix = gtb.Index(ctx, idx, gtb.OP_COUNT, flags)
            ix.add_host(q, offsets=off)
            assert np.array_equal(ix.finish(), want), ("host", flags)
            ix.reset()
            ix.add_device(dev, offsets=dev_off)
            assert np.array_equal(ix.finish(), want), ("device", flags)
            ix.close()
    rc, with_gaps, _ = oracle.count(q, short, gtb.MATCH_GAPS, qoff=off)
    rc2, without, _ = oracle.count(q, short, 0, qoff=off)
    assert rc == 0 and rc2 == 0 and not np.array_equal(with_gaps, without)             # (the input does tell the two semantics apart)
    # a malformed region and a fatal span keep their stream index
    bad = {k: v.copy() for k, v in q.items()}
    r_multi = int(np.flatnonzero(np.diff(off) > 1)[1234])
    bad["start"][off[r_multi] + 1] = bad["stop"][off[r_multi]]             # blocks overlap
    rc, _, ei = oracle.count(bad, genes, 0, qoff=off)
    assert rc != 0 and ei == r_multi
    ix = gtb.Index(ctx, genes, gtb.OP_COUNT, 0)
    ix.add_host(bad, offsets=off)
    with pytest.raises(gtb.GtbError) as e:
        ix.finish()
    assert (e.value.code, e.value.index) == (rc, r_multi)
    ix.close()


def test_pairs_without_offsets_skewed(ctx, oracle, gtb):
    """most pairs piled onto four loci, in random order: byte counters overflow before their spills land, the batch is dropped
    by the engine (its queries were spans formed in registers: no replay there) and goes down the general path -- or it is
    not, depending on timing; exact either way"""
    pairs, off = synth_pairs(1_500_000, seed=16)
    rng = np.random.default_rng(17)
    n = len(off) - 1
    hot = rng.random(n) < 0.9
    loc = (1_000_000 + rng.integers(0, 4, n) * 37).astype(np.int32)
    for m in (0, 1):
        pairs["chrom"][m::2][hot] = 2
    pairs["start"][0::2][hot] = loc[hot]; pairs["stop"][0::2][hot] = loc[hot] + 49
    pairs["start"][1::2][hot] = loc[hot] + 250; pairs["stop"][1::2][hot] = loc[hot] + 299
    regions = support.synth_regions(3_000, seed=32)
    short = support.synth_regions(3_000, 34, 20, 600)
    short["chrom"][:200] = 2; short["start"][:200] = (1_000_000 + np.arange(200) * 3).astype(np.int32); short["stop"][:200] = short["start"][:200] + 60
    for op, fn, flags, idx in ((gtb.OP_COVERAGE, oracle.coverage, gtb.MATCH_GAPS, regions), (gtb.OP_COUNT, oracle.count, gtb.MATCH_GAPS, regions),
                               (gtb.OP_COUNT, oracle.count, 0, short), (gtb.OP_COVERAGE, oracle.coverage, 0, regions)):
        rc, want, _ = fn(pairs, idx, flags, qoff=off)
        assert rc == 0
        ix = gtb.Index(ctx, idx, op, flags)
        for rep in range(2):
            ix.reset()
            ix.add_host(pairs, per_region=2)
            assert np.array_equal(ix.finish(), want), (op, flags, rep)
        ix.close()


def test_config3_paired_density_cli(tmp_path):
    if not support.have_ref():
        pytest.skip("reference binaries not built (oracle/_ref)")
    pairs, off = synth_pairs(1_000_000, seed=5)
    regions = support.synth_regions(20_000, 3)
    support.write_bed(str(tmp_path / "regions.bed"), regions, support.HG19_NAMES, labels=["g%06d" % k for k in range(20_000)])
    # REG: label <TAB> chrom strand start stop chrom strand start stop
    names = np.asarray(support.HG19_NAMES, dtype=object)
    with open(tmp_path / "pairs.reg", "w") as f:
        c = names[pairs["chrom"][0::2]]
        t = np.where(pairs["strand"][0::2] == ord("-"), "-", "+")
        a, b, c2, d2 = pairs["start"][0::2], pairs["stop"][0::2], pairs["start"][1::2], pairs["stop"][1::2]
        f.write("".join("p%d\t%s %s %d %d %s %s %d %d\n" % (k, c[k], t[k], a[k], b[k], c[k], t[k], c2[k], d2[k]) for k in range(len(c))))
    for op in ("coverage", "density"):
        same_as_reference("genomic_overlaps", [op, tmp_path / "regions.bed", tmp_path / "pairs.reg"])
        same_as_reference("genomic_overlaps", [op, "-gaps", tmp_path / "regions.bed", tmp_path / "pairs.reg"])


# ---------------------------------------------------------------------------------------------------------------------------------
# (f) scale limits
# ---------------------------------------------------------------------------------------------------------------------------------
def test_scale_limits_counts_past_2_32(ctx, gtb):
    """43 batches of 100 M reads, all on chr1 '+' inside one 1-Mbp region: that region's count passes 2^32, the
    chromosome-strand holds more than 2^31 intervals, and one batch's coverage of the region passes 2^32.  Expected values come
    from torch (an independent formulation: per-read overlap lengths summed in int64)."""
    import torch
    n, batches = 100_000_000, 43
    g = torch.Generator(device="cuda"); g.manual_seed(99)
    start = torch.randint(1_000_000, 2_000_000 - 49, (n,), device="cuda", dtype=torch.int32, generator=g)
    dev = {"chrom": torch.zeros(n, dtype=torch.int32, device="cuda"), "start": start, "stop": start + 49,
           "strand": torch.full((n,), ord("+"), dtype=torch.int8, device="cuda")}
    regions = {"chrom": np.zeros(4, np.int32), "start": np.array([1_000_000, 1_500_000, 1_999_990, 1_000_000], np.int32),
               "stop": np.array([2_000_000, 3_000_000, 2_500_000, 2_000_000], np.int32), "strand": np.array([43, 43, 43, 45], np.int8)}
    s64, e64 = dev["start"].long(), dev["stop"].long()
    want_count, want_cov = [], []
    for k in range(4):
        rs, re_ = int(regions["start"][k]), int(regions["stop"][k])
        hit = (s64 <= re_) & (e64 >= rs) if regions["strand"][k] == 43 else torch.zeros_like(s64, dtype=torch.bool)
        want_count.append(int(hit.sum().item()) * batches)
        ov = (torch.minimum(e64, torch.tensor(re_, device="cuda")) - torch.maximum(s64, torch.tensor(rs, device="cuda")) + 1).clamp(min=0)
        want_cov.append(int((ov * hit).sum().item()) * batches)
    assert want_count[0] == n * batches > 2 ** 32 and want_cov[0] > 2 ** 32
    for op, want in ((gtb.OP_COUNT, want_count), (gtb.OP_COVERAGE, want_cov)):
        ix = gtb.Index(ctx, regions, op, 0)
        for _ in range(batches):
            ix.add_device(dev)
        got = ix.finish()
        ix.close()
        assert [int(x) for x in got] == want, (op, got, want)


def test_scale_limits_weights(ctx, oracle, gtb):
    """--max-label-value weights near the 32-bit limit: a few thousand reads push a count and a coverage past 2^32."""
    rng = np.random.default_rng(5)
    n = 5000
    reads = {"chrom": np.zeros(n, np.int32), "start": rng.integers(1, 900, n).astype(np.int32), "strand": np.full(n, 43, np.int8)}
    reads["stop"] = (reads["start"] + rng.integers(0, 200, n)).astype(np.int32)
    w = rng.integers(2_000_000_000, 2_147_483_647, n).astype(np.int32)
    regions = {"chrom": np.zeros(2, np.int32), "start": np.array([1, 400], np.int32), "stop": np.array([1000, 600], np.int32), "strand": np.array([43, 43], np.int8)}
    rc, want, _ = oracle.count(reads, regions, 0, qw=w)
    assert rc == 0 and want[0] > 2 ** 32
    assert np.array_equal(ctx.overlap_count(reads, regions, 0, qweight=w), want)
    rc, want, _ = oracle.coverage(reads, regions, 0, qw=w)
    assert rc == 0 and want[0] > 2 ** 40
    assert np.array_equal(ctx.overlap_coverage(reads, regions, 0, qweight=w), want)


def test_scale_limits_scan_values_past_2_32(ctx, oracle, gtb):
    """Window counts whose micro-windows and window values pass 2^32 (weighted reads piled onto a few micro-windows, added in
    two batches with a reset in between): the table's carry plane and the upper halves of the emitted values come into use,
    and go out of use again after the reset."""
    rng = np.random.default_rng(9)
    n = 4000
    lens = np.array([support.HG19["chr21"], support.HG19["chr22"]], dtype=np.int64)
    reads = {"chrom": rng.integers(0, 2, n).astype(np.int32), "start": rng.integers(1_000_000, 1_000_400, n).astype(np.int32),
             "strand": np.where(rng.integers(0, 2, n) == 1, 43, 45).astype(np.int8)}
    reads["stop"] = (reads["start"] + 49).astype(np.int32)
    w = rng.integers(2_000_000_000, 2_147_483_647, n).astype(np.int32)
    n_want, want = oracle.scan_counts(reads, lens, 50, 200, "1", False, 1, weight=w)
    assert want["value"].max() > 2 ** 33
    sc = gtb.Scan(ctx, lens, 50, 200, "1", False, 1)
    sc.add_host({k: v[:n // 2] for k, v in reads.items()}, weight=w[:n // 2])
    sc.add_host({k: v[n // 2:] for k, v in reads.items()}, weight=w[n // 2:])
    assert sc.finish() == n_want
    got = sc.fetch(0, n_want)
    for k in ("chrom", "strand", "win", "value"):
        assert np.array_equal(got[k], want[k]), k
    # after a reset the same object counts small values again (the carry plane is cleared, the values are narrow)
    sc.reset()
    plain = support.synth_reads(500_000, seed=44, chrom_lens=lens)
    n_want, want = oracle.scan_counts(plain, lens, 50, 200, "1", False, 1)
    sc.add_host(plain)
    assert sc.finish() == n_want
    got = sc.fetch(0, n_want)
    for k in ("chrom", "strand", "win", "value"):
        assert np.array_equal(got[k], want[k]), k
    sc.close()
