"""Seeded random small cases rich in the edge conditions of SURVEY.md section 8a' (abutting,
contained and duplicate intervals, chromosomes present on one side only, multi-interval regions,
weights).  TEST INFRASTRUCTURE ONLY."""
import numpy as np

NAMES = ["chr1", "chr10", "chr2", "chrM", "chrX"]   # already in strcmp order


def rand_single(rng, n, n_chrom=4, span=2000, max_len=300, chrom_lo=0, strands="+-"):
    chrom = rng.integers(chrom_lo, chrom_lo + n_chrom, size=n).astype(np.int32)
    start = rng.integers(1, span, size=n).astype(np.int32)
    length = rng.integers(1, max_len, size=n)
    # sprinkle degenerate lengths: single-base and exactly abutting common coordinates
    length[rng.random(n) < 0.1] = 1
    snap = rng.random(n) < 0.15
    start[snap] = (start[snap] // 100) * 100 + 1
    stop = (start + length - 1).astype(np.int32)
    strand = np.array([ord(strands[i]) for i in rng.integers(0, len(strands), size=n)], dtype=np.int8)
    return {"chrom": chrom, "start": start, "stop": stop, "strand": strand}


def rand_grid(rng, n, n_chrom=3, span=1500, grid=50, max_len=6, strands="+-"):
    """Coordinates snapped to a coarse grid so that qs == re, qe == rs, qe == rs-1 all occur often."""
    chrom = rng.integers(0, n_chrom, size=n).astype(np.int32)
    start = (rng.integers(0, span // grid, size=n) * grid + rng.integers(0, 3, size=n)).astype(np.int32) + 1
    stop = (start + rng.integers(0, max_len, size=n) * grid + rng.integers(-1, 2, size=n)).astype(np.int32)
    stop = np.maximum(stop, start)
    strand = np.array([ord(strands[i]) for i in rng.integers(0, len(strands), size=n)], dtype=np.int8)
    return {"chrom": chrom, "start": start, "stop": stop.astype(np.int32), "strand": strand}


def rand_multi(rng, n_regions, n_chrom=3, span=3000, max_blocks=4, max_len=120, max_gap=150):
    """Well-formed multi-interval regions (same chrom/strand, sorted, non-overlapping)."""
    chrom, start, stop, strand, offsets = [], [], [], [], [0]
    for _ in range(n_regions):
        c = int(rng.integers(0, n_chrom)); s = ord("+-"[int(rng.integers(0, 2))])
        nb = int(rng.integers(1, max_blocks + 1))
        pos = int(rng.integers(1, span))
        for _b in range(nb):
            ln = int(rng.integers(1, max_len))
            chrom.append(c); strand.append(s); start.append(pos); stop.append(pos + ln - 1)
            pos = pos + ln + int(rng.integers(0, max_gap))   # gap 0 => abutting blocks (start = prev stop + 1)
        offsets.append(len(chrom))
    return ({"chrom": np.array(chrom, dtype=np.int32), "start": np.array(start, dtype=np.int32),
             "stop": np.array(stop, dtype=np.int32), "strand": np.array(strand, dtype=np.int8)},
            np.array(offsets, dtype=np.int64))
