"""Generates tests/golden/*.npz by running the UNMODIFIED reference binaries (oracle/_ref/, built by
`make -C oracle ref` from /root/reference) on seeded small inputs.  Run in the authoring container:

    python tests/golden/make_golden.py

Each .npz holds the packed SoA inputs (the arrays the C-ABI takes) and the values the reference
printed, so the golden tests need neither the reference sources nor its binaries."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import randcases  # noqa: E402
import support  # noqa: E402

FLAGS = [(0, []), (2, ["-i"]), (1, ["-gaps"]), (3, ["-gaps", "-i"])]


def ref_values(op, args, rf, qf):
    _, out, _ = support.run_ref("genomic_overlaps", [op] + args + [rf, qf])
    return np.array([int(v) for v in support.parse_label_values(out)[1]], dtype=np.uint64)


def overlaps_case(name, idx, q, ioff=None, qoff=None, labels=None, maxlab=1):
    with tempfile.TemporaryDirectory() as d:
        rf, qf = os.path.join(d, "ref"), os.path.join(d, "q")
        if ioff is None and qoff is None:
            support.write_bed(rf, idx, randcases.NAMES); support.write_bed(qf, q, randcases.NAMES, labels=labels)
        else:
            support.write_reg(rf, idx, randcases.NAMES, offsets=ioff)
            support.write_reg(qf, q, randcases.NAMES, labels=labels, offsets=qoff)
        out = {}
        for flags, args in FLAGS:
            a = args + (["--max-label-value", str(maxlab)] if maxlab > 1 else [])
            out["count_%d" % flags] = ref_values("count", a, rf, qf)
            out["coverage_%d" % flags] = ref_values("coverage", a, rf, qf)
    arrays = {"i_" + k: v for k, v in idx.items()}
    arrays.update({"q_" + k: v for k, v in q.items()})
    if ioff is not None: arrays["i_offsets"] = ioff
    if qoff is not None: arrays["q_offsets"] = qoff
    if labels is not None: arrays["q_weight"] = np.minimum(maxlab, labels).astype(np.int32)
    arrays.update(out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(name, {k: int(v.sum()) for k, v in out.items()})


def scans_case(name, reads, bounds, w, d, op, ignore, mn, labels=None, maxlab=1):
    with tempfile.TemporaryDirectory() as dd:
        gf, qf = os.path.join(dd, "g"), os.path.join(dd, "r")
        with open(gf, "w") as f:
            for c, b in enumerate(bounds):
                if b >= 0: f.write("%s\t0\t%d\n" % (randcases.NAMES[c], b))
        support.write_bed(qf, reads, randcases.NAMES, labels=labels)
        args = ["counts", "-g", gf, "-w", w, "-d", d, "-op", op, "-min", mn] + (["-i"] if ignore else [])
        if maxlab > 1: args += ["--max-label-value", maxlab]
        _, out, _ = support.run_ref("genomic_scans", args + [qf])
    chrom, strand, win, val = [], [], [], []
    for line in out.decode().splitlines():
        v, rest = line.split("\t"); c, s, a, _b = rest.split(" ")
        chrom.append(randcases.NAMES.index(c)); strand.append(ord(s)); win.append((int(a) - 1) // d + 1); val.append(int(v))
    arrays = {"r_" + k: v for k, v in reads.items()}
    if labels is not None: arrays["r_weight"] = np.minimum(maxlab, labels).astype(np.int32)
    arrays.update(bounds=np.asarray(bounds, dtype=np.int64), params=np.array([w, d, ord(op), int(ignore), mn], dtype=np.int64),
                  o_chrom=np.array(chrom, dtype=np.int32), o_strand=np.array(strand, dtype=np.int8),
                  o_win=np.array(win, dtype=np.int64), o_value=np.array(val, dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(name, len(val), "windows")


def main():
    assert support.have_ref(), "build the reference first: make -C oracle ref"
    rng = np.random.default_rng(20261018)
    overlaps_case("ov_single_uniform", randcases.rand_single(rng, 200), randcases.rand_single(rng, 5000))
    overlaps_case("ov_single_grid", randcases.rand_grid(rng, 150), randcases.rand_grid(rng, 4000))
    q = randcases.rand_single(rng, 3000)
    overlaps_case("ov_single_weighted", randcases.rand_grid(rng, 100), q, labels=rng.integers(-2, 9, size=3000), maxlab=5)
    idx, ioff = randcases.rand_multi(rng, 80); q, qoff = randcases.rand_multi(rng, 1500)
    overlaps_case("ov_multi_both", idx, q, ioff=ioff, qoff=qoff)
    idx, ioff = randcases.rand_multi(rng, 80); q = randcases.rand_single(rng, 3000, n_chrom=3, span=3500, max_len=60)
    overlaps_case("ov_multi_index_single_query", idx, q, ioff=ioff, qoff=np.arange(3001, dtype=np.int64))
    # long queries vs short regions (queries that swallow many regions)
    overlaps_case("ov_long_queries", randcases.rand_single(rng, 300, max_len=40), randcases.rand_single(rng, 800, max_len=1500))
    bounds = [5000, 120, 30, 3000, -1]
    reads = randcases.rand_single(rng, 4000, n_chrom=5, span=5200, max_len=120)
    scans_case("scan_w200_d50", reads, bounds, 200, 50, "1", False, 0)
    scans_case("scan_w200_d50_center_min3", reads, bounds, 200, 50, "c", False, 3)
    scans_case("scan_w100_d100_ignore", reads, bounds, 100, 100, "1", True, 1)
    scans_case("scan_w500_d25_weighted", reads, bounds, 500, 25, "1", False, 10, labels=rng.integers(0, 6, size=4000), maxlab=4)


if __name__ == "__main__":
    main()
