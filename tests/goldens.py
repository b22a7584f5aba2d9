"""Loader for tests/golden/*.npz (made by tests/golden/make_golden.py from the reference binaries)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _set(z, p):
    return {k: z[p + k] for k in ("chrom", "start", "stop", "strand")}


def overlap_cases():
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "ov_*.npz"))):
        z = np.load(path)
        out.append(dict(name=os.path.basename(path)[:-4], index=_set(z, "i_"), queries=_set(z, "q_"),
                        ioff=z["i_offsets"] if "i_offsets" in z else None,
                        qoff=z["q_offsets"] if "q_offsets" in z else None,
                        qw=z["q_weight"] if "q_weight" in z else None,
                        expect={(op, f): z["%s_%d" % (op, f)] for op in ("count", "coverage") for f in range(4)}))
    return out


def scan_cases():
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "scan_*.npz"))):
        z = np.load(path)
        w, d, op, ign, mn = [int(x) for x in z["params"]]
        out.append(dict(name=os.path.basename(path)[:-4], reads=_set(z, "r_"), bounds=z["bounds"],
                        rw=z["r_weight"] if "r_weight" in z else None,
                        win_size=w, win_step=d, op=chr(op), ignore_strand=bool(ign), min_reads=mn,
                        expect={k: z["o_" + k] for k in ("chrom", "strand", "win", "value")}))
    return out
