"""Wall clock of the command-line drivers on a small input (2 M reads, 60 000 regions), leaving the process by exit() (the CUDA
runtime unwinds; the default) against leaving it by _exit() after the flush (GT_FAST_EXIT=1).  (Call AG ran this when _exit() was
the default and exit() was selected by GT_FAST_EXIT=0.)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests")]
import support

BIN = os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "bin")
tmp = "/dev/shm/gt_exit_ab"
os.makedirs(tmp, exist_ok=True)
reads = os.path.join(tmp, "reads.bed")
with open(reads, "wb") as f:
    subprocess.check_call([os.path.join(BIN, "gt_synth_bed"), "2000000", "5"], stdout=f)
regions = support.synth_regions(60_000, 3)
support.write_bed(os.path.join(tmp, "genes.bed"), regions, support.HG19_NAMES, ["g%d" % k for k in range(60_000)])
genes = os.path.join(tmp, "genes.bed")
cases = {"count": [os.path.join(BIN, "genomic_overlaps"), "count", genes, reads],
         "coverage": [os.path.join(BIN, "genomic_overlaps"), "coverage", genes, reads],
         "gsort": [os.path.join(BIN, "genomic_regions"), "gsort", genes]}
out = {}
for name, cmd in cases.items():
    res = {}
    digest = None
    for mode, env in (("exit", {"GT_FAST_EXIT": "0"}), ("_exit", {"GT_FAST_EXIT": "1"})):
        ts = []
        for rep in range(7):
            t0 = time.perf_counter()
            p = subprocess.run(cmd, env=dict(os.environ, **env), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            ts.append(time.perf_counter() - t0)
            assert p.returncode == 0, p.stderr[-300:]
            assert digest is None or digest == p.stdout, "output differs"
            digest = p.stdout
        ts = sorted(ts[1:])
        res[mode] = {"median_s": ts[len(ts) // 2], "min_s": ts[0], "max_s": ts[-1]}
    p = subprocess.run(cmd, env=dict(os.environ, GT_TIMING="1"), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    res["phases"] = p.stderr.decode().strip().splitlines()[-1]
    out[name] = res
print(json.dumps(out))
