"""Several GPUs behind one index in ONE process (gtb_mgpu_*): host-resident reads in the packed form (5 B/read), cut into one
slice per device, each slice over its own host link; wall-clock per step (add + finish), 1 .. N devices, same total."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import gtb200
import support

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
n_dev = torch.cuda.device_count()
regions = support.synth_regions(60_000, 3)
# the reads are made on device 0 and brought to pinned host memory in the packed form
torch.cuda.set_device(0)
ctx = gtb200.Context(0)
start = torch.empty(n, dtype=torch.int32).pin_memory()
meta = torch.empty(n, dtype=torch.uint8).pin_memory()
PIECE = 100_000_000
for lo in range(0, n, PIECE):
    m = min(PIECE, n - lo)
    t = {"chrom": torch.empty(m, dtype=torch.int32, device="cuda"), "start": torch.empty(m, dtype=torch.int32, device="cuda"),
         "stop": torch.empty(m, dtype=torch.int32, device="cuda"), "strand": torch.empty(m, dtype=torch.int8, device="cuda")}
    ctx.synth_reads(2, lo, m, 50, support.HG19_LENS, t)
    start[lo:lo + m].copy_(t["start"])
    meta[lo:lo + m].copy_((t["chrom"].to(torch.uint8) | ((t["strand"] == ord("-")).to(torch.uint8) << 7)))
    del t
torch.cuda.synchronize()
ctx.close()
res, sums = {}, {}
counts = [k for k in (1, 2, 4, 8) if k <= n_dev]
for k in counts:
    mg = gtb200.MultiGpu(list(range(k)))
    ix = gtb200.MultiIndex(mg, regions, gtb200.OP_COUNT, 0)

    def step():
        ix.reset()
        ix.add_packed_ptr(n, start.data_ptr(), meta.data_ptr(), 50)
        return ix.finish()
    for _ in range(2):
        out = step()
    t0 = time.perf_counter()
    for _ in range(5):
        out = step()
    dt = (time.perf_counter() - t0) / 5
    res[k] = {"ms_per_step": dt * 1e3, "reads_per_s": n / dt, "h2d_GBps": 5 * n / dt / 1e9}
    sums[k] = int(out.sum())
    ix.close(); mg.close()
assert len(set(sums.values())) == 1, sums
print(json.dumps({"workload": "%d host-resident reads (packed, 5 B/read) vs 60 k regions, count, one process" % n, "devices": res, "checksum": sums[counts[0]]}))
