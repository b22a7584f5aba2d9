"""count without -gaps over spliced reads (regions of 1-4 blocks, CSR offsets): the one-pass engine on the spans + enumeration of the
multi-block regions whose span holds an evaluation point, against enumeration of every region (GTB_NO_MULTI_FAST=1)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import gtb200
import support
import test_baseline_configs as bc

n = 10_000_000
q, off = bc.synth_spliced(n, seed=92)
regions = support.synth_regions(60_000, 3)
torch.cuda.set_device(0)
ctx = gtb200.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
dev = {k: torch.from_numpy(v).cuda() for k, v in q.items()}
dev_off = torch.from_numpy(off).cuda()
dset, keep = gtb200.device_set(dev, offsets=dev_off)
out = torch.zeros(60_000, dtype=torch.int64, device="cuda")
res = {}
for name, env in (("one pass + enumeration of the exceptions", None), ("enumeration of every region", "1")):
    if env:
        os.environ["GTB_NO_MULTI_FAST"] = env
    ix = gtb200.Index(ctx, regions, gtb200.OP_COUNT, 0)

    def step():
        ix.reset(); ix.add_set(dset, gtb200.MEM_DEVICE); ix.finish_ptr(out.data_ptr(), gtb200.MEM_DEVICE)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        step()
    e1.record(stream); torch.cuda.synchronize()
    res[name] = {"ms_per_step": e0.elapsed_time(e1) / 3, "sum": int(out.sum().item())}
    ix.close()
assert len({v["sum"] for v in res.values()}) == 1, res
print(json.dumps({"workload": "%d spliced reads (%d blocks) vs 60 k regions, count without -gaps" % (n, len(q["chrom"])), "results": res}))
