"""--max-label-value weights: 100 M weighted 50-bp reads vs 60 k regions, count and coverage, the DIRECT engine's weighted form
against the general rank step (the only path for weighted queries before).  Device-resident, CUDA events, results compared."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python"), os.path.join(ROOT, "tests")]
import torch
import gtb200
import support

n = 100_000_000
torch.cuda.set_device(0)
ctx = gtb200.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
regions = support.synth_regions(60_000, 3)
t = {"chrom": torch.empty(n, dtype=torch.int32, device="cuda"), "start": torch.empty(n, dtype=torch.int32, device="cuda"),
     "stop": torch.empty(n, dtype=torch.int32, device="cuda"), "strand": torch.empty(n, dtype=torch.int8, device="cuda")}
ctx.synth_reads(2, 0, n, 50, support.HG19_LENS, t)
w = (torch.arange(n, device="cuda", dtype=torch.int64) * 2654435761 >> 9) % 5 + 1        # weights 1..5
w = w.int()
dset, keep = gtb200.device_set(t, weight=w)
out = torch.zeros(60_000, dtype=torch.int64, device="cuda")
res = {}
for op, opname in ((gtb200.OP_COUNT, "count"), (gtb200.OP_COVERAGE, "coverage")):
    sums = {}
    for eng, ename in ((0, "default (direct, weighted)"), (gtb200.ENGINE_RANK, "general rank step")):
        ix = gtb200.Index(ctx, regions, op, eng)

        def step():
            ix.reset(); ix.add_set(dset, gtb200.MEM_DEVICE); ix.finish_ptr(out.data_ptr(), gtb200.MEM_DEVICE)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            step()
        e1.record(stream); torch.cuda.synchronize()
        sums[ename] = int(out.sum().item())
        res["%s, %s" % (opname, ename)] = e0.elapsed_time(e1) / 5
        ix.close()
    assert len(set(sums.values())) == 1, sums
print(json.dumps({"workload": "100 M weighted reads (weights 1..5) vs 60 k regions", "ms_per_step": res}))
