"""Read pairs under `count` WITHOUT -gaps (a region counts a pair once if either mate overlaps it: not a sum over the mates, so the
rank formulation does not apply and the candidate-enumeration engine serves it) next to `count -gaps` (the pair's span) and
`coverage`: 20 M pairs vs 60 k regions, device-resident, CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python"), os.path.join(ROOT, "tests")]
import torch
import gtb200
import support

n_pairs = 20_000_000
torch.cuda.set_device(0)
ctx = gtb200.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
regions = support.synth_regions(60_000, 3)
m1 = {"chrom": torch.empty(n_pairs, dtype=torch.int32, device="cuda"), "start": torch.empty(n_pairs, dtype=torch.int32, device="cuda"),
      "stop": torch.empty(n_pairs, dtype=torch.int32, device="cuda"), "strand": torch.empty(n_pairs, dtype=torch.int8, device="cuda")}
ctx.synth_reads(5, 0, n_pairs, 50, support.HG19_LENS, m1)
lens = torch.from_numpy(support.HG19_LENS).cuda()
idx = torch.arange(n_pairs, device="cuda", dtype=torch.int64)
gap = ((idx * 2654435761) >> 7) % 301 + 100
s1 = m1["start"].long()
over = torch.clamp(s1 + 100 + gap - 1 - lens[m1["chrom"].long()], min=0)
s1 = torch.clamp(s1 - over, min=1)
dev = {"chrom": torch.repeat_interleave(m1["chrom"], 2), "strand": torch.repeat_interleave(m1["strand"], 2),
       "start": torch.stack([s1, s1 + 50 + gap], 1).reshape(-1).int()}
dev["stop"] = dev["start"] + 49
dset, keep = gtb200.device_set(dev, per_region=2)
out = torch.zeros(60_000, dtype=torch.int64, device="cuda")
res = {}
for name, op, flags in (("count", gtb200.OP_COUNT, 0), ("count -gaps", gtb200.OP_COUNT, gtb200.MATCH_GAPS), ("coverage", gtb200.OP_COVERAGE, 0)):
    ix = gtb200.Index(ctx, regions, op, flags)

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
print(json.dumps({"workload": "%d read pairs (two 50-bp mates, gap 100-400 bp) vs 60 k regions" % n_pairs, "results": res}))
