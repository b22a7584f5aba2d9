#!/usr/bin/env python
"""Secondary measurements for DESIGN.md: the other BASELINE.json configs on one GPU (not the driver's bench contract).

  python bench_configs.py [--reads N]

  configs[2]  genomic_scans counts, 200-bp windows step 50 over hg19 (-min 10), N reads   (BASELINE: 1 B reads)
  configs[3]  coverage over N intervals vs 60 k regions                                    (BASELINE: 1 B intervals)
  plus count with -i, and the sorted-input variant of configs[1]
Device-resident inputs, CUDA events, 3 warm-up + 5 timed steps each; one JSON line per measurement."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--regions", type=int, default=60_000, help="index regions (1000000 with --region-len 200,100000 = BASELINE configs[4])")
    ap.add_argument("--region-len", default="500,500000", help="min,max of the log-uniform region length")
    ap.add_argument("--only", default="", help="comma-separated subset of: count, coverage, sorted, scan")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)

    def wanted(tag):
        return not only or tag in only
    import torch
    import gtb200
    import support
    n = args.reads
    peak = 6547.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ctx = gtb200.Context(0)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    dev = {"chrom": torch.empty(n, dtype=torch.int32, device="cuda"), "start": torch.empty(n, dtype=torch.int32, device="cuda"),
           "stop": torch.empty(n, dtype=torch.int32, device="cuda"), "strand": torch.empty(n, dtype=torch.int8, device="cuda")}
    ctx.synth_reads(2, 0, n, 50, support.HG19_LENS, dev)
    dset, keep = gtb200.device_set(dev)
    rl = [int(x) for x in args.region_len.split(",")]
    regions = support.synth_regions(args.regions, 3 if args.regions == 60_000 else 7, rl[0], rl[1])
    m = args.regions

    def timed(name, fn, bytes_per_step, extra=None):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        line = {"measurement": name, "reads": n, "ms_per_step": ms, "intervals_per_s": n / (ms * 1e-3),
                "algorithmic_GBps": bytes_per_step / (ms * 1e-3) / 1e9, "frac_of_measured_peak": bytes_per_step / (ms * 1e-3) / 1e9 / peak}
        if extra:
            line.update(extra)
        ctx.profile(True)
        fn()
        torch.cuda.synchronize()
        line["kernel_ms"] = {k: round(v["total_ms"] / v["launches"], 4) for k, v in ctx.profile_report().items()}
        ctx.profile(False)
        print(json.dumps(line), flush=True)

    out = torch.zeros(m, dtype=torch.int64, device="cuda")
    for name, op, flags in (("count (configs[1])", gtb200.OP_COUNT, 0), ("count -i", gtb200.OP_COUNT, gtb200.IGNORE_STRAND),
                            ("coverage (configs[3], single-interval variant)", gtb200.OP_COVERAGE, 0)):
        if not wanted("coverage" if op == gtb200.OP_COVERAGE else "count"):
            continue
        ix = gtb200.Index(ctx, regions, op, flags)

        def step(ix=ix):
            ix.reset()
            ix.add_set(dset, gtb200.MEM_DEVICE)
            ix.finish_ptr(out.data_ptr(), gtb200.MEM_DEVICE)
        timed(name, step, 13 * n + 21 * m, {"regions": m})
        ix.close()

    if wanted("sorted"):
        # sorted input (what -S promises): same reads ordered by (chromosome, strand, start)
        key = (dev["chrom"].long() << 33) | ((dev["strand"] == ord("-")).long() << 32) | dev["start"].long()
        order = torch.argsort(key)
        del key
        sdev = {k: v[order].contiguous() for k, v in dev.items()}
        del order
        sset, keep2 = gtb200.device_set(sdev)
        ix = gtb200.Index(ctx, regions, gtb200.OP_COUNT, 0)

        def step_sorted():
            ix.reset()
            ix.add_set(sset, gtb200.MEM_DEVICE)
            ix.finish_ptr(out.data_ptr(), gtb200.MEM_DEVICE)
        timed("count, reads sorted by chromosome/strand/start", step_sorted, 13 * n + 21 * m, {"regions": m})
        ix.close()
        del sdev, sset

    if wanted("scan"):
        # scans: 200-bp windows, step 50, -min 10, strand-aware
        sc = gtb200.Scan(ctx, support.HG19_LENS, 50, 200, "1", False, 10)
        res = {}

        def step_scan():
            sc.reset()
            sc.add_set(dset, gtb200.MEM_DEVICE)
            res["n"] = sc.finish()
        windows = int(sum(max(int(L) // 50 - 3, 0) for L in support.HG19_LENS) * 2)
        timed("genomic_scans counts -w 200 -d 50 -min 10 (configs[2])", step_scan, 9 * n + 8 * windows, {"windows": windows})
        print(json.dumps({"qualifying_windows": res.get("n")}))
        sc.close()
    ctx.close()


if __name__ == "__main__":
    main()
