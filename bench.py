#!/usr/bin/env python
"""bench.py -- throughput of the overlap-count hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1 = BASELINE.json configs[1]: 100 M synthetic 50-bp hg19-shaped reads vs 60 000
gene-like regions, strand-aware `genomic_overlaps count`.  A step is one full pass of the hot path
over the batch: reset -> accumulate every query -> finalise per-region counts.
  value  : query intervals / s, device-timed (CUDA events), inputs already resident in HBM
  e2e    : the same through the C ABI with HOST (pinned) buffers, H2D + D2H inside the timed region
  roofline : dominant kernel's algorithmic bytes / its mean CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : the CPU path timed on this box's host cores on a bounded sample
For N > 1 (torchrun, one rank per GPU) every rank streams its own 100 M-read shard (weak scaling) and
the per-region counts of the genome shards are merged by ONE NCCL all-gather; value = all reads / max-over-ranks time.

  --config 1 (default)  the line above: the driver's bench and scaling runs
  --config 2            genomic_scans counts -w 200 -d 50 -min 10 over hg19, 1 B reads, one GPU
  --config 3            coverage over 1 B paired intervals (5e8 two-interval regions) vs 60 k regions, one GPU
  --config 4            count, 4 B reads vs 1 M regions, STRONG scaling over the GPUs given (the reads are split, not multiplied)
Configs 2-4 are the other BASELINE.json workloads, measured device-resident for the record (profiles/); they print the same
kind of line without the host-buffer and CPU legs.  Every config checks its result against an independent formulation in
torch (per-QUERY overlap counts by searchsorted; window sums by bincount + cumsum) before it prints: a line is only printed
for a correct result.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_READS = 100_000_000
N_REGIONS = 60_000
READ_LEN = 50
SEED_READS, SEED_REGIONS = 2, 3
BYTES_PER_QUERY, BYTES_PER_REGION = 13, 21
# BASELINE.json's metric; both arms print the same string so that the driver can divide one by the other (the reference arm's
# "device" is the host's cores: its clock also excludes parsing, like the GPU arm's)
METRIC = "query intervals/sec (overlap-count, device-timed)"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.splitlines()[0].split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_ref_engine(regions, reads_per_core, first_step, cores):
    """The UNMODIFIED reference engine (oracle/_ref/ref_engine: CountIndexOverlaps over in-memory sets) on `cores`
    processes, each on its own slice of the read stream.  Returns (reads processed, seconds = slowest process)."""
    import tempfile
    import support
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_engine")
    tmp = tempfile.mkdtemp(prefix="gtb_ref_")
    rf = os.path.join(tmp, "regions.bed")
    support.write_bed(rf, regions, support.HG19_NAMES)
    procs = [subprocess.Popen([exe, rf, str(SEED_READS), str((first_step * cores + p) * reads_per_core), str(reads_per_core), str(READ_LEN)],
                              stdout=subprocess.PIPE, text=True) for p in range(cores)]
    outs = [p.communicate()[0] for p in procs]
    secs = [float(json.loads(o.strip().splitlines()[-1])["engine_seconds"]) for o in outs]
    return reads_per_core * cores, max(secs)


def cpu_baseline(regions):
    """CPU path beside the GPU number: the reference's own engine on all host cores (kind "reference") when its
    binary travelled with the snapshot, else the scalar C port of the oracle on one core (kind "port").
    Bounded sample of the same workload, ~10-30 s of CPU work."""
    import support
    cores = os.cpu_count() or 1
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_engine")):
        per_core = 1_500_000
        n, secs = run_ref_engine(regions, per_core, 0, cores)
        return {"value": n / secs, "unit": "query intervals/s", "cores": cores, "kind": "reference",
                "sample": "%d reads per core x %d cores of the %d-read stream vs all %d regions; UnsortedGenomicRegionSetOverlaps + "
                          "CountIndexOverlaps only (sets pre-parsed in memory, no text parsing)" % (per_core, cores, N_READS, N_REGIONS)}
    sample = 40_000_000
    orc = support.Oracle()
    reads = support.synth_reads(sample, SEED_READS)
    t0 = time.perf_counter()
    rc, _, _ = orc.count(reads, regions, 0)
    dt = time.perf_counter() - t0
    assert rc == 0
    return {"value": sample / dt, "unit": "query intervals/s", "cores": 1, "kind": "port",
            "sample": "first %d of the %d reads vs all %d regions, engine only (no parsing)" % (sample, N_READS, N_REGIONS)}


def run_reference(args):
    """--impl reference: the CPU implementation on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import support
    regions = support.synth_regions(N_REGIONS, SEED_REGIONS)
    cores = os.cpu_count() or 1
    sample = 1_000_000
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_engine")
    kind = "reference" if os.path.exists(exe) else "port"
    times = []
    if kind == "reference":
        for step in range(args.warmup + args.steps):
            _, secs = run_ref_engine(regions, sample, step, cores)
            if step >= args.warmup:
                times.append(secs)
    else:
        orc = support.Oracle()
        cores = 1
        for step in range(args.warmup + args.steps):
            reads = support.synth_reads(sample, SEED_READS, first=step * sample)
            t0 = time.perf_counter()
            orc.count(reads, regions, 0)
            if step >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sample * cores * len(times)
    value = total / sum(times)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "query intervals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": "100M synthetic 50bp hg19 reads vs 60k gene regions, strand-aware count (configs[1])",
                       "n_regions": N_REGIONS, "read_len": READ_LEN},
            "cpu_baseline": {"value": value, "unit": "query intervals/s", "cores": cores, "kind": kind,
                             "sample": "%d reads per core per step x %d cores, engine only (CountIndexOverlaps), regions=%d" % (sample, cores, N_REGIONS)},
            "e2e": {"value": value, "unit": "query intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_other_config(args):
    """configs[2] (window counts, 1 B reads) and configs[3] (coverage over 1 B paired intervals) on one GPU, device-resident,
    CUDA events; the result is checked against an independent torch formulation before the line is printed."""
    import numpy as np
    import torch
    import gtb200
    import support
    torch.cuda.set_device(0)
    ctx = gtb200.Context(0)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    peak, peak_kind = measured_peak()
    SLICE = 1 << 30
    lens_dev = torch.from_numpy(support.HG19_LENS).cuda()

    def synth(n_reads, seed):
        t = {"chrom": torch.empty(n_reads, dtype=torch.int32, device="cuda"), "start": torch.empty(n_reads, dtype=torch.int32, device="cuda"),
             "stop": torch.empty(n_reads, dtype=torch.int32, device="cuda"), "strand": torch.empty(n_reads, dtype=torch.int8, device="cuda")}
        for lo in range(0, n_reads, SLICE):
            part = {k: v[lo:lo + SLICE] for k, v in t.items()}
            ctx.synth_reads(seed, lo, part["chrom"].numel(), READ_LEN, support.HG19_LENS, part)
        return t

    def timed(step):
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        launches0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(0) as clocks:
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        launches = ctx.launch_count() - launches0
        ctx.profile(True)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        prof = ctx.profile_report()
        ctx.profile(False)
        return ms, launches, clocks.summary(), prof

    def roofline_of(prof, bytes_dominant, bytes_step, ms):
        dom_name, dom = max(prof.items(), key=lambda kv: kv[1]["total_ms"])
        dom_ms = dom["total_ms"] / dom["launches"]
        total = sum(v["total_ms"] for v in prof.values())
        ach = bytes_dominant / (dom_ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel_ms": dom_ms, "kernel_share_of_step": dom["total_ms"] / total, "algorithmic_bytes_per_launch": bytes_dominant,
                "step_frac": bytes_step / (ms * 1e-3) / 1e9 / peak,
                "kernels": {k: {"launches_per_step": v["launches"] / 2, "ms_per_launch": v["total_ms"] / v["launches"]} for k, v in prof.items()}}

    n = args.reads
    if args.config == 2:
        step_bp, win_bp, min_reads = 50, 200, 10
        dev = synth(n, 4)
        sets, keep = [], []
        for lo in range(0, n, SLICE):
            st, k = gtb200.device_set({kk: v[lo:lo + SLICE] for kk, v in dev.items()})
            sets.append(st); keep.append(k)
        sc = gtb200.Scan(ctx, support.HG19_LENS, step_bp, win_bp, "1", False, min_reads)
        res = {}

        def step():
            sc.reset()
            for st in sets:
                sc.add_set(st, gtb200.MEM_DEVICE)
            res["n"] = sc.finish()
        ms, launches, clocks, prof = timed(step)
        # independent formulation: bincount of (slot, micro-window) + cumulative sums
        n_micro = support.HG19_LENS // step_bp
        off = np.concatenate([[0], np.cumsum(np.repeat(n_micro, 2))]).astype(np.int64)            # slot = 2 * chrom + (strand != '+')
        off_dev = torch.from_numpy(off).cuda()
        hist = torch.zeros(int(off[-1]) + 8, dtype=torch.int64, device="cuda")
        for lo in range(0, n, 50_000_000):
            sl = slice(lo, lo + 50_000_000)
            slot = dev["chrom"][sl].long() * 2 + (dev["strand"][sl] != ord("+")).long()
            w = (dev["start"][sl].long() - 1) // step_bp
            ok = w < torch.from_numpy(np.repeat(n_micro, 2)).cuda()[slot]
            hist.index_add_(0, (off_dev[slot] + w)[ok], torch.ones(int(ok.sum().item()), dtype=torch.int64, device="cuda"))
        combine = win_bp // step_bp
        want_n, want_sum = 0, 0
        for slot in range(len(off) - 1):
            h = hist[int(off[slot]):int(off[slot + 1])]
            if h.numel() < combine:
                continue
            cs = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(h, 0)])
            v = cs[combine:] - cs[:-combine]
            kept = v >= min_reads
            want_n += int(kept.sum().item()); want_sum += int(v[kept].sum().item())
        del hist
        got = sc.fetch(0, res["n"])
        assert res["n"] == want_n and int(got["value"].sum()) == want_sum, "window counts differ from the independent formulation: %d/%d windows, sums %d/%d" % (res["n"], want_n, int(got["value"].sum()), want_sum)
        windows = int(sum(max(int(L) // step_bp - combine + 1, 0) for L in support.HG19_LENS) * 2)
        bytes_step = 9 * n + 8 * windows                                 # chrom + start + strand per read, 8 B per window value
        line = {"metric": "reads/sec (genomic_scans counts, device-timed)", "value": n / (ms * 1e-3), "unit": "reads/s", "n_gpus": 1, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": {"workload": "genomic_scans counts -w 200 -d 50 -min 10 over hg19, %d synthetic 50bp reads, strand-aware (configs[2])" % n,
                           "reads": n, "windows": windows, "qualifying_windows": res["n"], "l2": "inputs (%.1f GB) and the window table far exceed the 126 MB L2" % (13 * n / 1e9)},
                "roofline": roofline_of(prof, 9 * min(n, SLICE), bytes_step, ms), "e2e": None, "gpu_launches": launches, "clocks": clocks,
                "checksum": want_sum, "checksum_verified": "window count and value sum equal torch bincount + cumsum"}
        print(json.dumps(line))
        sc.close()
    else:
        n_pairs = n // 2
        regions = support.synth_regions(N_REGIONS, SEED_REGIONS)
        m1 = synth(n_pairs, 5)
        # mate 2: same chromosome and strand, gap 100-400 bp (a hash of the pair index), pairs pulled back inside the chromosome
        idx = torch.arange(n_pairs, device="cuda", dtype=torch.int64)
        gap = ((idx * 2654435761) >> 7) % 301 + 100
        span = 2 * READ_LEN + gap
        s1 = m1["start"].long()
        over = torch.clamp(s1 + span - 1 - lens_dev[m1["chrom"].long()], min=0)
        s1 = torch.clamp(s1 - over, min=1)
        dev = {"chrom": torch.repeat_interleave(m1["chrom"], 2), "strand": torch.repeat_interleave(m1["strand"], 2),
               "start": torch.stack([s1, s1 + READ_LEN + gap], 1).reshape(-1).int(), "stop": None}
        dev["stop"] = dev["start"] + (READ_LEN - 1)
        del m1, idx, gap, span, s1, over
        off = torch.arange(0, 2 * n_pairs + 1, 2, dtype=torch.int64, device="cuda")
        dset_csr, keep = gtb200.device_set(dev, offsets=off)
        dset_uni, keep_u = gtb200.device_set(dev, per_region=2)           # the same pairs, declared as regions of two intervals: no offsets
        out = torch.zeros(N_REGIONS, dtype=torch.int64, device="cuda")
        lines = []
        for name, flags, dset, layout in (("coverage", 0, dset_uni, "regions of two intervals, no offsets"),
                                          ("coverage", 0, dset_csr, "CSR offsets"),
                                          ("coverage -gaps", gtb200.MATCH_GAPS, dset_uni, "regions of two intervals, no offsets"),
                                          ("coverage -gaps", gtb200.MATCH_GAPS, dset_csr, "CSR offsets")):
            index = gtb200.Index(ctx, regions, gtb200.OP_COVERAGE, flags)

            def step():
                index.reset()
                index.add_set(dset, gtb200.MEM_DEVICE)
                index.finish_ptr(out.data_ptr(), gtb200.MEM_DEVICE)
            ms, launches, clocks, prof = timed(step)
            # independent formulation: total overlap of [s, e] with its group's regions = F(e) - F(s - 1),
            # F(x) = sum_{rs <= x} (x + 1 - rs) - sum_{re < x} (x - re), from sorted region starts / stops and their prefix sums
            grp_r = torch.from_numpy(regions["chrom"].astype(np.int64) * 2 + (regions["strand"] == ord("-"))).cuda()
            rs = torch.from_numpy(regions["start"].astype(np.int64)).cuda(); re_ = torch.from_numpy(regions["stop"].astype(np.int64)).cuda()
            ks, order_s = torch.sort((grp_r << 32) + rs); ke, order_e = torch.sort((grp_r << 32) + re_)
            cs = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(rs[order_s], 0)])
            ce = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(re_[order_e], 0)])

            def F(g, x):
                a1, a0 = torch.searchsorted(ks, g + x, right=True), torch.searchsorted(ks, g, right=False)
                b1, b0 = torch.searchsorted(ke, g + x, right=False), torch.searchsorted(ke, g, right=False)
                return (a1 - a0) * (x + 1) - (cs[a1] - cs[a0]) - ((b1 - b0) * x - (ce[b1] - ce[b0]))
            want = 0
            for lo in range(0, n_pairs, 12_500_000):
                i0, i1 = 2 * lo, 2 * min(lo + 12_500_000, n_pairs)
                g = (dev["chrom"][i0:i1].long() * 2 + (dev["strand"][i0:i1] == ord("-")).long()) << 32
                s_, e_ = dev["start"][i0:i1].long(), dev["stop"][i0:i1].long()
                if flags & gtb200.MATCH_GAPS:                            # the span of the pair against the span of the region
                    g, s_, e_ = g[0::2], s_[0::2], e_[1::2]
                want += int((F(g, e_) - F(g, s_ - 1)).sum().item())
            got = int(out.sum().item())
            assert got == want, "%s: sum %d differs from the independent formulation's %d" % (name, got, want)
            bytes_step = 13 * 2 * n_pairs + (8 * (n_pairs + 1) if dset is dset_csr else 0) + BYTES_PER_REGION * N_REGIONS
            lines.append({"metric": "query intervals/sec (%s, device-timed)" % name, "value": 2 * n_pairs / (ms * 1e-3), "unit": "query intervals/s", "n_gpus": 1,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int64", "data": "synthetic",
                          "config": {"workload": "genomic_overlaps %s, %d synthetic read pairs (two 50-bp intervals per region, gap 100-400 bp) vs 60k regions (configs[3]); layout: %s" % (name, n_pairs, layout),
                                     "intervals": 2 * n_pairs, "n_regions": N_REGIONS, "l2": "inputs (%.1f GB) far exceed the 126 MB L2" % (bytes_step / 1e9)},
                          "roofline": roofline_of(prof, 13 * 2 * n_pairs, bytes_step, ms), "e2e": None, "gpu_launches": launches, "clocks": clocks,
                          "checksum": got, "checksum_verified": "sum of coverage equals the per-interval formulation's total (torch.searchsorted + prefix sums)"})
            index.close()
        for line in lines:
            print(json.dumps(line))
    ctx.close()


def main():
    global N_REGIONS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="BASELINE.json configs[N] (see the module docstring)")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (config 4: in total); 0 = what the config says")
    ap.add_argument("--regions", type=int, default=N_REGIONS, help="index regions (default: the BASELINE config; 1000000 = configs[4])")
    ap.add_argument("--engine", default="auto", choices=["auto", "rank", "bucket", "direct", "enumerate"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel iteration only: skip the host-buffer leg (the line is then not a valid bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    if args.reads <= 0:
        args.reads = {1: N_READS, 2: 1_000_000_000, 3: 1_000_000_000, 4: 4_000_000_000}[args.config]
    if args.config == 4 and args.regions == N_REGIONS:
        args.regions = 1_000_000
    if args.config in (2, 3):
        return run_other_config(args)

    import numpy as np
    import torch
    import gtb200
    import support

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    strong = args.config == 4
    SLICE = 1 << 30                                                     # reads per engine call (a batch's indices are 32-bit in places)
    workload = "100M synthetic 50bp hg19 reads vs 60k gene regions, strand-aware count (configs[1])"
    if args.regions != N_REGIONS:            # configs[4] shape: 1 M regions of 200 bp - 100 kb, seed 7
        N_REGIONS = args.regions
        regions = support.synth_regions(N_REGIONS, 7, 200, 100_000)
    else:
        regions = support.synth_regions(N_REGIONS, SEED_REGIONS)
    ctx = gtb200.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    engine = {"auto": 0, "rank": gtb200.ENGINE_RANK, "bucket": gtb200.ENGINE_BUCKET, "direct": gtb200.ENGINE_DIRECT, "enumerate": gtb200.ENGINE_ENUMERATE}[args.engine]

    # Genome shards: rank r owns a contiguous (chromosome, coordinate) range holding 1/world of the read mass and
    # every region whose span starts there.  P = the cut points on the axis of valid read starts.
    from gtb200 import sharded
    plan = sharded.ShardPlan(regions, world, chrom_extent=support.HG19_LENS)
    eff = np.maximum(support.HG19_LENS - READ_LEN + 1, 0)
    cum_eff = np.concatenate([[0], np.cumsum(eff)]).astype(np.int64)
    P = [0] + [int(cum_eff[c] + min(max(pos - 1, 0), int(eff[c]))) for c, pos in plan.cut_positions()] + [int(cum_eff[-1])]
    if strong:
        # the reads are SPLIT: a rank draws its share of the total inside its own range (first read index = what the ranks before it drew)
        share = [int(round(args.reads * (P[r + 1] - P[r]) / P[-1])) for r in range(world)]
        share[-1] = args.reads - sum(share[:-1])
        n, first_read = share[rank], sum(share[:rank])
        n_total = args.reads
        workload = "%d synthetic 50bp hg19 reads in total vs %d regions (200 bp - 100 kb), strand-aware count, genome-sharded over %d GPU(s) (configs[4])" % (n_total, N_REGIONS, world)
    else:
        n, first_read, n_total = args.reads, rank * args.reads, world * args.reads
        if args.regions != 60_000 or args.reads != N_READS:
            workload = "%d synthetic 50bp hg19 reads per GPU vs %d regions, strand-aware count" % (n, N_REGIONS)

    dev = {"chrom": torch.empty(n, dtype=torch.int32, device="cuda"), "start": torch.empty(n, dtype=torch.int32, device="cuda"),
           "stop": torch.empty(n, dtype=torch.int32, device="cuda"), "strand": torch.empty(n, dtype=torch.int8, device="cuda")}
    for lo in range(0, n, SLICE):
        part = {k: v[lo:lo + SLICE] for k, v in dev.items()}
        ctx.synth_reads(SEED_READS, first_read + lo, part["chrom"].numel(), READ_LEN, support.HG19_LENS, part,
                        p_range=(P[rank], P[rank + 1]) if world > 1 else None)
    out_dev = torch.zeros(N_REGIONS, dtype=torch.int64, device="cuda")

    def sets_of(t):
        """gtb_set views of a dict of device tensors, one per SLICE reads"""
        views, keep = [], []
        for lo in range(0, t["chrom"].numel(), SLICE):
            st, k = gtb200.device_set({kk: v[lo:lo + SLICE] for kk, v in t.items()})
            views.append(st); keep.append(k)
        return views, keep

    # ---- what the result must add up to, from an independent per-QUERY formulation (torch.searchsorted over the regions'
    # sorted starts / stops): sum_r count[r] == sum_q #{r in q's group: r.start <= q.stop} - #{r in q's group: r.stop < q.start}
    def dual_checksum(t):
        grp_r = torch.from_numpy(regions["chrom"].astype(np.int64) * 2 + (regions["strand"] == ord("-"))).cuda()
        ks = torch.sort((grp_r << 32) + torch.from_numpy(regions["start"].astype(np.int64)).cuda()).values
        ke = torch.sort((grp_r << 32) + torch.from_numpy(regions["stop"].astype(np.int64)).cuda()).values
        total = 0
        for lo in range(0, t["chrom"].numel(), 25_000_000):
            sl = slice(lo, lo + 25_000_000)
            g = (t["chrom"][sl].long() * 2 + (t["strand"][sl] == ord("-")).long()) << 32
            a = torch.searchsorted(ks, g + t["stop"][sl].long(), right=True) - torch.searchsorted(ks, g, right=False)
            bb = torch.searchsorted(ke, g + t["start"][sl].long(), right=False) - torch.searchsorted(ke, g, right=False)
            total += int((a - bb).sum().item())
        return total
    expect_sum = dual_checksum(dev)                                     # this rank's own reads (the halo below is other ranks')

    sharding = {}
    if world == 1:
        index = gtb200.Index(ctx, regions, gtb200.OP_COUNT, engine)
        dsets, keep = sets_of(dev)
        n_local = n

        # The library's kernels run on the context's own stream.  A step is queued without a host wait
        # (gtb_index_finish_async), ordered against torch's stream -- where the timing events and the consumers of the
        # result live -- by stream events; the engine's status is read once after the timed loop.
        lib_stream = torch.cuda.ExternalStream(ctx.stream_ptr())

        def step_device():
            lib_stream.wait_stream(torch.cuda.current_stream())
            index.reset()
            for st in dsets:
                index.add_set(st, gtb200.MEM_DEVICE)
            index.finish_async_ptr(out_dev.data_ptr())
            torch.cuda.current_stream().wait_stream(lib_stream)
            return out_dev
    else:
        # Reads that reach regions owned by a neighbour are replicated to it once, at ingest (untimed, like parsing);
        # a step then has no data-path exchange besides the final all-gather of per-region counts.
        send = {}
        for s_ in range(world):
            if s_ == rank:
                continue
            ids = torch.nonzero(sharded.route_mask_torch(plan, dev, s_)).flatten()
            send[s_] = {k: v[ids].cpu().numpy() for k, v in dev.items()}
        own_mask_ok = bool(sharded.route_mask_torch(plan, dev, rank).sum().item() <= n)
        gathered = [None] * world
        dist.all_gather_object(gathered, send)
        halo = [gathered[r][rank] for r in range(world) if r != rank and rank in gathered[r]]
        n_halo = sum(len(h["chrom"]) for h in halo)
        own = dev
        dev = {k: torch.cat([own[k]] + [torch.from_numpy(h[k]).cuda() for h in halo]) for k in own}
        n_local = n + n_halo
        dsets, keep = sets_of(dev)
        sh = sharded.ShardedDeviceOverlap(ctx, regions, plan, gtb200.OP_COUNT, engine)

        def step_device():
            # the host does not wait inside a step: kernels, all-gather and scatter are ordered by stream events, and the
            # engine's status is read once after the timed loop (sh.check() below)
            return sh.step(dsets, gtb200.MEM_DEVICE, defer_status=True)

        # cross-check (untimed): the query-sharded decomposition -- every rank counts its OWN reads against ALL regions,
        # one sum-reduction -- must give the same per-region counts as ownership + gather
        full = gtb200.Index(ctx, regions, gtb200.OP_COUNT, engine)
        own_sets, keep_own = sets_of(own)
        for st in own_sets:
            full.add_set(st, gtb200.MEM_DEVICE)
        full.finish_ptr(out_dev.data_ptr(), gtb200.MEM_DEVICE)
        dist.all_reduce(out_dev)
        got = step_device()
        sharding = {"decomposition": "genome ranges, region ownership, all-gather", "halo_reads_this_rank": n_halo,
                    "reads_this_rank": n, "owned_regions_this_rank": int(len(plan.owned[rank])),
                    "matches_query_sharded_allreduce": bool(torch.equal(got, out_dev)) and own_mask_ok}
        assert sharding["matches_query_sharded_allreduce"], "genome-sharded result differs from the query-sharded all-reduce"
        full.close()
        del own_sets, keep_own
        index = sh.index
        t = torch.tensor([expect_sum], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        expect_sum = int(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`) ----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ev1.record(stream)
        barrier()
    if world > 1:
        sh.check()
    else:
        index.status()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    counts_check = int(step_device().sum().item())
    if world > 1:
        sh.check()
    else:
        index.status()
    # a line is only printed for a correct result: the per-region counts must add up to the per-query formulation's total
    assert counts_check == expect_sum, "checksum %d differs from the independent per-query total %d" % (counts_check, expect_sum)

    # ---- per-kernel CUDA-event profile for the roofline (outside the timed region) -------------------
    ctx.profile(True)
    for _ in range(3):
        step_device()
    torch.cuda.synchronize()
    prof = ctx.profile_report()
    ctx.profile(False)
    dom_name, dom = max(prof.items(), key=lambda kv: kv[1]["total_ms"])
    dom_ms = dom["total_ms"] / dom["launches"]
    total_prof_ms = sum(v["total_ms"] for v in prof.values())
    peak, peak_kind = measured_peak()
    alg_bytes = BYTES_PER_QUERY * min(n_local, SLICE)    # one launch of the dominant kernel streams one batch (the whole of it at configs[1])
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of this workload (never measured
    # under the profiler here); only quoted when the batch size is the one that was captured
    traffic, traffic_src = None, None
    for name in ("r2_traffic.json", "r1g_traffic.json"):               # newest capture that knows this kernel
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tr = json.load(f)
            if tr.get("reads") == n_local and dom_name in tr:
                traffic = tr[dom_name]["dram_bytes_read"] + tr[dom_name]["dram_bytes_write"]
                traffic_src = tr["source"]
                break
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": dom_ms, "kernel_share_of_step": dom["total_ms"] / total_prof_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "step_frac": (BYTES_PER_QUERY * n_local + BYTES_PER_REGION * N_REGIONS) / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernels": {k: {"launches_per_step": v["launches"] / 3, "ms_per_launch": v["total_ms"] / v["launches"]} for k, v in prof.items()}}

    # ---- end to end through the C ABI with host buffers (`e2e`) -------------------------------------
    n_e2e = n
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms_per_step, "kernels": {k: round(v["ms_per_launch"], 4) for k, v in roofline["kernels"].items()},
                              "checksum": counts_check}))
        return
    if strong:
        # configs[4] is measured device-resident for the record (the driver's bench line is configs[1]): no host-buffer leg
        # (52 GB of pinned memory) and no CPU leg
        if rank == 0:
            line = {"metric": METRIC, "value": value, "unit": "query intervals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                    "config": {"workload": workload, "reads_total": n_total, "reads_this_rank": n, "n_regions": N_REGIONS, "read_len": READ_LEN,
                               "engine": args.engine, "l2": "inputs (%.1f GB per rank and step) far exceed the 126 MB L2; no flush needed" % (BYTES_PER_QUERY * n / 1e9),
                               "parallelism": "genome-sharded x%d: region ownership by range, boundary reads replicated at ingest, one NCCL all-gather of per-region counts" % world},
                    "roofline": roofline, "e2e": None, "gpu_launches": launches, "clocks": clocks.summary(), "checksum": counts_check,
                    "checksum_verified": "equals the per-query formulation's total (torch.searchsorted)", "sharding": sharding}
            print(json.dumps(line))
        (sh.close() if world > 1 else index.close())
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    # Host buffers in the two forms the ABI takes: the packed one a parser writes directly (gtb_index_add_packed: int32 start + one
    # byte of chromosome | strand per read, one read length -- 5 B/read) and the general SoA gtb_set (13 B/read, which the library
    # re-encodes to the packed form on the host before it crosses the link).  `e2e` is the packed form; `e2e_soa` the other.
    host = {k: torch.empty(n_local, dtype=v.dtype).pin_memory() for k, v in dev.items()}
    for k in host:
        host[k].copy_(dev[k])
    hset, keep2 = gtb200.pinned_set(host)
    host_meta = torch.empty(n_local, dtype=torch.uint8).pin_memory()
    host_meta.copy_((dev["chrom"] | ((dev["strand"] == ord("-")).int() << 7)).to(torch.uint8))
    packed = {"n": n_local, "start": host["start"].data_ptr(), "meta": host_meta.data_ptr(), "read_len": READ_LEN}
    out_host = torch.zeros(N_REGIONS, dtype=torch.int64).pin_memory()

    def step_e2e(form):
        if world == 1:
            index.reset()
            if form == "packed":
                index.add_packed_ptr(packed["n"], packed["start"], packed["meta"], packed["read_len"], gtb200.MEM_HOST)
            else:
                index.add_set(hset, gtb200.MEM_HOST)
            index.finish_ptr(out_host.data_ptr(), gtb200.MEM_HOST)
        else:
            out_host.copy_(sh.step(packed if form == "packed" else hset, gtb200.MEM_HOST))   # H2D of the local reads, gather, D2H of all counts

    def time_e2e(form, steps):
        step_e2e(form)
        step_e2e(form)                                                   # (two untimed steps: staging buffers allocated, the host packing pool's rate known)
        assert int(out_host.sum().item()) == expect_sum, "e2e (%s) result differs" % form
        barrier()
        xfer0 = ctx.transfer_stats()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e(form)
        barrier()
        secs = (time.perf_counter() - t0) / steps
        if dist is not None:
            t = torch.tensor([secs], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        xfer1 = ctx.transfer_stats()
        # bytes that actually crossed PCIe per step on this rank; host_buffer_bytes is what the caller handed over
        return {"value": n_total / secs, "unit": "query intervals/s",
                "h2d_bytes_per_step": (xfer1["h2d_bytes"] - xfer0["h2d_bytes"]) // steps,
                "d2h_bytes_per_step": (xfer1["d2h_bytes"] - xfer0["d2h_bytes"]) // steps + (8 * N_REGIONS if world > 1 else 0),
                "host_buffer_bytes_per_step": (5 if form == "packed" else BYTES_PER_QUERY) * n_local,
                "host_form": "packed: int32 start + uint8 chromosome|strand, one read length (gtb_index_add_packed)" if form == "packed"
                             else "gtb_set SoA: int32 chrom/start/stop + int8 strand (gtb_index_add_queries)",
                "packed_chunks": xfer1["packed_chunks"] - xfer0["packed_chunks"], "raw_chunks": xfer1["raw_chunks"] - xfer0["raw_chunks"],
                "ms_per_step": secs * 1e3}

    e2e_steps = max(2, min(args.steps, 5))
    e2e = time_e2e("packed", e2e_steps)
    e2e_soa = time_e2e("soa", max(2, e2e_steps // 2))

    # ---- the same reads in position order (what -S promises and aligners emit), device-resident: a secondary figure next to `value`
    sorted_input = None
    if world == 1:
        key = (dev["chrom"].long() << 33) | ((dev["strand"] == ord("-")).long() << 32) | dev["start"].long()
        order = torch.argsort(key)
        del key
        sdev = {k: v[order].contiguous() for k, v in dev.items()}
        del order
        ssets, keep3 = sets_of(sdev)

        def step_sorted():
            lib_stream.wait_stream(torch.cuda.current_stream())
            index.reset()
            for st in ssets:
                index.add_set(st, gtb200.MEM_DEVICE)
            index.finish_async_ptr(out_dev.data_ptr())
            torch.cuda.current_stream().wait_stream(lib_stream)
        for _ in range(args.warmup):
            step_sorted()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(args.steps):
            step_sorted()
        s1.record(stream)
        torch.cuda.synchronize()
        index.status()
        assert int(out_dev.sum().item()) == expect_sum, "sorted-input result differs"
        s_ms = s0.elapsed_time(s1) / args.steps
        ctx.profile(True)
        for _ in range(3):
            step_sorted()
        torch.cuda.synchronize()
        sprof = ctx.profile_report()
        ctx.profile(False)
        s_dom = sprof[dom_name]["total_ms"] / sprof[dom_name]["launches"] if dom_name in sprof else None
        sorted_input = {"order": "chromosome, strand, start", "ms_per_step": s_ms, "value": n / (s_ms * 1e-3),
                        "step_frac": (BYTES_PER_QUERY * n + BYTES_PER_REGION * N_REGIONS) / (s_ms * 1e-3) / 1e9 / peak,
                        "kernel": dom_name, "kernel_ms": s_dom, "frac": None if s_dom is None else alg_bytes / (s_dom * 1e-3) / 1e9 / peak}
        del sdev, ssets, keep3

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "query intervals/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": {"workload": workload,
                           "reads_per_gpu": n, "n_regions": N_REGIONS, "read_len": READ_LEN, "engine": args.engine,
                           "l2": "inputs (%.1f GB per step) far exceed the 126 MB L2; no flush needed" % (BYTES_PER_QUERY * n / 1e9),
                           "parallelism": "genome-sharded x%d: region ownership by range, boundary reads replicated at ingest, one NCCL "
                                          "all-gather of per-region counts" % world if world > 1 else "single GPU"},
                "roofline": roofline, "e2e": e2e, "e2e_soa": e2e_soa, "gpu_launches": launches, "clocks": clocks.summary(),
                "checksum": counts_check, "checksum_verified": "equals the per-query formulation's total (torch.searchsorted)"}
        if sorted_input:
            line["sorted_input"] = sorted_input
        if sharding:
            line["sharding"] = sharding
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(regions)
        print(json.dumps(line))
    if world == 1:
        index.close()
    else:
        sh.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
