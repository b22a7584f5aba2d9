#!/usr/bin/env python
"""End-to-end timing of the drop-in command-line drivers against the reference binaries: text file in, text out.

    python bench_cli.py [--reads N] [--regions M] [--ref-reads K] [--out FILE]

Writes N synthetic hg19 reads (BED6, the stream of DESIGN.md section 6) and M regions to a scratch directory, then runs
  * oracle/_ref/genomic_overlaps count / genomic_scans counts   (the unmodified reference, single-threaded as shipped) on the
    first K reads -- the reference parses about 2 M reads per second, so K is kept small and the rate is what is compared;
  * ibm-cbc-genomic-tools_b200/bin/genomic_overlaps count / coverage and genomic_scans counts on all N reads,
    and on the first K reads, whose stdout must equal the reference's byte for byte.
Wall-clock of the whole process (CUDA context creation, parsing, transfers, kernels, printing).  One JSON line per case.
Secondary measurement; the headline numbers are bench.py's."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import support  # noqa: E402

BIN = os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


LAST_PHASES = {}


def timed(cmd, stdout_path):
    """wall-clock of the whole process; our drivers also report their phases on stderr (GT_TIMING)"""
    t0 = time.perf_counter()
    with open(stdout_path, "wb") as f:
        p = subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE, env=dict(os.environ, GT_TIMING="1"))
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("%s failed: %s" % (cmd, p.stderr.decode()[-400:]))
    LAST_PHASES.clear()
    for line in p.stderr.decode().splitlines():
        if line.startswith("[gt timing]"):
            for tok in line.split()[2:]:
                k, v = tok.split("=")
                LAST_PHASES[k] = float(v.rstrip("s"))
    return dt


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--regions", type=int, default=60_000)
    ap.add_argument("--ref-reads", type=int, default=10_000_000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--tmp", default=None)
    args = ap.parse_args()
    tmp = args.tmp or tempfile.mkdtemp(prefix="gt_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    os.makedirs(tmp, exist_ok=True)
    reads = os.path.join(tmp, "reads.bed")
    sub = os.path.join(tmp, "reads_sub.bed")
    t0 = time.perf_counter()
    with open(reads, "wb") as f:
        subprocess.run([os.path.join(BIN, "gt_synth_bed"), str(args.reads), "2"], stdout=f, check=True)
    with open(sub, "wb") as f:
        subprocess.run([os.path.join(BIN, "gt_synth_bed"), str(args.ref_reads), "2"], stdout=f, check=True)
    regions = os.path.join(tmp, "regions.bed")
    support.write_bed(regions, support.synth_regions(args.regions, 3), support.HG19_NAMES, labels=["g%06d" % k for k in range(args.regions)])
    genome = os.path.join(tmp, "genome.bed")
    with open(genome, "w") as f:
        for name, ln in zip(support.HG19_NAMES, support.HG19_LENS):
            f.write("%s\t0\t%d\n" % (name, ln))
    lines = []

    def emit(**kw):
        kw["host_cores"] = os.cpu_count()
        lines.append(kw)
        print(json.dumps(kw), flush=True)

    emit(case="inputs", reads=args.reads, reads_bytes=os.path.getsize(reads), regions=args.regions, ref_reads=args.ref_reads,
         seconds_to_write=round(time.perf_counter() - t0, 2), tmp=tmp)
    have_ref = os.path.exists(os.path.join(REF, "genomic_overlaps"))
    cases = [("genomic_overlaps", ["count"], [regions]), ("genomic_overlaps", ["coverage"], [regions]),
             ("genomic_scans", ["counts", "-g", genome, "-w", "200", "-d", "50", "-min", "10"], [])]
    for tool, opts, pre in cases:
        name = tool + " " + opts[0]
        out_ref, out_sub, out_new = (os.path.join(tmp, x) for x in ("ref.out", "sub.out", "new.out"))
        t_ref = None
        if have_ref:
            t_ref = timed([os.path.join(REF, tool)] + opts + pre + [sub], out_ref)
        timed([os.path.join(BIN, tool)] + opts + pre + [sub], out_sub)          # also warms the page cache and the driver
        t_sub = timed([os.path.join(BIN, tool)] + opts + pre + [sub], out_sub)
        phases_sub = dict(LAST_PHASES)
        t_new = timed([os.path.join(BIN, tool)] + opts + pre + [reads], out_new)
        phases_new = dict(LAST_PHASES)
        stream = phases_new.get("stream_queries", phases_new.get("stream_reads"))
        same = (md5(out_ref) == md5(out_sub)) if have_ref else None
        emit(case=name, reference_seconds=t_ref, reference_reads=args.ref_reads,
             reference_reads_per_s=(args.ref_reads / t_ref) if t_ref else None,
             ours_seconds_same_input=round(t_sub, 3), stdout_identical=same,
             ours_seconds=round(t_new, 3), ours_reads=args.reads, ours_reads_per_s=args.reads / t_new,
             ours_text_GBps=os.path.getsize(reads) / t_new / 1e9,
             ours_phases_same_input=phases_sub, ours_phases=phases_new,
             ours_stream_reads_per_s=(args.reads / stream) if stream else None,
             speedup_same_input=(t_ref / t_sub) if t_ref else None,
             speedup_rate=(args.reads / t_new) / (args.ref_reads / t_ref) if t_ref else None)
        if have_ref and not same:
            raise SystemExit("stdout differs from the reference for " + name)
    if args.out:
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")
    for fn in ("reads.bed", "reads_sub.bed", "ref.out", "sub.out", "new.out", "regions.bed", "genome.bed"):
        try:
            os.remove(os.path.join(tmp, fn))
        except OSError:
            pass


if __name__ == "__main__":
    main()
