"""Top CUDA source lines of an .ncu-rep by instructions executed and by sampled stall reason."""
import csv
import subprocess
import sys


def main(path, top=22):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    ci = {n: i for i, n in enumerate(hdr)}
    tab = []
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[0] == "Line No":
            break
        def num(name):
            try:
                return float(r[ci[name]] or 0)
            except Exception:
                return 0.0
        tab.append(dict(line=r[0], src=r[1].strip()[:100], inst=num("Instructions Executed"), smp=num("# Samples"),
                        long_sb=num("stall_long_sb"), short_sb=num("stall_short_sb"), wait=num("stall_wait"), barrier=num("stall_barrier"),
                        branch=num("stall_branch_resolving"), tinst=num("Thread Instructions Executed")))
    ti = sum(t["inst"] for t in tab); ts = sum(t["smp"] for t in tab)
    print("total warp-inst %.0f  samples %.0f" % (ti, ts))
    for key in ("inst", "smp", "long_sb"):
        print("--- top by", key)
        tot = sum(t[key] for t in tab) or 1
        for t in sorted(tab, key=lambda t: -t[key])[:top]:
            print("%5.1f%%  L%-4s inst=%4.1f%% smp=%4.1f%% thr/inst=%4.1f | %s" % (100 * t[key] / tot, t["line"], 100 * t["inst"] / ti, 100 * t["smp"] / ts,
                                                                                 t["tinst"] / t["inst"] if t["inst"] else 0, t["src"]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
