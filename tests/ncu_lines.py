"""Per CUDA source line of an .ncu-rep: instructions, shared-memory wavefronts (actual / ideal), stall samples."""
import csv
import subprocess
import sys


def main(path, kernel, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    ci = {}
    for i, n in enumerate(hdr):
        ci.setdefault(n, i)
    tab = []
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        f = lambda n: float(r[ci[n]] or 0) if n in ci and r[ci[n]] not in ("", "-") else 0.0
        try:                                          # source lines with quotes and commas (inline asm) do not survive the CSV export
            f("Instructions Executed"), f("# Samples"), f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Ideal"), f("stall_short_sb")
            f("stall_long_sb"), f("stall_barrier"), f("stall_wait"), f("stall_mio"), f("stall_selected"), f("L2 Theoretical Sectors Global")
        except ValueError:
            continue
        tab.append(dict(line=int(r[0]), src=r[1].strip()[:90], inst=f("Instructions Executed"), smp=f("# Samples"), wf=f("L1 Wavefronts Shared"),
                        wfi=f("L1 Wavefronts Shared Ideal"), ssb=f("stall_short_sb"), lsb=f("stall_long_sb"), bar=f("stall_barrier"),
                        wait=f("stall_wait"), mio=f("stall_mio"), sel=f("stall_selected"), g=f("L2 Theoretical Sectors Global")))
    T = {k: sum(t[k] for t in tab) or 1 for k in ("inst", "smp", "wf", "wfi", "ssb", "lsb", "bar", "wait", "mio", "g")}
    print("totals: warp-inst %.3g  samples %.0f  shared wavefronts %.3g (ideal %.3g)  global sectors %.3g" % (T["inst"], T["smp"], T["wf"], T["wfi"], T["g"]))
    print("stall samples: short_sb %.0f long_sb %.0f barrier %.0f wait %.0f mio %.0f" % (T["ssb"], T["lsb"], T["bar"], T["wait"], T["mio"]))
    for t in sorted(tab, key=lambda t: -t["smp"])[:top]:
        print("L%-4d smp %4.1f%% inst %4.1f%% wf %4.1f%% (x%.1f of ideal) ssb %4.1f%% bar %4.1f%% lsb %4.1f%% | %s" % (
            t["line"], 100 * t["smp"] / T["smp"], 100 * t["inst"] / T["inst"], 100 * t["wf"] / T["wf"], t["wf"] / t["wfi"] if t["wfi"] else 0,
            100 * t["ssb"] / T["ssb"], 100 * t["bar"] / T["bar"], 100 * t["lsb"] / T["lsb"], t["src"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
