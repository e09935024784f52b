"""Dynamic SASS mix of one kernel of an ncu report: executed warp instructions per unit by opcode, shared-memory
wavefronts, and where the stall samples sit.

    python tools/sass_mix.py report.ncu-rep <kernel-substring> <units> [--instance k] [--dump]

`units` = frames (or whatever the kernel's unit of work is) per launch; --dump prints every hot instruction
(executed at least units/8 times) with its count and samples, in address order.
"""
import csv
import subprocess
import sys
from collections import defaultdict


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    kernels, cur, hdr = [], None, None
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r[0] == "Address":
            hdr = r
        elif cur is not None and r[0].startswith("0x"):
            cur["rows"].append(dict(zip(hdr, r)))
    return kernels


def num(s):
    try:
        return int(s)
    except (ValueError, TypeError):
        return 0


def main():
    rep, pat, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
    inst = int(sys.argv[sys.argv.index("--instance") + 1]) if "--instance" in sys.argv else 0
    ks = [k for k in load(rep) if pat in k["name"]]
    k = ks[inst]
    print(k["name"][:140])
    ops, wf, wf_ideal, samples = defaultdict(float), defaultdict(float), defaultdict(float), defaultdict(int)
    tot = 0
    stall_cols = [c for c in k["rows"][0] if c.startswith("stall_") and "Not Issued" not in c]
    stalls = defaultdict(int)
    for r in k["rows"]:
        src = r["Source"].strip()
        parts = src.split()
        op = parts[0]
        if op.startswith("@"):
            op = parts[1]
        op = op.rstrip(";")
        n = num(r["Instructions Executed"])
        tot += n
        ops[op] += n
        base = op.split(".")[0]
        w = num(r.get("L1 Wavefronts Shared"))
        if w:
            wf[base + "." + r.get("Access Size", "")] += w
            wf_ideal[base + "." + r.get("Access Size", "")] += num(r.get("L1 Wavefronts Shared Ideal"))
        samples[base] += num(r["# Samples"])
        for c in stall_cols:
            stalls[c] += num(r[c])
    print("instructions / unit: %.1f" % (tot / units))
    byb = defaultdict(float)
    for op, n in ops.items():
        byb[op.split(".")[0]] += n
    for op, n in sorted(byb.items(), key=lambda x: -x[1])[:40]:
        print("  %-10s %8.1f   samples %5.1f%%" % (op, n / units, 100.0 * samples[op] / max(1, sum(samples.values()))))
    print("shared wavefronts / unit (actual, ideal):")
    for key in sorted(wf, key=lambda x: -wf[x]):
        print("  %-14s %8.1f %8.1f" % (key, wf[key] / units, wf_ideal[key] / units))
    print("  total          %8.1f %8.1f" % (sum(wf.values()) / units, sum(wf_ideal.values()) / units))
    ts = sum(stalls.values())
    print("stall samples:", ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / ts) for c, v in sorted(stalls.items(), key=lambda x: -x[1])[:9]))
    if "--dump" in sys.argv:
        for r in k["rows"]:
            n = num(r["Instructions Executed"])
            if n >= units / 8:
                print("%s %9.2f %5d  wf %6.2f/%6.2f  %s" % (r["Address"][-5:], n / units, num(r["# Samples"]),
                                                          num(r.get("L1 Wavefronts Shared")) / units,
                                                          num(r.get("L1 Wavefronts Shared Ideal")) / units, r["Source"].strip()))


if __name__ == "__main__":
    main()
