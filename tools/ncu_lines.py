"""Per-source-line instruction counts / stall samples from an ncu report (needs -lineinfo at compile time).

    python tools/ncu_lines.py report.ncu-rep [kernel-substring] [frames]
"""
import csv, subprocess, sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
out, cur, kern, hdr, seen = {}, None, None, None, []
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        kern = r[1]
        if kern not in seen:
            seen.append(kern)
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and pat in kern:
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        def num(s):
            try:
                return int(s)
            except ValueError:
                return 0
        key = (kern, cur, int(r[0]), r[1].strip()[:100])
        v = out.setdefault(key, [0, 0])
        v[0] += num(r[ie]); v[1] += num(r[ss])
for kk in [k for k in seen if pat in k]:
    items = [(k, v) for k, v in out.items() if k[0] == kk]
    tot, tots = sum(v[0] for _, v in items), sum(v[1] for _, v in items)
    print(kk[:110], "| inst/unit %.1f | samples %d" % (tot / units, tots))
    for k, v in sorted(items, key=lambda x: -x[1][0])[:int(__import__("os").environ.get("TOP", "40"))]:
        print("%8.1f inst  %5.1f%% samples  %-16s:%-4d %s" % (v[0] / units, 100.0 * v[1] / max(tots, 1), k[1], k[2], k[3]))
    break
