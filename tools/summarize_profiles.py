"""Distil gpurun_out/ ncu captures into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep>[,<more.ncu-rep>...] [bench json ...]
"""
import csv, json, os, subprocess, sys
from collections import defaultdict

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)


def short(name):
    name = name.replace("acids::", "").replace("(int)", "")
    return name[:110]


# ---- launch list: per-kernel totals and shares (cold-cache, serialised: compare SHARES) ----
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
h = next(r for r in rows if r[0] == "ID")
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[rows.index(h) + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    tot[short(r[ki])] += v
    cnt[short(r[ki])] += 1
s = sum(tot.values())
with open(os.path.join(OUT, "%s_launches.csv" % tag), "w") as f:
    f.write("kernel,launches,total_us,share_pct\n")
    for n, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write('"%s",%d,%.1f,%.2f\n' % (n, cnt[n], v, 100 * v / s))

# ---- full capture: key metrics per captured launch ----
rr, row_units = [], []
for one in rep.split(","):
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if not rr:
        rr = part[:2]
    idx = [part[0].index(c) if c in part[0] else -1 for c in rr[0]]
    for row in part[2:]:
        rr.append([row[i] if i >= 0 else "" for i in idx])
        row_units.append([part[1][i] if i >= 0 else "" for i in idx])      # units are per report (Mbyte here, Gbyte there)
hh = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
stalls = [w for w in hh if "issue_stalled" in w and "per_issue_active" in w and "not_issued" not in w]
traffic = {}
with open(os.path.join(OUT, "%s_ncu_summary.md" % tag), "w") as f:
    f.write("# ncu --set full summary (%s)\n\nSource: `%s` (scratch, not tracked).  One block per captured launch; byte\n"
            "and time units as printed by ncu.  Stall columns are warps stalled per issued instruction.\n\n" % (tag, rep))
    for r, units in zip(rr[2:], row_units):
        name = short(r[hh.index("Kernel Name")])
        f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name)
        vals = {}
        for w in want:
            if w in hh:
                vals[w] = r[hh.index(w)]
                f.write("| %s | %s | %s |\n" % (w, r[hh.index(w)], units[hh.index(w)]))
        st = sorted(((float(r[hh.index(w)] or 0), w) for w in stalls), reverse=True)[:8]
        for v, w in st:
            f.write("| %s | %.3f | warps/issue |\n" % (w.replace("smsp__average_warps_issue_stalled_", "stall: ").replace("_per_issue_active.ratio", ""), v))
        f.write("\n")
        try:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(vals["dram__bytes_read.sum"]) * mult[units[hh.index("dram__bytes_read.sum")]]
            wr = float(vals["dram__bytes_write.sum"]) * mult[units[hh.index("dram__bytes_write.sum")]]
            traffic.setdefault(name, []).append(rd + wr)
        except (KeyError, ValueError):
            pass
# the bench kernels are the LARGEST launch of each name
tj = {}
for n, v in traffic.items():
    key = "stft_fwd_kernel<Fwd1024R,MODE_REAL>" if ("stft_fwd_kernel<Plan<1024" in n and ">, 1, " in n) else \
          "istft_ola_kernel<Inv1024>" if "istft_ola_kernel<Plan<1024" in n else n
    tj[key] = max(tj.get(key, 0), max(v))
json.dump(tj, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
for i, src in enumerate(sys.argv[4:]):
    dst = os.path.join(OUT, "%s_%s" % (tag, os.path.basename(src)))
    lines = [l for l in open(src).read().splitlines() if l.startswith("{")]
    open(dst, "w").write("\n".join(lines) + "\n")
print("wrote", sorted(os.listdir(OUT)))
