"""The instructions that collect the most warp-stall samples, per captured launch, from an .ncu-rep (source page), plus the
stall-reason totals of the raw page.     python tools/ncu_hot_lines.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sections, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
for n, sec in enumerate(sections):
    h = sec["hdr"]
    try:
        si, ii = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    except ValueError:
        continue
    rows = [r for r in sec["rows"] if len(r) > max(si, ii) and r[si].isdigit()]
    tot = sum(int(r[si]) for r in rows) or 1
    inst = sum(int(r[ii]) for r in rows if r[ii].isdigit())
    print(f"## launch {n} hot instructions ({tot} stall samples, {inst} warp instructions executed)")
    for r in sorted(rows, key=lambda r: -int(r[si]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
        print(f"   {100 * int(r[si]) / tot:5.1f} %  x{r[ii]:>9}  {r[1].strip()[:110]}")
    if len(raw) > 2 + n:
        d = dict(zip(raw[0], raw[2 + n]))
        st = sorted(((k.replace("smsp__pcsamp_warps_issue_stalled_", ""), int(float(v))) for k, v in d.items()
                     if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "n/a")),
                    key=lambda kv: -kv[1])
        print("   stall reasons: " + ", ".join(f"{k} {v}" for k, v in st[:6]))
        for k in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
                  "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"):
            if k in d:
                print(f"   {k} {d[k]}")
    print()
