"""ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) -> per-kernel summary of the LAST
run in the capture + profiles/traffic.json entry.  usage: mk_profile.py in.csv out_summary.csv workload 'command'"""
import csv, sys, json, collections
from pathlib import Path
src, out, workload, command = sys.argv[1:5]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = collections.OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(r[ii], {"name": r[ki]})
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        d["us"] = v / 1000 if r[ui] in ("ns", "nsecond") else (v * 1000 if r[ui] in ("ms", "msecond") else v)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
        d["dram"] = d.get("dram", 0.0) + v * scale
L = list(launches.values())
starts = [i for i, d in enumerate(L) if "state_reset_kernel" in d["name"]]
run = L[starts[-1]:] if starts else L
agg = collections.OrderedDict()
for d in run:
    name = d["name"].replace("pprb200::", "").replace("void ", "").split("(")[0]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("dram", 0.0)
tot = sum(a[1] for a in agg.values())
with open(out, "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none :\n#   {command}\n")
    f.write("# last run of the command; per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"# total {tot/1000:.2f} ms over {sum(a[0] for a in agg.values())} launches\nkernel,launches,total_us,share,dram_bytes\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{a[0]},{a[1]:.1f},{a[1]/tot:.4f},{a[2]:.0f}\n")
merge = {k: a for k, a in agg.items() if k.startswith("merge_") or k.startswith("mc_walk")}
tj = Path("profiles/traffic.json")
t = json.loads(tj.read_text()) if tj.exists() else {}
t[workload] = {"command": command, "kernel": "merge_dense_kernel + merge_seq_kernel + merge_par_kernel" + (" + mc_walk_kernel" if any(k.startswith("mc_walk") for k in merge) else "") + ", all launches of one run",
               "dram_bytes_per_step": sum(a[2] for a in merge.values()), "ncu_merge_us": sum(a[1] for a in merge.values()),
               "source": f"{out} (ncu dram__bytes_read.sum + dram__bytes_write.sum)"}
tj.write_text(json.dumps(t, indent=1))
print(open(out).read())
