"""Condense gpurun_out/ ncu outputs into small tracked files under profiles/.
    python tools/summarise_profiles.py <round-tag> [launches.csv] [prof.ncu-rep]"""
import collections, csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof_r1.ncu-rep")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

UNIT = {"usecond": 1e-3, "us": 1e-3, "nsecond": 1e-6, "ns": 1e-6, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}

if os.path.exists(launches):
    txt = open(launches).read().splitlines()
    start = [i for i, l in enumerate(txt) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(txt[start:]))))
    agg, ours = collections.OrderedDict(), []
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        ms = float(r["Metric Value"].replace(",", "")) * UNIT[r["Metric Unit"]]
        short = name.split("(")[0].replace("void ", "")
        short = short if "lgnn::" in short else "lib: " + short.split("<")[0][-48:]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ms
        if "lgnn::" in name:
            ours.append((r["ID"], short, r["Grid Size"], r["Block Size"], ms))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {len(rows)} launches, {tot:.2f} ms total\n")
        f.write("# (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write(f"{'ms':>10s} {'share':>6s} {'n':>5s}  kernel\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{a[1]:10.3f} {100 * a[1] / tot:5.1f}% {a[0]:5d}  {k}\n")
    with open(os.path.join(out_dir, f"{tag}_launches_lgnn.csv"), "w") as f:
        f.write("id,kernel,grid,block,ms\n")
        for o in ours:
            f.write(",".join(str(v) for v in o) + "\n")

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(out_dir, f"{tag}_ncu_full_summary.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}), one block per captured launch\n")
        for r in rows[2:]:
            f.write("----\n")
            for k in keep:
                if k in idx:
                    f.write(f"{k} [{units[idx[k]]}] = {r[idx[k]]}\n")
print("written to", out_dir)
