"""Turn an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of
the bench command into profiles/ncu_traffic.json (read back by bench.py for `roofline.traffic`).
    python tools/ncu_traffic.py gpurun_out/traffic.csv products 1.0 1 <nnz> <d> [kind [kernel-name substring]]"""
import csv, io, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, workload, scale, world, nnz, d = sys.argv[1], sys.argv[2], float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
kind = sys.argv[7] if len(sys.argv) > 7 else "spmm"
names = [sys.argv[8]] if len(sys.argv) > 8 else ["spmm_bulk_kernel", "spmm_vec_kernel"]
txt = open(path).read().splitlines()
start = [i for i, l in enumerate(txt) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(io.StringIO("\n".join(txt[start:]))))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
per = {}
for r in rows:
    if not any(nm in r["Kernel Name"] for nm in names):
        continue
    e = per.setdefault(r["ID"], {})
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"].startswith("dram__bytes"):
        e[r["Metric Name"]] = v * UNIT[r["Metric Unit"]]
    elif r["Metric Name"] == "gpu__time_duration.sum":
        e["ms"] = v * {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
# keep the launches of the widest group (longest ones)
launches = sorted(per.values(), key=lambda e: -e.get("ms", 0))
top = [e for e in launches if e.get("ms", 0) > 0.8 * launches[0]["ms"]]
tot = sum(e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"] for e in top) / len(top)
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[f"{workload}:{scale}:{world}:{kind} d={d}"] = {
    "nnz": nnz, "dram_bytes_per_launch": tot, "launches_averaged": len(top),
    "avg_ms_under_ncu": sum(e["ms"] for e in top) / len(top),
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline`"}
json.dump(data, open(out_path, "w"), indent=1)
print(json.dumps(data, indent=1))
