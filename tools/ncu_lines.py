"""Stall samples of one kernel per CUDA source line, from an .ncu-rep captured with --set full --import-source on
(compile with -lineinfo):   python tools/ncu_lines.py gpurun_out/x.ncu-rep [top=30]
Reads `ncu -i rep --page source --print-source cuda,sass --csv`; rows with a line number are the per-line sums."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
ci = {}
for i, n in enumerate(hdr):
    ci.setdefault(n, i)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
lines, fname = [], ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    try:
        s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    lines.append((s, fname, r))
tot = sum(s for s, _, _ in lines)
print(f"{rep}: {tot} samples on {len(lines)} source lines")
for s, f, r in sorted(lines, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[ci[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
    print(f"{100 * s / tot:5.1f}%  {f}:{r[0]:>4}  {r[1].strip()[:100]:100s}  " + " ".join(f"{n}={v}" for v, n in st if v))
