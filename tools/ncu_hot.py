"""Hot spots of an .ncu-rep captured with --import-source on: SASS instructions ranked by warp-stall samples.
usage: python tools/ncu_hot.py report.ncu-rep [top=40] [context=0]"""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"]) for r in rows)
inst = sum(int(r["Instructions Executed"]) for r in rows)
print(f"{lines[0][:150]}\ntotal samples {tot}, warp instructions {inst}, SASS lines {len(rows)}")
reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[k]) for r in rows) for k in reasons}
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 > tot))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"]))[:top]
print(f"{'#':>5} {'samples':>8} {'%':>5} {'exec':>10}  top reason      SASS")
for i in sorted(idx):
    r = rows[i]; s = int(r["# Samples"])
    rs = max(reasons, key=lambda k: int(r[k]))
    print(f"{i:5d} {s:8d} {100*s/tot:5.1f} {int(r['Instructions Executed']):10d}  {rs[6:]:<14} {r['Source'].strip()}")
