"""Per-source-line totals of an .ncu-rep captured with --import-source on (-lineinfo build): warp instructions executed,
average active threads, stall samples and excess shared-memory wavefronts per source line (innermost inlined location).
usage: python tools/ncu_lines.py report.ncu-rep [top=40]"""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []; cur = None; hdr = None
for rec in csv.reader(io.StringIO(out)):
    if not rec: continue
    if rec[0] == "File Path": cur = rec[1].split("/")[-1]; continue
    if rec[0] == "Function Name": continue
    if rec[0] == "Line No": hdr = rec; continue
    if hdr and rec[0] != "" and cur:
        if not rec[0].isdigit(): continue
        extra = len(rec) - len(hdr)                      # unescaped quotes in the source text split it into more fields
        if extra > 0: rec = [rec[0], ",".join(rec[1:2 + extra])] + rec[2 + extra:]
        d = dict(zip(hdr, rec))
        # the header has two "Source" columns (source text, sass text): csv dict keeps the last; take the text by position
        rows.append((cur, int(rec[0]), rec[1], int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0),
                     float(d["Avg. Threads Executed"] or 0), int(d.get("L1 Wavefronts Shared Excessive", 0) or 0), int(d.get("L1 Wavefronts Shared", 0) or 0)))
tot_i = sum(r[3] for r in rows); tot_s = sum(r[4] for r in rows)
print(f"{path}: {tot_i} warp instructions, {tot_s} samples, shared wavefronts {sum(r[7] for r in rows)} (excess {sum(r[6] for r in rows)})")
byfile = {}
for r in rows: byfile[r[0]] = byfile.get(r[0], 0) + r[3]
print("by file:", ", ".join(f"{k} {100*v/tot_i:.1f}%" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])))
print(f"{'file':<22}{'line':>5} {'instr%':>7} {'smpl%':>6} {'thr':>5} {'wfX':>9}  source")
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print(f"{r[0]:<22}{r[1]:>5} {100*r[3]/tot_i:7.2f} {100*r[4]/max(tot_s,1):6.2f} {r[5]:5.1f} {r[6]:>9}  {r[2].strip()[:110]}")
