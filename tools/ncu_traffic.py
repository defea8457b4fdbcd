"""Write profiles/ncu_traffic.json (DRAM bytes per launch of the decode and fast-encode kernels) from the two
`ncu --set full` captures tools/gpu_round.sh makes.  usage: python tools/ncu_traffic.py TAG BLOCKS"""
import csv, io, json, subprocess, sys
tag, blocks = sys.argv[1], int(sys.argv[2])
out = {"source": f"ncu --set full --clock-control none, bench.py --blocks {blocks} (profiles/ncu_{tag}_decode_encode.md), class E50", "blocks": blocks}
for key, rep in (("lz4_decode_kernel", f"gpurun_out/dec_{tag}.ncu-rep"), ("lz4_encode_fast_kernel", f"gpurun_out/enc_{tag}.ncu-rep")):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tot = sum(float(d[m]) * scale[u[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out[key] = {"dram_bytes": tot, "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "s": 1e3}[u["gpu__time_duration.sum"]]}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out))
