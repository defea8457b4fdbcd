"""Regenerates tests/golden/golden_v1.json from the REFERENCE's own code (oracle/_ref/liblz4net_ref.so, i.e.
/root/reference/original/lz4.c + lz4hc.c compiled in place by oracle/Makefile).  Run only in the build container:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these are outputs of the reference itself on
reproducible inputs (tests/cases.py), including the reference-defined AutoTest text (src/LZ4/LZ4Codec.cs:175-184).
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests import cases  # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    out = []
    todo = [("autotest", "autotest", len(cases.AUTOTEST), 0)]
    for m in cases.MODELS:
        todo.append((f"{m}-64k", m, 65536, 1))
    for i, n in enumerate((0, 1, 12, 13, 20, 64, 300, 4096, 32768, 65535, 65546, 65547, 100000)):
        for m in ("mixed", "lowent", "ETEXT", "periodic"):
            todo.append((f"{m}-{n}", m, n if not (m == 'mixed' and n > 70000) else 70000, 50 + i))
    seen = set()
    for name, model, n, seed in todo:
        if name in seen:
            continue
        seen.add(name)
        data = cases.AUTOTEST if model == "autotest" else cases.content(model, n, seed).tobytes()
        c = {"name": name, "model": model, "n": len(data), "seed": seed, "input_sha256": sha(data)}
        for mode, fn in (("fast", oracle.encode), ("hc", oracle.encode_hc)):
            r, o = fn(data, impl="ref")
            r2, _ = fn(data, cap=len(data), impl="ref")
            c[mode] = {"len": r, "sha256": sha(o), "len_cap_n": r2}
            if r <= 512:
                c[mode]["hex"] = o.hex()
        out.append(c)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")
    json.dump({"source": "oracle/_ref (reference original/lz4.c + lz4hc.c, LZ4_ARCH64=1, LZ4_MK_OPT)", "cases": out},
              open(path, "w"), indent=1)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
