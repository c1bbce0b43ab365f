"""Matrix view (rows n, columns d) of one field of a fullbench JSONL file: python tools/fbtable.py file [field]"""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1])]
field = sys.argv[2] if len(sys.argv) > 2 else "roofline_frac"
ns = sorted({r["n"] for r in rows}); ds = sorted({r["d"] for r in rows})
tab = {(r["n"], r["d"]): r for r in rows}
code = {"tiny": "t", "pairtile": "p", "pairtile-multipass": "P", "generic": "g", "generic-multipass": "G", "dmma": "D",
        "dmma-multipass": "M", "dmma-l2": "L", "wspec5": "w", "sym5": "s", "sym4": "h", "wspec": "W", "regtile": "r", "rows2": "R"}
print("n\\d " + "".join(f"{d:>10d}" for d in ds))
for n in ns:
    line = f"{n:3d} "
    for d in ds:
        r = tab.get((n, d))
        line += f"{'-':>10s}" if r is None else f"{r[field]:>8.3f}{code.get(r['path'], '?'):>2s}"
    print(line)
