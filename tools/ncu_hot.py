"""Per-instruction view of an ncu source page: python tools/ncu_hot.py <rep> [start end]  (rows of the SASS listing)
Prints idx, samples, executed, dominant stall reasons, instruction.  Also a per-region summary split at BAR/WARPSYNC/SYNCS."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
st_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(data))
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
for k in range(lo, min(hi, len(data))):
    r = data[k]
    s = int(r[ix["# Samples"]] or 0)
    sts = sorted(((int(r[ix[c]] or 0), c[6:]) for c in st_cols), reverse=True)[:3]
    print(f"{k:5d} {s:6d} {r[ix['Instructions Executed']]:>9s}  {' '.join(f'{n}:{v}' for v, n in sts if v):40s} {r[1].strip()[:90]}")
