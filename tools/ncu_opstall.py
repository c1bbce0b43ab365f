"""opcode x stall-reason sample matrix of an ncu source page: python tools/ncu_opstall.py <rep>"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
st_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
M = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for r in data:
    op = r[1].split()
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    for c in st_cols:
        v = int(r[ix[c]] or 0)
        M[o][c[6:]] += v
        tot[c[6:]] += v
T = sum(tot.values())
print("all:", " ".join(f"{k}:{100*v/T:.1f}%" for k, v in tot.most_common(10)))
for o, cnt in sorted(M.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
    s = sum(cnt.values())
    print(f"{o:8s} {100*s/T:5.1f}%  " + " ".join(f"{k}:{100*v/T:.1f}" for k, v in cnt.most_common(5) if v))
