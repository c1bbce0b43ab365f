"""Shared-memory wavefronts per opcode from the source page of an .ncu-rep:  python tools/ncu_smem_ops.py <rep> [...]
(executed warp instructions, wavefronts, ideal wavefronts -- shows which accesses load the shared-memory pipe)."""
import collections, csv, io, subprocess, sys
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rep, rows[0][1][:110] if rows and len(rows[0]) > 1 else "")
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0, 0, 0]); tot = 0
    for r in rows[2:]:
        if len(r) != len(hdr): continue
        ins = r[1].strip().split()
        op = ins[1] if ins[0].startswith('@') else ins[0]
        w = int(r[ix["L1 Wavefronts Shared"]] or 0); wi = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0); ex = int(r[ix["Instructions Executed"]] or 0)
        tot += w
        if w:
            a = agg[op]; a[0] += ex; a[1] += w; a[2] += wi; a[3] += 1
    for op, (ex, w, wi, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {op:26s} static {cnt:4d} executed {ex:11d} wavefronts {w:11d} ({100*w/tot:4.1f} %) ideal {wi:11d}  per-exec {w/ex:5.2f} ideal {wi/ex:5.2f}")
