"""Side-by-side view of two fullbench JSONL files: python tools/fbshow.py new.jsonl [old.jsonl]"""
import json, sys
def load(f):
    return {(r['n'], r['d']): r for r in map(json.loads, open(f))}
new = load(sys.argv[1]); old = load(sys.argv[2]) if len(sys.argv) > 2 else {}
for k, r in new.items():
    o = old.get(k)
    print(f"n={k[0]:2d} d={k[1]} nb={r['nb']:9d} {r['path']:18s} {r['ms']:9.4f} ms {r['gflops']:9.1f} GF {r['alg_gbs']:7.1f} GB/s frac {r['roofline_frac']:.3f}"
          + (f"   was {o['ms']:9.4f} ms ({o['path']}) x{o['ms']/r['ms']:.2f}" if o else ""))
