"""Per-source-line instruction counts of an ncu report:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_src.csv ; python profiles/src_hot.py X_src.csv [problems]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nprob = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
fn = None; hdr = None; agg = {}; local = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fn = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[2] != "-": continue
    d = dict(zip(hdr, r))
    key = (fn, int(r[0]))
    inst = float(d["Instructions Executed"] or 0)
    e = agg.setdefault(key, [r[1], 0.0, 0.0, 0.0])
    e[1] += inst
    e[2] += float(d.get("L1 Wavefronts Shared") or 0)
    e[3] += float(d.get("L2 Theoretical Sectors Local") or 0)
tot = sum(v[1] for v in agg.values())
print(f"total inst {tot:.4g}  per problem {tot / nprob:.1f}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{v[1] / nprob:9.1f} {100 * v[1] / tot:5.1f}%  smem_wf {v[2] / nprob:8.1f} local {v[3]/nprob:6.1f}  {f}:{ln}: {v[0].strip()[:110]}")
