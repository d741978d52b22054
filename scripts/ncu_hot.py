"""Hot spots of one kernel from an `ncu --page source --csv` dump: stall samples grouped into phases between barriers and
by the instruction class that the stalled warps were waiting at.   usage: python scripts/ncu_hot.py file.csv 'XlCztAxis<4096, 0, 1, 1>' [nth]"""
import csv, sys, collections
path, want = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
# split into kernels
ker = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1].replace("(int)", ""), "rows": []}; ker.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sel = [k for k in ker if want in k["name"]]
k = sel[nth]
hdr = k["rows"][0]; body = k["rows"][1:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in body)
print(k["name"], "instructions", len(body), "samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# phases split at BAR / SYNCS
phase = 0; acc = collections.OrderedDict()
for r in body:
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    key = phase
    a = acc.setdefault(key, {"n": 0, "samples": 0, "ops": collections.Counter(), "stalls": collections.Counter(), "first": src})
    s = int(r[ix["# Samples"]])
    a["n"] += 1; a["samples"] += s
    a["ops"][op.split(".")[0]] += s
    for c in stall_cols:
        a["stalls"][c] += int(r[ix[c]])
    if op.startswith("BAR") or op.startswith("SYNCS.PHASECHK"):
        phase += 1
for key, a in acc.items():
    if a["samples"] < 0.01 * tot: continue
    top_ops = ", ".join(f"{o}:{100*v/tot:.1f}" for o, v in a["ops"].most_common(6))
    top_st = ", ".join(f"{o.replace('stall_','')}:{100*v/tot:.1f}" for o, v in a["stalls"].most_common(5))
    print(f"phase {key:2d}: {a['n']:5d} instr  {100*a['samples']/tot:5.1f}% of samples | at: {top_ops} | why: {top_st}")
