"""Group the SASS of one kernel (ncu --page source --csv export) into regions of equal execution
count and print each region's share of warp instructions and stall samples.
usage: ncu_regions.py file.csv [instance] [min_total_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 10:
        cur["rows"].append(r)
b = blocks[inst]; h = b["hdr"]
iS, iN, iE, iT = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
tot = sum(int(r[iE]) for r in b["rows"]); tots = sum(int(r[iN]) for r in b["rows"])
print(b["name"][:110], "| instances", len(blocks), "| warp-inst", tot, "| samples", tots)
groups = []
for n, r in enumerate(b["rows"]):
    ie, te, sm = int(r[iE]), int(r[iT]), int(r[iN])
    if groups and groups[-1]["ie"] == ie:
        g = groups[-1]; g["end"] = n; g["sum"] += ie; g["smp"] += sm; g["te"] += te; g["ops"].append(r[iS].split()[0] if r[iS].split() else "")
    else:
        groups.append({"start": n, "end": n, "ie": ie, "sum": ie, "smp": sm, "te": te, "ops": [r[iS].strip().split()[0]]})
for g in groups:
    if 100 * g["sum"] / tot >= thr:
        ops = {}
        for o in g["ops"]:
            o = o.lstrip("@!P0123456789 ").split(".")[0] if o.startswith("@") else o.split(".")[0]
            ops[o] = ops.get(o, 0) + 1
        top = " ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"{g['start']:5d}-{g['end']:5d} n={g['end']-g['start']+1:4d} inst {100*g['sum']/tot:6.2f}% smp {100*g['smp']/max(tots,1):6.2f}% "
              f"thr {g['te']/max(g['sum'],1):5.1f} | {top}")
