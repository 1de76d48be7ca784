"""Print the SASS of one kernel from an `ncu --page source --csv` export with executed-instruction
counts, stall samples and shared-memory wavefronts (run here, no GPU needed).
usage: ncu_sass.py file.csv [instance] [min_share_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 10:
        cur["rows"].append(r)
b = blocks[inst]
h = b["hdr"]
ix = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed",
                              "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal")}
tot = sum(int(r[ix["Instructions Executed"]]) for r in b["rows"])
tots = sum(int(r[ix["# Samples"]]) for r in b["rows"])
print(b["name"][:100], "instances", len(blocks), "warp-inst", tot, "samples", tots)
for n, r in enumerate(b["rows"]):
    ie = int(r[ix["Instructions Executed"]]); te = int(r[ix["Thread Instructions Executed"]])
    print(f"{n:5d} {100*ie/tot:6.2f}% smp {100*int(r[ix['# Samples']])/max(tots,1):5.2f}% thr {te/max(ie,1):5.1f} "
          f"wf {r[ix['L1 Wavefronts Shared']]:>10s}/{r[ix['L1 Wavefronts Shared Ideal']]:>10s} {r[ix['Source']].strip()}")
