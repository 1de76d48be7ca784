"""Bucketed instruction/sample shares over the SASS of one kernel (first launch in the report)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] if len(his) > 1 else len(rows)
hdr = rows[hi]
ii, si, ti, wi = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = [r for r in rows[hi + 1:end] if len(r) > ii and r[ii].isdigit()]
tot = sum(int(r[ii]) for r in data); tots = sum(int(r[ti]) for r in data)
print(len(data), 'sass lines; total warp instr', tot)
for k in range(0, len(data), B):
    blk = data[k:k + B]
    n = sum(int(r[ii]) for r in blk); s = sum(int(r[ti]) for r in blk)
    if n > 0.012 * tot or s > 0.015 * tots:
        ops = ' '.join((r[si].split()[1] if r[si].strip().startswith('@') else r[si].split()[0]) for r in blk)
        thr = sum(float(r[wi]) for r in blk) / len(blk)
        print(f"{k:4d} instr {100*n/tot:5.1f}% smp {100*s/tots:5.1f}% thr~{thr:4.1f} | {ops[:170]}")
