"""Aggregate the SASS source page of one kernel by executed instructions (top regions)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ii, si, ti, wi = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = [r for r in rows[hi + 1:] if len(r) > ii and r[ii].isdigit()]
tot = sum(int(r[ii]) for r in data)
tots = sum(int(r[ti]) for r in data)
print("total warp instr", tot, "samples", tots)
# print contiguous listing with share, only lines >= 0.4% of instr or samples
for k, r in enumerate(data):
    n, s = int(r[ii]), int(r[ti])
    if n >= 0.004 * tot or s >= 0.006 * tots:
        print(f"{k:5d} {100*n/tot:5.1f}% instr {100*s/tots:5.1f}% smp thr={r[wi]:>5s} {r[si].strip()[:90]}")
