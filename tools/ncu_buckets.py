"""Bucket the warp-stall samples of an ncu source page (SASS view) into runs of instructions: quick where-is-the-time view."""
import csv, subprocess, sys
rep, step = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))[2:]
tot = sum(int(r[2]) for r in rows)
print("total samples", tot, "instructions", len(rows))
for i in range(0, len(rows), step):
    blk = rows[i:i + step]; s = sum(int(r[2]) for r in blk)
    if s * 200 < tot: continue
    ops = {}
    for r in blk:
        w = r[1].split(); op = w[1] if w[0].startswith("@") else w[0]
        ops[op] = ops.get(op, 0) + int(r[2])
    print(i, s, "%.1f%%" % (100 * s / tot), sorted(ops.items(), key=lambda x: -x[1])[:4])
