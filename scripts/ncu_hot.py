"""Top stall instructions of a kernel from an ncu report's source page (SASS view)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, iall, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > iall and r[iall].isdigit()]
tot = sum(int(r[iall]) for r in body)
print("total samples", tot, "instructions", len(body))
# classes
from collections import Counter
c = Counter()
for r in body:
    op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
    c[op.split(".")[0]] += int(r[iall])
print("by opcode:", [(k, round(100 * v / tot, 1)) for k, v in c.most_common(12)])
idx = sorted(range(len(body)), key=lambda i: -int(body[i][iall]))[:topn]
for i in sorted(idx):
    r = body[i]
    print(f"{i:5d} {100*int(r[iall])/tot:5.1f}%  ex={r[iex]:>9}  {r[isrc].strip()[:90]}")
