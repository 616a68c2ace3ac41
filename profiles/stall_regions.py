"""Aggregate the warp-stall samples of an `ncu --set full --import-source on` report by barrier region
(SASS between two BAR.SYNC) and list the hottest instructions.
usage: python profiles/stall_regions.py report.ncu-rep   (needs ncu on PATH; reads no GPU)"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
h = rows[1]
si, ni, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for idx, r in enumerate(rows[2:]):
    try:
        data.append((idx, int(r[ni]), int(r[ie]), r[si].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data)
print(f"{tot} stall samples over {len(data)} SASS instructions")
cur, start = 0, 0
for idx, n, e, s in data:
    cur += n
    if "BAR.SYNC" in s or "EXIT" in s:
        if cur > 0.02 * tot:
            print(f"  SASS [{start:5d},{idx:5d}] up to {s.split()[0]:9s} {cur:6d} samples {100 * cur / tot:5.1f} %")
        cur, start = 0, idx + 1
print("hottest instructions (index, samples, executed, SASS):")
for idx, n, e, s in sorted(data, key=lambda d: -d[1])[:12]:
    print(f"  {idx:5d} {n:5d} {e:8d}  {s}")
