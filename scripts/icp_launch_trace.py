"""Prints the kernel sequence of the last ICP call in an ncu launch list (C = certify, S = search, A = accumulate; us)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, st = r, i
        break
ki, mi = h.index("Kernel Name"), h.index("Metric Value")
seq = [(r[ki].split('(')[0].replace('void ', '').replace('opb::', ''), float(r[mi].replace(',', '')) / 1e3) for r in rows[st + 1:] if len(r) > mi]
idx = [i for i, (k, _) in enumerate(seq) if k.startswith('icp_bbox')]
s2 = seq[idx[-1]:] if idx else seq
line = []
for k, t in s2:
    tag = {'icp_certify': 'C', 'icp_search_': 'S', 'icp_accumul': 'A'}.get(k[:11])
    line.append(f"{tag}{t:.1f}" if tag else f"{k[:16]} {t:.1f}")
print(' | '.join(line))
