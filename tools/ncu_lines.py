"""Development aid: join an `ncu --page source --csv --print-source sass` export with `nvdisasm --print-line-info` of the
same build (instruction order is identical) and print stall samples per source line.

    python tools/ncu_lines.py <ncu csv> <nvdisasm sass> <kernel substring> [top N]
"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
kern = sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = []
inside, cur = False, None
for line in open(sys.argv[2]):
    if line.startswith("//--------------------- .text."):
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", line):
        lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.defaultdict(lambda: collections.Counter())
for ln, r in zip(lines, data):
    a = agg[ln]
    a["samples"] += int(r[isamp] or 0)
    a["exec"] += int(r[iex] or 0)
    a["n"] += 1
    for k in stalls:
        a[k] += int(r[hdr.index(k)] or 0)
tot = sum(a["samples"] for a in agg.values())
print("total samples", tot)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    top = sorted(((k[6:], a[k]) for k in stalls if a[k]), key=lambda kv: -kv[1])[:4]
    print(f"{ln[0] if ln else None}:{ln[1] if ln else 0:5d} {a['samples']:8d} {100*a['samples']/tot:5.1f}%  n={a['n']:4d} exec={a['exec']:10d}  " + " ".join(f"{k}={v}" for k, v in top))
