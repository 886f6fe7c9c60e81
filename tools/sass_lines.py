"""Development aid: SASS instructions per source line of one kernel (nvdisasm --print-line-info of an object's cubin).

    python tools/sass_lines.py <sass file> <kernel substring> [file substring] [lo] [hi]
"""
import re, sys, collections
sass, kern = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
inside = False
cur = None
cnt = collections.Counter()
ops = collections.defaultdict(collections.Counter)
for line in open(sass):
    if line.startswith("//--------------------- .text."):
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        cnt[cur] += 1
        ops[cur][m.group(2).split(".")[0]] += 1
tot = 0
for (f, l), n in sorted(cnt.items()):
    if fsub in f and lo <= l <= hi:
        tot += n
        print(f"{f}:{l:5d} {n:5d}  " + " ".join(f"{k}x{v}" for k, v in ops[(f, l)].most_common(6)))
print("total", tot, "of", sum(cnt.values()))
