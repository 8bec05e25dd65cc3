"""Executed instructions per CUDA source line: joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with
`nvdisasm -g <cubin>` of the same build (line info from -lineinfo).  The two listings are matched by instruction order
inside the kernel (ncu prints absolute addresses, nvdisasm offsets).
  python profiles/lines.py src.csv disasm.txt '<mangled kernel substring>' <flip-evaluations> [top]"""
import csv, re, sys
from collections import defaultdict
src, dis, kern, steps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text') or l.startswith('.text.'):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        seq.append(cur)
assert len(seq) == len(data), (len(seq), len(data))
per, samp = defaultdict(int), defaultdict(int)
for loc, r in zip(seq, data):
    per[loc] += int(r[iE]); samp[loc] += int(r[iSm])
tot, tots = sum(per.values()), sum(samp.values())
print("total %.1f instructions per flip-evaluation" % (tot / steps))
srcs = {}
for loc, n in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    f, ln = loc if loc else ("?", 0)
    if f not in srcs:
        try:
            srcs[f] = open([p for p in sys.argv[6:] if p.endswith(f)][0]).read().split('\n')
        except Exception:
            srcs[f] = None
    text = srcs[f][ln - 1].strip()[:80] if srcs[f] and 0 < ln <= len(srcs[f]) else ""
    print("%-22s %4d  %7.1f /flip  %5.1f%% samples   %s" % (f, ln, n / steps, 100.0 * samp[loc] / max(tots, 1), text))
