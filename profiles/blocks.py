"""Group the SASS of an ncu source page (ncu -i X.ncu-rep --page source --csv) into runs of equal execution count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 15
hdr, data = rows[1], rows[2:]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
out, cur, start, n, samp = [], None, 0, 0, 0
for i, r in enumerate(data):
    e = int(r[iE])
    if cur is None or abs(e - cur) > 0.02 * max(cur, 1):
        if cur is not None:
            out.append((start, i - 1, cur / steps, n, samp))
        cur, start, n, samp = e, i, 0, 0
    n += 1
    samp += int(r[iSm])
out.append((start, len(data) - 1, cur / steps, n, samp))
tot = sum(int(r[iSm]) for r in data)
for s, e, c, n, sm in out:
    if c * n > thr:
        print("%5d-%5d exec/flip %7.2f x %3d instrs = %7.1f   samples %4.1f%%   %s" % (s, e, c, n, c * n, 100 * sm / tot, data[s][iS].strip()[:50]))
