"""Summarise an ncu raw.csv + src.csv pair (ncu -i X.ncu-rep --page raw|source --csv) for the SA kernel."""
import csv, sys
from collections import Counter
raw, src, steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum ',
        'smsp__average_warps_issue_stalled', 'lts__throughput.avg.pct', 'launch__registers_per_thread ',
        'launch__occupancy_limit', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ', 'launch__grid_size', 'smsp__warps_eligible.avg.per_cycle_active']
for h, u, v in zip(hdr, units, vals):
    if any(w in h + ' ' for w in want):
        try:
            if float(v) == 0:
                continue
        except ValueError:
            pass
        print(h, u, v)
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data)
print('instructions per flip-evaluation: %.1f' % (tot / steps))
c, cs = Counter(), Counter()
for r in data:
    toks = r[iS].split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    c[op] += int(r[iE]); cs[op] += int(r[iSm])
for op, n in c.most_common(22):
    print('%-12s %8.1f per flip   samples %d' % (op, n / steps, cs[op]))
print('--- most sampled instructions')
for r in sorted(data, key=lambda r: -int(r[iSm]))[:18]:
    print(r[iSm], '%.2f' % (int(r[iE]) / steps), r[iS].strip()[:90])
