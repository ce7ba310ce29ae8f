"""Summarise an ncu report exported as CSV (raw page + source page) — development tool."""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for i, h in enumerate(hdr):
    if h in want:
        print('%-70s %-12s %s' % (h, units[i], vals[i]))
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
tots = sum(agg.values())
print('stalls:', [(k, '%.1f%%' % (100 * v / tots)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
seg, acc = 0, {}
for i, r in enumerate(data):
    n = int(r[ix['Instructions Executed']]); sm = int(r[ix['# Samples']])
    a = acc.setdefault(seg, [0, 0, i, i]); a[0] += n; a[1] += sm; a[3] = i
    if 'BAR.SYNC' in r[ix['Source']]: seg += 1
tot = sum(a[0] for a in acc.values()); ts = sum(a[1] for a in acc.values())
print('total warp instr', tot)
for k, a in acc.items():
    if a[0] / tot > 0.005 or a[1] / ts > 0.01:
        print('seg %2d lines %4d-%4d instr %5.1f%% samples %5.1f%%' % (k, a[2], a[3], 100 * a[0] / tot, 100 * a[1] / ts))
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]
for i in sorted(top):
    r = data[i]
    st = sorted([(int(r[ix[s]]), s) for s in stalls], reverse=True)[:2]
    print(i, r[ix['Source']].strip()[:58].ljust(58), r[ix['# Samples']], r[ix['Instructions Executed']], st)
