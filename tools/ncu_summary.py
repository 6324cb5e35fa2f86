"""Summarise an exported ncu report: python tools/ncu_summary.py raw.csv [src.csv]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:80])
    for w in want:
        if w in hdr:
            print('  %-80s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = sorted([(float(r[i]), n) for i, n in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in n and n.endswith('per_issue_active.ratio')], reverse=True)[:7]
    for v, n in st:
        print('  stall %-40s %.2f' % (n.split('stalled_')[1].split('_per_issue')[0], v))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    h = rows[hi[0]]
    body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
    si = h.index('Warp Stall Sampling (All Samples)'); ie = h.index('Instructions Executed')
    tot = sum(int(r[si]) for r in body if len(r) > si and r[si].isdigit())
    print('samples', tot, 'sass lines', len(body))
    top = sorted([(int(r[si]), i, r[1].strip(), r[ie]) for i, r in enumerate(body) if len(r) > si and r[si].isdigit()], reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
    for s, i, src, ex in top:
        print('%5d %4.1f%% line %4d exec %s  %s' % (s, 100 * s / tot, i, ex, src[:90]))
