"""Developer tool: opcode mix (warp instructions per warp-step) and the headline counters of one ncu report.
usage: python tools/ncu_mix.py raw.csv source.csv <warp-steps>"""
import csv, collections, sys
raw, src, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
want = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', '_per_issue_active.ratio']
for h, v in zip(hdr, vals):
    if any(w in h for w in want) and not v.startswith('0.0'):
        print(h, '=', v)
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[ie]) for r in data)
mix, smp = collections.Counter(), collections.Counter()
for r in data:
    t = r[ia].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    mix[op] += int(r[ie]); smp[op] += int(r[isamp])
print('total warp instr', tot, 'per warp-step', round(tot / per, 1))
for op, c in mix.most_common(28):
    print(f"{op:10s} {c / per:7.1f} {100 * c / tot:5.1f}%  samples {smp[op]}")
