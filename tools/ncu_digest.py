"""Digest of an .ncu-rep: key raw metrics, stall totals, hot SASS lines. Usage: python tools/ncu_digest.py rep [nlines]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_active.avg', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in want or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.03):
        print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot_s = sum(int(r[ix['# Samples']]) for r in data)
tot_i = sum(int(r[ix['Instructions Executed']]) for r in data)
print('samples', tot_s, 'warp-inst', tot_i)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    o = m.group(2).split('.')[0] if m else '?'
    op[o] += int(r[ix['Instructions Executed']]); ops[o] += int(r[ix['# Samples']])
print(' '.join(f'{o}:{c / tot_i * 100:.1f}/{ops[o] / tot_s * 100:.1f}' for o, c in op.most_common(18)))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
for r in data:
    for s in st:
        tot[s] += int(r[ix[s]] or 0)
T = sum(tot.values())
print(' '.join(f'{s[6:]}:{c / T * 100:.1f}' for s, c in tot.most_common(10)))
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:nl]
for i in sorted(top):
    r = data[i]
    reasons = sorted(((int(r[ix[k]] or 0), k[6:]) for k in st if k != 'stall_selected'), reverse=True)[:2]
    print(i, r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:70], reasons)
