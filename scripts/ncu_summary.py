"""Summarises an ncu report: key raw metrics, dynamic instruction mix and top stall sites.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [n_top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
print('kernel:', d.get('Kernel Name', ('', ''))[1])
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for k in keys:
    if k in d:
        print(f'{k:75s} {d[k][0]:16s} {d[k][1]}')
for k in d:
    if 'issue_stalled' in k and 'per_issue_active' in k:
        try:
            v = float(d[k][1])
        except ValueError:
            continue
        if v > 0.15:
            print(f"stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):22s} {v:.3f}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sn = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
byop, byop_s = collections.Counter(), collections.Counter()
tot_i = tot_s = 0
top = []
for r in rows[2:]:
    try:
        n, s = int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
    except (ValueError, IndexError):
        continue
    op = r[ix['Source']].split()
    o = (op[1] if op and op[0].startswith('@') else (op[0] if op else '')).split('.')[0]
    byop[o] += n; byop_s[o] += s; tot_i += n; tot_s += s
    top.append((s, n, r[ix['Source']][:60], {h[6:]: int(r[ix[h]]) for h in sn if r[ix[h]] not in ('', '0')}))
print(f'\ndynamic instruction mix ({tot_i} warp instructions, {tot_s} stall samples)')
for o, n in byop.most_common(16):
    print(f'  {o:10s} inst {100 * n / tot_i:5.1f}%   samples {100 * byop_s[o] / max(tot_s, 1):5.1f}%')
top.sort(key=lambda x: -x[0])
print('\ntop stall-sample instructions (samples, executed, SASS, reasons)')
for t in top[:ntop]:
    print('  ', t[0], t[1], t[2], dict(sorted(t[3].items(), key=lambda kv: -kv[1])[:3]))
