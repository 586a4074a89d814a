#!/usr/bin/env python
"""Summarise an ncu --metrics CSV of the tcgen05 GEMM launches of one iteration (per kernel / grid averages).
usage: python scripts/summarize_tc_metrics.py gpurun_out/tc_metrics_r1.csv > profiles/r1_tc_metrics.md"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
per = {}
for r in csv.DictReader(lines):
    d = per.setdefault(r['ID'], {'name': re.sub(r'^void tc::', '', r['Kernel Name']).split('(CUtensorMap')[0], 'grid': r['Grid Size'], 'block': r['Block Size']})
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    if r['Metric Name'] == 'gpu__time_duration.sum':
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(u, v)
    v *= {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    d[r['Metric Name']] = v
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault((d['name'], d['grid'], d['block']), collections.defaultdict(float)); a['n'] += 1
    for m, v in d.items():
        if isinstance(v, float): a[m] += v
print('# ncu metrics of the tcgen05 GEMM launches of one REINFORCE iteration (B=64, K=5, T_v=80)\n')
print('`ncu --metrics gpu__time_duration.sum,dram__bytes_*,lts__t_bytes.sum,sm__pipe_tensor_cycles_active... --clock-control none` (caches flushed before every kernel: DRAM reads include the L2-resident weights)\n')
print('| kernel | grid | block | launches | us / launch | DRAM read MB | DRAM write MB | L2 MB | tensor pipe % | warps active % | regs |')
print('|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|')
tot_step_bytes = tot_step_n = 0
for k, a in sorted(agg.items(), key=lambda x: -x[1]['gpu__time_duration.sum']):
    n = a['n']
    print('| `%s` | %s | %s | %d | %.1f | %.2f | %.2f | %.1f | %.1f | %.1f | %d |' % (k[0], k[1], k[2], n, a['gpu__time_duration.sum'] / n, a['dram__bytes_read.sum'] / n / 1e6, a['dram__bytes_write.sum'] / n / 1e6, a['lts__t_bytes.sum'] / n / 1e6, a['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'] / n, a['sm__warps_active.avg.pct_of_peak_sustained_active'] / n, a['launch__registers_per_thread'] / n))
    if 'EpiLstm' in k[0]:
        tot_step_bytes += a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']; tot_step_n += n
if tot_step_n:
    print('\nrecurrent-step kernels: mean DRAM traffic per launch = %.0f bytes over %d launches' % (tot_step_bytes / tot_step_n, tot_step_n))
