"""Per-launch table from an `ncu --metrics ... --csv` log (tools/gpu_iter.sh): time, DMMA pipe, issue, DRAM / L2 bytes.
usage: python tools/ncu_counters.py <counters.csv> [min_us]"""
import csv, re, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    k = int(r[idx['ID']])
    per.setdefault(k, {'name': re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '')})
    v = float(r[idx['Metric Value']].replace(',', '')); u = r[idx['Metric Unit']]
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}.get(u, 1.0)
    per[k][r[idx['Metric Name']]] = v * scale
tot = sum(v['gpu__time_duration.sum'] for v in per.values())
print("%d launches, %.1f us serialised" % (len(per), tot))
agg = collections.defaultdict(lambda: [0.0, 0.0])
for k, v in per.items():
    t = v['gpu__time_duration.sum']; dm = v.get('sm__inst_executed_pipe_tensor_subpipe_dmma.sum', 0)
    agg[v['name']][0] += t; agg[v['name']][1] += dm * 512
    if t >= min_us:
        print("%3d %-36s %9.1f us %5.1f%%  dmma %5.1f%% issue %5.1f%%  issued %5.1f TF/s  dram rd %6.2f wr %6.2f GB (%4.2f TB/s)  l2 %6.2f GB (%4.2f TB/s)" % (
            k, v['name'][:36], t, 100 * t / tot, v['sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'],
            v['smsp__issue_active.avg.pct_of_peak_sustained_active'], dm * 512 / t / 1e6, v['dram__bytes_read.sum'], v['dram__bytes_write.sum'],
            (v['dram__bytes_read.sum'] + v['dram__bytes_write.sum']) / t * 1e3, v['lts__t_bytes.sum'], v['lts__t_bytes.sum'] / t * 1e3))
print("per kernel:")
for n, (t, f) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("  %-40s %10.1f us %5.1f %%   issued %5.1f TF/s" % (n, t, 100 * t / tot, f / t / 1e6 if t else 0))
