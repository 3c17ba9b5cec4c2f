#!/usr/bin/env python
"""Fold the raw evidence CSVs of tools/r02_flops.sh and tools/r02_final_profile.sh (in gpurun_out/) into
profiles/ncu_summary.json (what bench.py reads) and copy them to profiles/ (developer tool)."""
import collections, csv, importlib.util, json, os, re, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)


def metrics(path):
    out = {}
    for r in csv.reader(open(path)):
        if len(r) > 14 and r[0].isdigit():
            out[r[12]] = float(r[14].replace(',', ''))
            out['kernel'] = r[4]
    return out


res = {}
for name, f, evals in [('ensemble_dias', 'r02_flops_dias.csv', 1024 * 128 * 201), ('ensemble_shin', 'r02_flops_shin.csv', 1024 * 128 * 201),
                       ('ensemble_colecole2_n20', 'r02_flops_colecole2_n20.csv', 1024 * 64 * 201)]:
    m = metrics('gpurun_out/' + f)
    fma, mul, add = [m[f'smsp__sass_thread_inst_executed_op_d{k}_pred_on.sum'] for k in ('fma', 'mul', 'add')]
    res[name] = {'kernel': m['kernel'], 'evals': evals, 'dfma': fma, 'dmul': mul, 'dadd': add, 'flop_per_eval': (2 * fma + mul + add) / evals,
                 'fp64_inst_per_eval': (fma + mul + add) / evals,
                 'fp64_pipe_pct_active': m['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'], 'ms': m['gpu__time_duration.sum'] / 1e6}
    print(name, m['kernel'][:70], {k: round(v, 1) for k, v in res[name].items() if isinstance(v, float)})
    shutil.copy('gpurun_out/' + f, 'profiles/' + f)
t = metrics('gpurun_out/r02_traffic_bench_size.csv')
spec = importlib.util.spec_from_file_location('bench', 'bench.py'); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
p = 'profiles/ncu_summary.json'
j = json.load(open(p))
j['fp64_flop_per_eval'] = {'source': 'ncu --metrics smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on.sum on the config kernels (tools/r02_flops.sh; raw CSVs '
                                     "profiles/r02_flops_*.csv): flops = 2 dfma + dmul + dadd, per log-prob evaluation incl. the sampler's own FP64 work",
                           **{k: v['flop_per_eval'] for k, v in res.items()}, 'detail': res}
e = j['ensemble_decomp']
e['dram_bytes_per_launch_at_bench_size'] = t['dram__bytes_read.sum'] + t['dram__bytes_write.sum']
e['dram_bytes_source'] = ('measured at bench size (12,500 spectra, 2000 steps, 100 kept steps) on ' + t['kernel'] + ': ncu --metrics dram__bytes_read.sum,'
                          'dram__bytes_write.sum on that launch (profiles/r02_traffic_bench_size.csv, round 2, tools/r02_final_profile.sh)')
e['kernel_sources_sha'] = b.kernel_sources_sha()
e['kernel_at_bench_size'] = t['kernel']
json.dump(j, open(p, 'w'), indent=1)
print('traffic GB', e['dram_bytes_per_launch_at_bench_size'] / 1e9, 'sha', e['kernel_sources_sha'], 'kernel ms', t['gpu__time_duration.sum'] / 1e6)
for f in ('r02_traffic_bench_size.csv', 'r02_launches.csv'):
    shutil.copy('gpurun_out/' + f, 'profiles/' + f)
rows = [r for r in csv.reader(open('gpurun_out/r02_launches.csv')) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    k = re.sub(r'\(.*', '', r[4]); tot.setdefault(k, [0, 0.0]); tot[k][0] += 1; tot[k][1] += float(r[14].replace(',', '')) / 1e6
for k, (n, ms) in tot.items():
    if 'bisip' in k:
        print(f'{n:4d} {ms:10.2f} ms  {ms / n:9.2f} each  {k[:100]}')
