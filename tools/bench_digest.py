#!/usr/bin/env python
"""Print the key numbers of a bench.py JSON line (developer tool)."""
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
if j.get('impl') == 'reference':
    print('reference: %.4e evals/s on %d cores (%s)' % (j['value'], j['cpu_baseline']['cores'], j['cpu_baseline']['sample'])); sys.exit()
r = j['roofline']
print('N=%d value %.4e  e2e %.4e  ms/step %.1f  spectra/s %.0f  launches %d' % (j['n_gpus'], j['value'], j['e2e']['value'], j['ms_per_step'], j['spectra_per_s'], j['gpu_launches']))
print('roofline: %.2f TF / %.2f = %.3f  kernel_ms %.1f share %.4f  stats %.2f ms (%.0f GB/s)  traffic %s' % (r['achieved'], r['peak'], r['frac'], r['kernel_ms'], r['kernel_share_of_step'], r['stats_kernel_ms'], r['stats_kernel_GBps'], r['traffic']))
print('clocks', j['clocks'], 'acc %.3f' % j['acceptance_fraction'], 'nan', j['nan_flags'])
if j.get('cpu_baseline', {}).get('value'): print('cpu_baseline %.3e on %d cores' % (j['cpu_baseline']['value'], j['cpu_baseline']['cores']))
s = j.get('strong_scaling')
if s: print('strong: %d spectra on %d GPUs in %.2f s = %.3e evals/s' % (s['spectra'], s['n_gpus'], s['seconds'], s['evals_per_s']))
for k, v in j['variants'].items():
    if isinstance(v, dict): print('variant %-15s kernel %.3e  e2e %s  %s' % (k, v['evals_per_s'], ('%.3e (%d GPUs, %.3f s)' % (v['e2e_evals_per_s'], v['e2e_gpus'], v['e2e_ms_per_step'] / 1e3)) if 'e2e_evals_per_s' in v else '-', v.get('spectra_with_percentiles_identical_to_fp64', '')))
for k, v in (j.get('configs') or {}).items():
    if not isinstance(v, dict): print(k, v); continue
    if k.startswith('C4'):
        for p, e in v.items():
            if isinstance(e, dict): print('%s %-15s kernel %.3e e2e %.3e  %s %s' % (k, p, e['evals_per_s'], e['e2e_evals_per_s'], ('frac %.3f' % e['roofline']['frac']) if e.get('roofline') else 'fwd err %.1e shift %.2f sd x%.1f' % (e['forward_rel_error_vs_fp64'], e['max_median_shift_in_posterior_sd'], e['speedup_vs_fp64']), e['kernel']))
    else:
        print('%-14s kernel %.3e e2e %.3e  pipe frac %.3f (flops %.3f)' % (k, v['evals_per_s'], v['e2e_evals_per_s'], v['roofline']['frac'], v['roofline']['frac_in_flops']))
