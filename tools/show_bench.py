#!/usr/bin/env python
"""Print the essentials of bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith('{'):
            continue
        d = json.loads(l)
        print('== %s: N=%d value %.4e frac %.3f ms/step %.2f launches %d' % (path, d['n_gpus'], d['value'], d['roofline']['frac'], d['ms_per_step'], d['gpu_launches']))
        print('   ', d['config']['partition'], '| exchanges', d['config'].get('halo_exchanges_total'), 'timeouts', d['config'].get('halo_wait_timeouts'))
        if 'e2e' in d:
            print('    e2e %.4e (%.2f ms/step) h2d %d d2h %d %s' % (d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['e2e'].get('matches_device_run')))
        print('    clocks', d['clocks'])
        if 'cpu_baseline' in d:
            print('    cpu %.3e %s' % (d['cpu_baseline']['value'], d['cpu_baseline']['sample']))
        for k, v in d.get('also', {}).items():
            print('    also %s %.4e frac %.3f ms %.2f clocks %s MHz %s' % (k, v['value'], v['frac'], v['ms_per_step'], v['clocks'].get('sm_mhz'), v['clocks'].get('reasons')))
