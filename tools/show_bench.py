"""Print the per-kernel table of a bench.py JSON line (development helper)."""
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d.get(k) for k in ['value', 'ms_per_step', 'gpu_launches', 'loss', 'kernel_time_accounted']})
        print('e2e', d.get('e2e') and {k: d['e2e'][k] for k in ('value', 'ms_per_step')}, 'clocks', d.get('clocks'))
        steps = d['steps']
        for k in d.get('kernels', []):
            print('%-20s ms/step=%7.3f n=%3d frac=%.3f ach=%9.1f %s share=%.3f' % (k['name'], k['ms'] / steps, k['launches'], k['frac'], k['achieved'], k['unit'], k['share_of_step']))
        print('variants', d.get('variants'))
        print('cpu_baseline', d.get('cpu_baseline') and d['cpu_baseline'].get('value'))
