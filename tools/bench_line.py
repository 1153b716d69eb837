"""One-line digest of a bench.py JSON line on stdin: python tools/bench_line.py <label>"""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
v = d.get('verify') or {}
keys = ('on_facet', 'on_ccd', 'mean_order', 'ccd_hit_fraction', 'on_detector', 'mean_probability_detected')
checks = {k: (float('%.5g' % x) if isinstance(x, float) else x) for k, x in (d.get('checks') or {}).items() if k in keys}
print('%s kernel_ms %.4f frac %.4f verify %s pol %s  %s %s' % (
    sys.argv[1] if len(sys.argv) > 1 else '', r['kernel_ms'], r['frac'], v.get('indices_bit_exact'),
    (v.get('max_rel_err') or {}).get('polarization'), d['kernel_path'][:70], checks))
