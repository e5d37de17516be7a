"""Print the headline numbers of bench.py JSON lines (one per line in the given files)."""
import json, sys
for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith('{'):
            continue
        d = json.loads(l)
        r = d['roofline']; c = d['config']
        da = r.get('distance_plus_argmin') or {}
        s = r.get('sustained_leg') or {}
        print(f"{c['name']:10s} {d['scaling']:6s} N={d['n_gpus']} tok/GPU={c['tokens_per_gpu']:7d} value={d['value']/1e6:8.1f} M/s  ms/step={d['ms_per_step']*1e3:7.1f} us  "
              f"search={r['avg_launch_ms']*1e3:6.1f} us frac_burst={r['frac_vs_burst_peak']:.3f} d+a={da.get('frac_vs_burst_peak', 0):.3f}  "
              f"sustained={s.get('value', 0)/1e6:8.1f} M/s  e2e={(d['e2e'] or {}).get('value', 0)/1e6:7.1f} M/s  "
              f"module={(d.get('module_path') or {}).get('ms_per_step', 0)*1e3:6.1f} us  parity={ {k: v for k, v in d['parity'].items() if isinstance(v, bool)} }")
