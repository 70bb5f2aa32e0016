"""one-line digest of bench.py JSON lines: show_bench.py file..."""
import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r = j["roofline"]
        print(f"{f}: value {j['value']:.3e} e2e {j['e2e']['value']:.3e} roofline {r['frac']:.3f} ({r['kernel']}, {r['launch_ms'] * 1e3:.1f} us) "
              f"step {r['step_including_estimator']['frac']:.3f} ms/step {j['ms_per_step']:.1f} cpu {j.get('cpu_baseline', {}).get('value', 0):.3e} check {j['check']}")
    except Exception as ex:
        print(f, "FAILED", ex)
