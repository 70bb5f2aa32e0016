#!/bin/bash
# r02x: the headline bench line on HEAD (default build, with the CPU legs) and C5 (Energy pass with KM = 2 inside the chain-major kernel)
tag=r02x; out=gpurun_out; mkdir -p $out
timeout -k 5 200 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; python scripts/show_bench.py $out/${tag}_bench_c2.json
timeout -k 5 120 python bench.py --workload c5 --steps 3 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; python scripts/show_bench.py $out/${tag}_bench_c5.json
python - <<'PY'
import json
for w in ("c2","c5"):
    try:
        j=json.loads(open(f"gpurun_out/r02x_bench_{w}.json").read().strip().splitlines()[-1]); f=j["roofline"]["by_family"]
        print(w, "families:", {k: ("%.3e" % v["bead_moves_per_s"] if "bead_moves_per_s" in v else "%.1f us" % (1e3*v["launch_ms_marginal"])) for k,v in f.items()})
    except Exception as ex: print(w, "no families", ex)
PY
