#!/bin/bash
tag=r02t; out=gpurun_out; mkdir -p $out
timeout -k 5 700 python -m pytest tests -m gpu -x -q --durations=3 2>&1 | tail -12 > $out/${tag}_tests.log; tail -2 $out/${tag}_tests.log
b() { name=$1; shift; timeout -k 5 150 python bench.py --no-cpu-baseline "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; python scripts/show_bench.py $out/${tag}_bench_$name.json; python - $out/${tag}_bench_$name.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); f=j["roofline"]["by_family"]
    print("   families:", {k: ("%.3e" % v["bead_moves_per_s"] if "bead_moves_per_s" in v else "%.1f us" % (1e3*v["launch_ms_marginal"])) for k,v in f.items()})
except Exception as ex: print("   no families", ex)
PY
tail -2 $out/${tag}_bench_$name.err | cut -c1-300; }
b c3 --workload c3 --steps 3
b c4 --workload c4 --steps 3
