#!/bin/bash
# r02y: A/B on the headline workload -- default (2-stage TMA ring in the Energy pass) vs A (3 stages) vs B (A + permutation entry of the next
# centre-of-mass proposal prefetched) vs C (A + k_measure compiled for 3 CTAs per SM: 80 registers, no spills); a quick Energy parity check per variant
tag=r02y; out=gpurun_out; mkdir -p $out
for v in "" mA mB mC; do
  name=c2${v:+_$v}
  if [ -n "$v" ]; then export PIMC_B200_SO=$PWD/pimc_jl_b200/libpimc_b200_$v.so; else unset PIMC_B200_SO; fi
  timeout -k 5 100 python bench.py --no-cpu-baseline > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err
  python - $out/${tag}_bench_$name.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); f=j["roofline"]["by_family"]
    print(sys.argv[1], "value %.4e e2e %.4e ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), {k: ("%.4e" % v["bead_moves_per_s"] if "bead_moves_per_s" in v else "%.1f us" % (1e3*v["launch_ms_marginal"])) for k,v in f.items()}, j["check"])
except Exception as ex: print(sys.argv[1], "FAILED", ex)
PY
done
for v in mA mC; do
  PIMC_B200_SO=$PWD/pimc_jl_b200/libpimc_b200_$v.so timeout -k 5 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline_shapes or default_dispatch or energy_objects" 2>&1 | tail -2 > $out/${tag}_tests_$v.log; cat $out/${tag}_tests_$v.log
done
