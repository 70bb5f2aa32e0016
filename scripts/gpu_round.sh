#!/bin/bash
# one gpurun call: GPU tests, bench lines of every SURVEY 8d configuration, launch list + ncu --set full captures
# usage: scripts/gpu_round.sh <tag> [what...]   what in: tests bench configs ncu ncu2   (gpurun copies back at most 64 MiB: one or two .ncu-rep per call)
tag=$1; shift
what=${*:-tests bench configs ncu}
out=gpurun_out; mkdir -p $out
for w in $what; do case $w in
tests)   timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_tests.log; tail -3 $out/${tag}_tests.log ;;
bench)   timeout 600 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; tail -c 600 $out/${tag}_bench_c2.json ;;
configs) for c in c2s c2f c1 c5 c3 c4 c3i c4i; do
           timeout 300 python bench.py --workload $c --steps 3 --warmup 3 > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err
           python - $out/${tag}_bench_$c.json $c <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "value %.3e e2e %.3e frac %.3f cpu %.3e"%(j["value"], j["e2e"]["value"], j["roofline"]["frac"], j.get("cpu_baseline",{}).get("value",0)), j["check"])
except Exception as ex: print(sys.argv[2], "FAILED", ex)
PY
           tail -3 $out/${tag}_bench_$c.err
         done ;;
ncu)     timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --iters 40 --therm 60 --no-cpu-baseline > $out/${tag}_ncu_b.log 2>&1
         timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 200 -c 1 -f -o $out/${tag}_k_sweep python scripts/probe_ncu.py 4096 260 > $out/${tag}_ncu_sweep.log 2>&1
         ls -la $out | grep $tag ;;
ncu2)    timeout 900 ncu --set full --clock-control none -k regex:k_measure -s 60 -c 1 -f -o $out/${tag}_k_measure python scripts/probe_ncu.py 4096 260 > $out/${tag}_ncu_meas.log 2>&1
         timeout 900 ncu --set full --clock-control none -k regex:k_swap_iter -s 100 -c 1 -f -o $out/${tag}_k_swap python bench.py --workload c2s --steps 1 --warmup 3 --iters 60 --therm 60 --no-cpu-baseline > $out/${tag}_ncu_swap.log 2>&1
         ls -la $out | grep $tag ;;
esac; done
