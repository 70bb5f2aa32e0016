#!/bin/bash
# r02w: final check of the round on HEAD -- the whole GPU suite, smoke(), the headline bench line, and the A/B of 128-thread k_sweep CTAs
# (eight per SM) against the default 256 (four per SM); the variant is parity-tested only when it wins.
tag=r02w; out=gpurun_out; mkdir -p $out
timeout -k 5 400 python -m pytest tests -m gpu -x -q --durations=3 2>&1 | tail -12 > $out/${tag}_tests.log; tail -2 $out/${tag}_tests.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
fam() { python - $1 <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); f=j["roofline"]["by_family"]
    print("   families:", {k: ("%.3e" % v["bead_moves_per_s"] if "bead_moves_per_s" in v else "%.1f us" % (1e3*v["launch_ms_marginal"])) for k,v in f.items()})
except Exception as ex: print("   no families", ex)
PY
}
for so in "" t128; do
  name=c2${so:+_$so}
  PIMC_B200_SO=${so:+$PWD/pimc_jl_b200/libpimc_b200_$so.so} timeout -k 5 120 python bench.py --no-cpu-baseline > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err
  python scripts/show_bench.py $out/${tag}_bench_$name.json; fam $out/${tag}_bench_$name.json
done
win=$(python - <<'PY'
import json
v=lambda f: json.loads(open(f).read().strip().splitlines()[-1])["value"]
try: print(1 if v("gpurun_out/r02w_bench_c2_t128.json") > 1.015 * v("gpurun_out/r02w_bench_c2.json") else 0)
except Exception: print(0)
PY
)
echo "t128 wins: $win"
if [ "$win" == "1" ]; then
PIMC_B200_SO=$PWD/pimc_jl_b200/libpimc_b200_t128.so timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline_shapes or trajectory or sweep_edge or default_dispatch" 2>&1 | tail -3 > $out/${tag}_tests_t128.log; cat $out/${tag}_tests_t128.log
fi
