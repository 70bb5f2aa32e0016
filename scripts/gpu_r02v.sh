#!/bin/bash
# r02v: (1) ncu --set full captures of the kernels / instantiations without one so far (verdict, missing #7): k_measure after the round-2
# rewrite, the staging and centre-of-mass halves of k_sweep<0,4> alone, k_sweep<3,8>/k_measure<3,8> (C4 lattice, Density), k_sweep<1,4> +
# Density (C3), k_sweep<1,2> (C5), k_chain<1,2> (C5 default), the optimistic interacting kernels + k_cells_build (C4i);
# (2) A/B of 128-thread k_sweep CTAs (eight per SM) against the default 256 (four per SM) on the headline workload.
tag=r02v; out=gpurun_out; mkdir -p $out
capm() { # capm <name> <kernel regex> <skip> <count> <cmd...>: several launches of one process, summarised on the box
  name=$1; re=$2; skip=$3; cnt=$4; shift 4
  timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:"$re" -s $skip -c $cnt -f -o /tmp/${tag}_$name "$@" > $out/${tag}_ncu_$name.log 2>&1
  python scripts/ncu_summary.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_${name}_ncu_raw_summary.txt 2>&1
  grep -E "Kernel Name|gpu__time_duration.sum" $out/${tag}_${name}_ncu_raw_summary.txt | paste - - | cut -c1-260
  tail -3 $out/${tag}_ncu_$name.log | cut -c1-200
}
timeout -k 5 300 python -m pytest tests -m gpu -x -q -k "structure or checkpoint or paircorr or c_driver" 2>&1 | tail -6 > $out/${tag}_tests_new.log; cat $out/${tag}_tests_new.log
# C2: 260 thermalisation launches, then com alone x2, reshape alone x2, mix x4 with 2 k_measure
capm c2 "k_sweep|k_measure" 260 10 python scripts/probe_cfg.py c2 2 260
python scripts/ncu_lines.py /tmp/${tag}_c2.ncu-rep k_measureILi0ELi4E 30 > $out/${tag}_k_measure_source_lines.txt 2>&1
capm c4 "k_sweep|k_measure" 150 12 python scripts/probe_cfg.py c4 2 150
python scripts/ncu_lines.py /tmp/${tag}_c4.ncu-rep k_sweepILi3ELi8ELb0E 30 > $out/${tag}_k_sweep_c4_source_lines.txt 2>&1
capm c3 "k_sweep|k_measure" 150 16 python scripts/probe_cfg.py c3 2 150
capm c5chain "k_chain" 1 3 python scripts/probe_cfg.py c5 0 150
capm c4i "k_isweep|k_iswap|k_cells_build|k_relink" 0 8 python scripts/probe_cfg.py c4i 0 3 2
# (2) the A/B
for so in "" t128; do
  name=c2${so:+_$so}
  PIMC_B200_SO=${so:+$PWD/pimc_jl_b200/libpimc_b200_$so.so} timeout -k 5 150 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err
  python scripts/show_bench.py $out/${tag}_bench_$name.json
  python - $out/${tag}_bench_$name.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); f=j["roofline"]["by_family"]
    print("   families:", {k: ("%.3e" % v["bead_moves_per_s"] if "bead_moves_per_s" in v else "%.1f us" % (1e3*v["launch_ms_marginal"])) for k,v in f.items()})
except Exception as ex: print("   no families", ex)
PY
done
win=$(python - <<'PY'
import json
v=lambda f: json.loads(open(f).read().strip().splitlines()[-1])["value"]
try: print(1 if v("gpurun_out/r02v_bench_c2_t128.json") > 1.02 * v("gpurun_out/r02v_bench_c2.json") else 0)
except Exception: print(0)
PY
)
echo "t128 wins: $win"
if [ "$win" == "1" ]; then
PIMC_B200_SO=$PWD/pimc_jl_b200/libpimc_b200_t128.so timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline_shapes or trajectory or sweep_edge or default_dispatch" 2>&1 | tail -3 > $out/${tag}_tests_t128.log; cat $out/${tag}_tests_t128.log
fi
