#!/bin/bash
# r02z: final check of the round on HEAD: whole GPU suite, smoke(), the headline bench line with the CPU legs, ncu capture of k_structure
tag=r02z; out=gpurun_out; mkdir -p $out
timeout -k 5 400 python -m pytest tests -m gpu -x -q --durations=3 2>&1 | tail -12 > $out/${tag}_tests.log; tail -2 $out/${tag}_tests.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout -k 5 200 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; python scripts/show_bench.py $out/${tag}_bench_c2.json
timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:k_structure -s 1 -c 1 -f -o /tmp/${tag}_k_structure python scripts/probe_estim.py 1024 > $out/${tag}_ncu_k_structure.log 2>&1
python scripts/ncu_summary.py /tmp/${tag}_k_structure.ncu-rep > $out/${tag}_k_structure_ncu_raw_summary.txt 2>&1
python scripts/ncu_lines.py /tmp/${tag}_k_structure.ncu-rep k_structure 25 > $out/${tag}_k_structure_source_lines.txt 2>&1
python scripts/ncu_digest.py $out/${tag}_k_structure_ncu_raw_summary.txt; tail -2 $out/${tag}_ncu_k_structure.log | cut -c1-300
