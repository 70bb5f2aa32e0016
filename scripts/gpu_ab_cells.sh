for c in c3i c4i; do
  PIMC_PROF=1 timeout 300 python bench.py --workload $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01h_bench_${c}.json 2> gpurun_out/r01h_bench_${c}.err
  python scripts/show_bench.py gpurun_out/r01h_bench_${c}.json; grep "pimc prof" gpurun_out/r01h_bench_${c}.err | tail -2
done
