export PIMC_PROF=1
for args in "c3i 8 5 1" "c3i 8 5 0" "c3i 64 20 1" "c3i 1024 10 1 reshape" "c3i 1024 10 1 pcom" "c3i 1024 10 1 swap" "c3i 1024 10 1" "c3i 1024 10 0" "c4i 1024 10 1" "c4i 1024 10 0"; do
  echo "=== $args"; timeout 100 python scripts/probe_isweep.py $args 2>&1 | tail -8
done
