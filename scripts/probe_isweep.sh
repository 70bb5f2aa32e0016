export PIMC_PROF=1
for args in "c3i 1024 10 2 reshape" "c3i 1024 10 2" "c3i 1024 10 0" "c3i 1024 10 1" "c4i 1024 10 2" "c4i 1024 10 0"; do
  echo "=== $args"; timeout 150 python scripts/probe_isweep.py $args 2>&1 | tail -8
done
