#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k sharded > $O/r02_pytest_sharded.log 2>&1; echo "pytest sharded rc=$?" | tee -a $O/r02_rc3.log
tail -5 $O/r02_pytest_sharded.log
B="python bench.py --workload cfg5 --trees 6000 --steps 1 --warmup 1 --no-e2e --no-cpu --quick-check"
if timeout 200 $B > $O/r02_tol_plain.json 2>&1; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tolerant_kernel -c 1 -f -o $O/r02_tol_full $B > $O/r02_ncu_tol.log 2>&1
  echo "ncu tol rc=$?" | tee -a $O/r02_rc3.log
fi
head -c 400 $O/r02_tol_plain.json
