#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_agglo.py -x -q > $O/r02_pytest_agglo3.log 2>&1; echo "pytest agglo rc=$?" | tee -a $O/r02_rc5.log
tail -3 $O/r02_pytest_agglo3.log
timeout 400 python scripts/consumers_probe.py 20000 200 8 4 --dendrogram > $O/r02_consumers_probe3.json 2> $O/r02_consumers_probe3.err; echo "probe rc=$?" | tee -a $O/r02_rc5.log
cat $O/r02_consumers_probe3.json; tail -3 $O/r02_consumers_probe3.err
B="python bench.py --workload cfg5 --trees 6000 --steps 1 --warmup 1 --no-e2e --no-cpu --quick-check"
if timeout 200 $B > $O/r02_tol_plain2.json 2>&1; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tolerant_kernel -c 1 -f -o $O/r02_tol_full2 $B > $O/r02_ncu_tol2.log 2>&1
  echo "ncu tol rc=$?" | tee -a $O/r02_rc5.log
fi
