#!/bin/bash
# Second round-2 GPU call: the new agglomeration kernel, the rewritten tolerant kernel (+ repeated-label form), cfg5 bench,
# consumers probe with the device agglomeration.
set -u
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_agglo.py -x -q > $O/r02_pytest_agglo.log 2>&1; echo "pytest agglo rc=$?" | tee -a $O/r02_rc2.log
tail -15 $O/r02_pytest_agglo.log
timeout 420 python -m pytest tests/test_gpu_hashrf.py -x -q -k "tolerant or flat_clusters or kat or one_shot" > $O/r02_pytest_tol.log 2>&1; echo "pytest tolerant rc=$?" | tee -a $O/r02_rc2.log
tail -5 $O/r02_pytest_tol.log
timeout 300 python bench.py --workload cfg5 --no-cpu > $O/r02_bench_cfg5_v2.json 2> $O/r02_bench_cfg5_v2.err; echo "bench cfg5 rc=$?" | tee -a $O/r02_rc2.log
cat $O/r02_bench_cfg5_v2.json | head -c 600; echo
timeout 400 python scripts/consumers_probe.py 20000 200 8 4 --host --dendrogram > $O/r02_consumers_probe.json 2> $O/r02_consumers_probe.err; echo "probe rc=$?" | tee -a $O/r02_rc2.log
cat $O/r02_consumers_probe.json; tail -3 $O/r02_consumers_probe.err
