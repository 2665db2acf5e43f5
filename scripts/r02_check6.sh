#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_hashrf.py tests/test_gpu_agglo.py -x -q -k "tolerant or kat or one_shot" > $O/r02_pytest_tol3.log 2>&1; echo "pytest tolerant rc=$?" | tee -a $O/r02_rc6.log
tail -3 $O/r02_pytest_tol3.log
timeout 300 python bench.py --workload cfg5 --no-cpu > $O/r02_bench_cfg5_v4.json 2> $O/r02_bench_cfg5_v4.err; echo "bench cfg5 rc=$?" | tee -a $O/r02_rc6.log
python -c "
import json; d=json.load(open('$O/r02_bench_cfg5_v4.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
