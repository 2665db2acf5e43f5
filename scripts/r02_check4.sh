#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_agglo.py -x -q > $O/r02_pytest_agglo2.log 2>&1; echo "pytest agglo rc=$?" | tee -a $O/r02_rc4.log
tail -5 $O/r02_pytest_agglo2.log
timeout 420 python -m pytest tests/test_gpu_hashrf.py -x -q -k "tolerant or flat_clusters or kat or one_shot" > $O/r02_pytest_tol2.log 2>&1; echo "pytest tolerant rc=$?" | tee -a $O/r02_rc4.log
tail -3 $O/r02_pytest_tol2.log
timeout 100 python scripts/wordop_peak.py > $O/wordop_peak.json 2> $O/wordop_peak.err; echo "wordop rc=$?" | tee -a $O/r02_rc4.log
cat $O/wordop_peak.json
mkdir -p profiles; cp $O/wordop_peak.json profiles/wordop_peak.json
timeout 300 python bench.py --workload cfg5 --no-cpu > $O/r02_bench_cfg5_v3.json 2> $O/r02_bench_cfg5_v3.err; echo "bench cfg5 rc=$?" | tee -a $O/r02_rc4.log
python -c "
import json; d=json.load(open('$O/r02_bench_cfg5_v3.json')); print(d['value'], d['ms_per_step'], d['roofline'])"
timeout 400 python scripts/consumers_probe.py 20000 200 8 4 --dendrogram > $O/r02_consumers_probe2.json 2> $O/r02_consumers_probe2.err; echo "probe rc=$?" | tee -a $O/r02_rc4.log
cat $O/r02_consumers_probe2.json; tail -3 $O/r02_consumers_probe2.err
