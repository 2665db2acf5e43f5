#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_hashrf.py tests/test_gpu_multi.py -x -q -k "golden or synthetic or config2 or config3 or config4 or multifurcating or quirk or duplicate or editdist_mask or majority_flip or dedup_table or multi_matches or multi_nonuniform or team_of_one" > $O/r02_pytest_build.log 2>&1; echo "pytest build rc=$?" | tee $O/r02_rc7.log
tail -3 $O/r02_pytest_build.log
if grep -q "rc=0" $O/r02_rc7.log; then
  timeout 200 python bench.py > $O/r02_bench_cfg3_batched.json 2> $O/r02_bench_cfg3_batched.err; echo "bench cfg3 rc=$?" | tee -a $O/r02_rc7.log
  python -c "
import json; d=json.load(open('$O/r02_bench_cfg3_batched.json')); print(d['value'], d['ms_per_step'], d['phases_ms'], d['selfcheck'])"
  timeout 100 python bench.py --workload cfg4 --no-cpu --no-e2e > $O/r02_bench_cfg4_batched.json 2> $O/r02_bench_cfg4_batched.err; echo "bench cfg4 rc=$?" | tee -a $O/r02_rc7.log
  python -c "
import json; d=json.load(open('$O/r02_bench_cfg4_batched.json')); print(d['value'], d['ms_per_step'], d['phases_ms'])"
fi
