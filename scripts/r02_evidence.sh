#!/bin/bash
# Round-2 evidence run (one gpurun call, one GPU): GPU tests, the bench lines of every BASELINE config,
# the ncu launch list of the default bench command and one `--set full` capture of the dominant kernel.
# Every ncu pass runs only after the same command has exited 0 without ncu.
# usage: gpurun --timeout 1500 -- 'bash scripts/r02_evidence.sh'
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02_rc.log
tail -3 $O/r02_pytest.log
timeout 400 python bench.py > $O/r02_bench_cfg3_default.json 2> $O/r02_bench_cfg3_default.err; echo "bench cfg3 rc=$?" | tee -a $O/r02_rc.log
cat $O/r02_bench_cfg3_default.json
for w in ${WORKLOADS:-cfg4 cfg5 cfg4_nni cfg3_nni cfg2}; do
  timeout 300 python bench.py --workload $w --no-cpu > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; echo "bench $w rc=$?" | tee -a $O/r02_rc.log
done
if [ -n "${REF_ARM:-}" ]; then timeout 300 python bench.py --impl reference > $O/r02_bench_cfg3_reference_arm.json 2> $O/r02_bench_ref.err; echo "ref arm rc=$?" | tee -a $O/r02_rc.log; fi
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --quick-check"
if timeout 200 $B > $O/r02_b_plain.json 2>&1; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_cfg3.csv $B > $O/r02_ncu_launches.log 2>&1
  echo "ncu launches rc=$?" | tee -a $O/r02_rc.log
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -c 1 -f -o $O/r02_rows_full $B > $O/r02_ncu_full.log 2>&1
  echo "ncu full rc=$?" | tee -a $O/r02_rc.log
fi
