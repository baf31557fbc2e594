#!/bin/bash
# round-2 GPU call 2 (one B200): full parity suite, bench line, C2 re-profile
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
for wl in c2; do
  timeout 300 python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2b_plain_$wl.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_${wl}_launches.csv python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2b_ncu_$wl.log 2>&1
done
