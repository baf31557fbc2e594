#!/bin/bash
# round-2 GPU call 1 (one B200): parity tests, bench line, launch lists + full captures of the new kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json
for wl in c2 gemma3_head; do
  timeout 300 python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2a_plain_$wl.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_${wl}_launches.csv python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2a_ncu_$wl.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 6 -c 2 -o gpurun_out/r2a_${wl}_scan python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2a_ncufull_$wl.log 2>&1
done
ls -la gpurun_out | tail -20
