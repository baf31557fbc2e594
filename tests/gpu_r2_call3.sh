#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
timeout 300 python tests/gpu_rowkernels.py > gpurun_out/r2c_rowkernels.log 2>&1; cat gpurun_out/r2c_rowkernels.log
timeout 300 python tests/gpu_rowkernels.py --quick > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_mean -s 3 -c 1 -o gpurun_out/r2c_gather_mean python tests/gpu_rowkernels.py --quick > gpurun_out/r2c_ncu_gm.log 2>&1
for wl in c1; do
  timeout 300 python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2c_plain_$wl.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_${wl}_launches.csv python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2c_ncu_$wl.log 2>&1
done
