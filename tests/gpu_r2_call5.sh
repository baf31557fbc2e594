#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shims.py -m gpu -q -x > gpurun_out/r2e_pytest_shims.log 2>&1; echo "pytest shims rc=$?" >> gpurun_out/r2e_pytest_shims.log
tail -30 gpurun_out/r2e_pytest_shims.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_shims.py > gpurun_out/r2e_pytest_rest.log 2>&1; echo "pytest rest rc=$?" >> gpurun_out/r2e_pytest_rest.log
tail -6 gpurun_out/r2e_pytest_rest.log
timeout 300 python tests/gpu_rowkernels.py > gpurun_out/r2e_rowkernels.log 2>&1; cat gpurun_out/r2e_rowkernels.log
timeout 300 python bench.py --workload c1 --no-sweep --steps 40 > gpurun_out/r2e_bench_c1.json 2> gpurun_out/r2e_bench_c1.err; echo "bench c1 rc=$?"; head -c 600 gpurun_out/r2e_bench_c1.json
