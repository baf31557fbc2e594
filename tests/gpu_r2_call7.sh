#!/bin/bash
# panel scan (one-launch small-batch path): parity, then C1 timings old vs new
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
timeout 300 python tests/gpu_panel_timing.py > gpurun_out/r2g_panel_timing.log 2>&1; echo "timing rc=$?"; cat gpurun_out/r2g_panel_timing.log
timeout 300 python bench.py --workload c1 --no-sweep > gpurun_out/r2g_bench_c1.json 2> gpurun_out/r2g_bench_c1.err; echo "bench c1 rc=$?"; tail -3 gpurun_out/r2g_bench_c1.err
