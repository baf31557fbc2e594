#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shims.py -m gpu -q -x --timeout=120 -k "gather or a3 or multi_token or imagenet" > gpurun_out/r2i_pytest_gather.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_gather.log
tail -6 gpurun_out/r2i_pytest_gather.log
timeout 300 python tests/gpu_rowkernels.py > gpurun_out/r2i_rowkernels.log 2>&1; echo "rowkernels rc=$?"; cat gpurun_out/r2i_rowkernels.log
