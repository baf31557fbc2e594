#!/bin/bash
# ncu --set full pages of the hand-written kernels that had none yet (final tree): the grad-epilogue
# scan + the two tcgen05 gradient GEMMs of the backward (section 8 f1), and the k = 1 epilogue of the scan
# at the reference-native LM-head shape (a4/a5).  Each ncu run only after its plain run exited 0.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tests/gpu_bwd_profile.py > gpurun_out/r2d_plain_bwd.log 2>&1 && cat gpurun_out/r2d_plain_bwd.log &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'scan_tc_kernel|gemm_tc_kernel' -s 76 -c 3 -o gpurun_out/r2d_bwd_kernels python tests/gpu_bwd_profile.py > gpurun_out/r2d_ncufull_bwd.log 2>&1
echo "ncu bwd rc=$?"
timeout 200 python bench.py --steps 3 --warmup 3 --profile --workload gemma3_head > gpurun_out/r2d_plain_gemma3_head.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 4 -c 1 -o gpurun_out/r2d_gemma3_head_scan python bench.py --steps 3 --warmup 3 --profile --workload gemma3_head > gpurun_out/r2d_ncufull_gemma3_head.log 2>&1
echo "ncu gemma3_head rc=$?"
