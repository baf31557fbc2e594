#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
timeout 1200 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench.err
timeout 300 python tests/gpu_rowkernels.py > gpurun_out/r2f_rowkernels.log 2>&1; grep -E "D=3584|D=1152 Q=65536 ids/row in \[1,5\) normalize=True" gpurun_out/r2f_rowkernels.log
timeout 600 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "ref rc=$?"
for wl in c3 c4 c5; do
  timeout 300 python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2f_plain_$wl.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2f_${wl}_launches.csv python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2f_ncu_$wl.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 4 -c 1 -o gpurun_out/r2f_${wl}_scan python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2f_ncufull_$wl.log 2>&1
done
ls -la gpurun_out | grep r2f
