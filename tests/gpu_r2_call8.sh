#!/bin/bash
# after the panel scan: full GPU tests, full bench (sweep incl. gemma3_eval), C1 launch list + full ncu capture of the panel kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
timeout 300 python tests/gpu_panel_timing.py > gpurun_out/r2h_panel_timing.log 2>&1; echo "timing rc=$?"; grep -E "panel|stamps|CTA start" gpurun_out/r2h_panel_timing.log
timeout 300 python bench.py --workload c1 --no-sweep > gpurun_out/r2h_bench_c1.json 2> gpurun_out/r2h_bench_c1.err; echo "bench c1 rc=$?"; tail -3 gpurun_out/r2h_bench_c1.err
timeout 1200 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench.err
for wl in c1 gemma3_eval; do
  timeout 300 python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2h_plain_$wl.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_${wl}_launches.csv python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2h_ncu_$wl.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:panel_scan_kernel -s 4 -c 1 -o gpurun_out/r2h_${wl}_panel python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2h_ncufull_$wl.log 2>&1
done
ls -la gpurun_out | grep r2h
