#!/bin/bash
# final single-GPU call of round 2: full GPU tests, smoke, bench (ours + reference arm), row kernels,
# launch lists + full ncu captures of the dominant kernels (each only after its plain run exited 0)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 1200 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --workload c1 --no-sweep > gpurun_out/r2z_bench_c1.json 2> gpurun_out/r2z_bench_c1.err; echo "bench c1 rc=$?"
timeout 300 python tests/gpu_rowkernels.py --quick > gpurun_out/r2z_rowkernels.log 2>&1
timeout 300 python tests/gpu_phases.py > gpurun_out/r2z_phases.log 2>&1; cat gpurun_out/r2z_phases.log
for wl in c3 c2 c1; do
  timeout 300 python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2z_plain_$wl.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2z_${wl}_launches.csv python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2z_ncu_$wl.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 6 -c 1 -o gpurun_out/r2z_c2_scan python bench.py --steps 3 --warmup 3 --profile --workload c2 > gpurun_out/r2z_ncufull_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:panel_scan_kernel -s 4 -c 1 -o gpurun_out/r2z_c1_panel python bench.py --steps 3 --warmup 3 --profile --workload c1 > gpurun_out/r2z_ncufull_c1.log 2>&1
ls -la gpurun_out | grep r2z | wc -l
