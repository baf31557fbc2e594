#!/bin/bash
# last single-GPU call of round 2 (final tree): full GPU tests, smoke, bench (ours + reference arm),
# then ncu --set full of the MAIN scan kernel on C2 and C5 (the earlier captures of these two
# workloads caught the threshold-seeding pre-pass: scan_tc_kernel launches alternate seed / main,
# so the capture skips an odd number of them); each ncu run only after its plain run exited 0
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2c_smoke.log 2>&1; tail -1 gpurun_out/r2c_smoke.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; echo "ref rc=$?"
for wl in c2 c5; do
  timeout 300 python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2c_plain_$wl.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 7 -c 1 -o gpurun_out/r2c_${wl}_scan python bench.py --steps 3 --warmup 3 --profile --workload $wl > gpurun_out/r2c_ncufull_$wl.log 2>&1
  echo "ncu $wl rc=$?"
done
ls -la gpurun_out | grep r2c | wc -l
