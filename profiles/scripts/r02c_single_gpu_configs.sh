#!/bin/bash
# contract lines of the other single-GPU BASELINE configs on the final tree (N = 1): configs[1] (C2),
# configs[3] (C4), configs[4] (C5), each with its own e2e / cpu_baseline / parity_check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for wl in c2 c5 c4; do
  timeout 60 python bench.py --workload $wl --no-sweep --steps 10 --warmup 3 > gpurun_out/r2g_bench_${wl}_n1.json 2> gpurun_out/r2g_bench_${wl}_n1.err
  echo "$wl rc=$? $(head -c 200 gpurun_out/r2g_bench_${wl}_n1.json)"
done
