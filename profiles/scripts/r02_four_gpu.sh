#!/bin/bash
# 4 x B200, final build: C3 at N = 4 and N = 2 (peer-memory exchange), plus the N = 1 line on the same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() {  # workload gpus steps tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $2 --steps $3 --warmup 5 --workload $1 > gpurun_out/r3b_bench_$1_n$2_$4.json 2> gpurun_out/r3b_bench_$1_n$2_$4.err
  echo "bench $1 n$2 $4 rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3b_bench_$1_n$2_$4.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["parity_check"]["ok"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("no json:", e)
PY
)"
}
run c3 4 50 p2p
run c3 2 50 p2p
timeout 300 python bench.py --no-sweep --steps 50 > gpurun_out/r3b_bench_c3_n1.json 2> gpurun_out/r3b_bench_c3_n1.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r3b_bench_c3_n1.json").read().strip().splitlines()[-1])
print("n1", round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]))
PY
