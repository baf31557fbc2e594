#!/bin/bash
# 8 x B200: multi-rank parity (peer-memory exchange and NCCL), then the scaling lines of the final build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 420 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x --timeout=200 -k "8-auto or 8-nccl or 4-auto or 2-auto" > gpurun_out/r2p_pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_sharded.log
tail -5 gpurun_out/r2p_pytest_sharded.log
run() {  # workload gpus steps tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $2 --steps $3 --warmup 5 --workload $1 > gpurun_out/r2p_bench_$1_n$2_$4.json 2> gpurun_out/r2p_bench_$1_n$2_$4.err
  echo "bench $1 n$2 $4 rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2p_bench_$1_n$2_$4.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["parity_check"]["ok"])
except Exception as e:
    print("no json:", e)
PY
)"
}
run c3 8 50 p2p
MCL_SHARDED_EXCHANGE=nccl run c3 8 50 nccl
run c3 4 50 p2p
run c4 8 10 p2p
run c5 8 10 p2p
