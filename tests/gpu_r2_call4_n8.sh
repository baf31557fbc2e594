#!/bin/bash
# round-2 GPU call 4 (8 x B200): multi-rank parity tests, then the scaling lines with parity_check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r2d_pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_sharded.log
tail -6 gpurun_out/r2d_pytest_sharded.log
run() {  # workload gpus steps
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $2 --steps $3 --warmup 5 --workload $1 > gpurun_out/r2d_bench_$1_n$2.json 2> gpurun_out/r2d_bench_$1_n$2.err
  echo "bench $1 n$2 rc=$? $(head -c 400 gpurun_out/r2d_bench_$1_n$2.json)"
}
run c3 8 50
run c3 4 50
run c3 2 50
run c4 8 10
run c5 8 10
run c4 4 10
run c5 4 10
run c4 2 10
run c5 2 10
