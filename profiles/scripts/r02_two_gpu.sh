#!/bin/bash
# N = 2: sharded parity (peer-memory exchange and NCCL), then C3 at N=2 with both exchanges
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x --timeout=200 > gpurun_out/r2o_pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_sharded.log
tail -12 gpurun_out/r2o_pytest_sharded.log
run() {  # workload gpus steps tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $2 --steps $3 --warmup 5 --workload $1 > gpurun_out/r2o_bench_$1_n$2_$4.json 2> gpurun_out/r2o_bench_$1_n$2_$4.err
  echo "bench $1 n$2 $4 rc=$? $(head -c 300 gpurun_out/r2o_bench_$1_n$2_$4.json)"
}
run c3 2 30 p2p
MCL_SHARDED_EXCHANGE=nccl run c3 2 30 nccl
