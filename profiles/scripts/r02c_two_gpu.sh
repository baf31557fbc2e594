#!/bin/bash
# N = 2 after the parity-counter change of the peer-memory exchange: sharded parity test (peer memory,
# incl. 12 local-rows batches + mixed full / local scans), then C3 at N=2 through bench.py
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 200 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x --timeout=150 -k "2-auto" > gpurun_out/r2c_pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_sharded.log
tail -4 gpurun_out/r2c_pytest_sharded.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_c3_n2.json 2> gpurun_out/r2c_bench_c3_n2.err
echo "bench c3 n2 rc=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c_bench_c3_n2.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check']['ok'])")"
