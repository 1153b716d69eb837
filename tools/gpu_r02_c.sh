#!/bin/bash
# round 2, pass C (2 GPUs): default bench with the warmed collective, C5 at 1.25e8 photons per GPU, new tests
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
tail -c 1800 gpurun_out/r02c_bench_n2.json; tail -3 gpurun_out/r02c_bench_n2.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,P2P NCCL_DEBUG_FILE=gpurun_out/r02c_nccl_%h_%p.log \
  $TR bench.py --gpus 2 --config c5 --c5-photons 2.5e8 --steps 3 > gpurun_out/r02c_c5_n2.json 2> gpurun_out/r02c_c5_n2.err
tail -c 2200 gpurun_out/r02c_c5_n2.json; tail -3 gpurun_out/r02c_c5_n2.err
python -m pytest tests/test_parity_round2.py -m gpu -q -x 2>&1 | tail -5
ls gpurun_out | head -30
