#!/bin/bash
# round 2, pass D (8 GPUs): default bench at N=8 and N=4 with the warmed collective, C5 at 1e9 photons
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02d_topo.txt 2>&1
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02d_bench_n$N.json 2> gpurun_out/r02d_bench_n$N.err
tail -c 1500 gpurun_out/r02d_bench_n$N.json; tail -2 gpurun_out/r02d_bench_n$N.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL NCCL_DEBUG_FILE=gpurun_out/r02d_nccl_%h_%p.log \
  $TR bench.py --gpus 8 --config c5 --c5-photons 1e9 --steps 3 > gpurun_out/r02d_c5_n8.json 2> gpurun_out/r02d_c5_n8.err
tail -c 2500 gpurun_out/r02d_c5_n8.json; tail -3 gpurun_out/r02d_c5_n8.err
ls gpurun_out/r02d_nccl* | head -3; rm -f $(ls gpurun_out/r02d_nccl* | tail -n +3)
