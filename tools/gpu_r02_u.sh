#!/bin/bash
# round 2, pass U (8 GPUs): weak scaling at N = 8 / 4 / 2 and C5 at 1e9 photons with the final kernels
mkdir -p gpurun_out
for N in 8 4 2; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-api 2>gpurun_out/r02u_bench_n$N.err | tail -1 > gpurun_out/r02u_bench_n$N.json
python -c "import json; d=json.load(open('gpurun_out/r02u_bench_n$N.json')); print('N=$N value %.4g ms %.4f collective_ms %.3f e2e %.4g lean %.4g frac %.3f'%(d['value'], d['ms_per_step'], d['collective_ms'], d['e2e']['value'], d['e2e']['lean']['value'], d['roofline']['frac']))"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620"
$TR bench.py --gpus 8 --config c5 --c5-photons 1e9 --steps 3 2> gpurun_out/r02u_c5_n8.err | tail -1 > gpurun_out/r02u_c5_n8.json
python -c "import json; d=json.load(open('gpurun_out/r02u_c5_n8.json')); print('C5 N=8', d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline']['trace_only']['frac'], d['gathered_on_rank0'])"
