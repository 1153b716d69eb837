for p in 0 1; do
MXB_JIT_PIPE=$p ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/prof_r01_jit_pipe$p python bench.py --no-cpu --no-e2e --steps 2 > gpurun_out/ncu_jit_pipe$p.log 2>&1
done
