#!/bin/bash
# round 2, pass Z2: why is the split kernel slower?  ncu of the split C3 kernel
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02z_prof_c3_split \
    python bench.py --config c3 --photons 9999872 --steps 2 > gpurun_out/r02z_ncu_c3.log 2>&1
ls -la gpurun_out | grep r02z
