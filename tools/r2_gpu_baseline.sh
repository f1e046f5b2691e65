#!/bin/bash
# round-2 baseline measurements of the round-1 kernels (run under gpurun, one GPU)
set -x
O=gpurun_out/r2_base
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 5 > $O/probe_sphere_d1000_f64.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --blocks 4 > $O/probe_sphere_d64_f64.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --blocks 4 > $O/probe_sphere_d64_f32.txt 2>&1
python tools/probe_de.py --pop 262144 --dim 4096 --objective rosenbrock --strategy best --x0 4.096 --blocks 3 > $O/probe_cfg4_2p18.txt 2>&1
python tools/probe_de.py --pop 1024 --dim 64 --objective sphere --blocks 3 > $O/probe_sphere_d64_p1k.txt 2>&1
python tests/tools/quick_time_pso.py 4194304 64 20 0 0 0 > $O/pso_vanilla_d64_f32.txt 2>&1
python tests/tools/quick_time_pso.py 4194304 64 20 0 0 1 > $O/pso_vanilla_d64_f64.txt 2>&1
python tests/tools/quick_time_pso.py 2097152 256 20 3 1 1 > $O/pso_accel_cfg3.txt 2>&1
# ncu: full sets of the kernels the verdict names (plain runs above exited 0 or this is skipped)
CMD32="python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --blocks 1 --gens 6"
$CMD32 > $O/plain32.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:de_generation_kernel -s 3 -c 1 -f -o $O/de_gen_f32_d64 $CMD32 > $O/ncu32.log 2>&1
$CMD32 > $O/plain32r.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:de_repair_kernel -s 3 -c 1 -f -o $O/de_repair_f32_d64 $CMD32 > $O/ncu32r.log 2>&1
CMD4="python tools/probe_de.py --pop 262144 --dim 4096 --objective rosenbrock --strategy best --x0 4.096 --blocks 1 --gens 4"
$CMD4 > $O/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:de_generation_kernel -s 2 -c 1 -f -o $O/de_gen_cfg4 $CMD4 > $O/ncu4.log 2>&1
CMDA="python tools/probe_de.py --pop 262144 --dim 1000 --objective sphere --F 0.2 --blocks 1 --gens 24"
$CMDA > $O/plainA.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:de_repair_kernel -s 22 -c 1 -f -o $O/de_repair_d1000 $CMDA > $O/ncuA.log 2>&1
ls -la $O
