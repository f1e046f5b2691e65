#!/bin/bash
# session 2, call D: A/B of the accelerated-PSO U = 2 trip and the batched repair scan
set -x
O=gpurun_out/r2_s2d
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -6 $O/pytest.txt
for u in 1 2 1 2; do
NLS_PSO_ACCEL_U=$u python tests/tools/quick_time_pso.py 2097152 256 20 3 1 1 2>&1 | grep -v Warn
NLS_PSO_ACCEL_U=$u python tests/tools/quick_time_pso.py 2097152 256 20 3 1 0 2>&1 | grep -v Warn
NLS_PSO_ACCEL_U=$u python tests/tools/quick_time_pso.py 1048576 1000 10 2 1 1 2>&1 | grep -v Warn
done
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --F 0.3 --blocks 2
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --F 0.3 --blocks 2
python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 3
python tools/probe_de.py --pop 65536 --dim 64 --objective sphere --F 0.3 --blocks 2
