#!/bin/bash
# session 2, call C: branch-free div / sqrt in rnorm + branch-free Ackley terms: GPU tests, config-3 and SANN timings
set -x
O=gpurun_out/r2_s2c
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -6 $O/pytest.txt
for i in 1 2; do
python tests/tools/quick_time_pso.py 2097152 256 20 3 1 1 2>&1 | grep -v Warn
python tests/tools/quick_time_pso.py 2097152 256 20 3 1 0 2>&1 | grep -v Warn
python tests/tools/quick_time_pso.py 4194304 64 20 0 1 1 2>&1 | grep -v Warn
python tests/tools/quick_time_pso.py 4194304 64 20 0 1 0 2>&1 | grep -v Warn
python tests/tools/quick_time_pso.py 4194304 64 20 3 0 1 2>&1 | grep -v Warn
done
python tools/bench_sann.py > $O/sann.json 2> $O/sann.err; tail -c 900 $O/sann.json
