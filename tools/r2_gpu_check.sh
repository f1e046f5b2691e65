#!/bin/bash
# parity tests + per-kernel probes of the shapes the verdict names (run under gpurun, one GPU)
set -x
O=gpurun_out/r2_check
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -5 $O/pytest.txt
python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.8 --blocks 2 > $O/probe_sphere_d1000_F08.txt 2>&1
python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 4 > $O/probe_sphere_d1000_F02.txt 2>&1
python tools/probe_de.py --pop 1048576 --dim 1000 --objective rastrigin --blocks 2 > $O/probe_rastrigin_d1000.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --blocks 3 > $O/probe_sphere_d64_f64.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --blocks 3 > $O/probe_sphere_d64_f32.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --F 0.3 --blocks 3 > $O/probe_sphere_d64_f32_F03.txt 2>&1
python tools/probe_de.py --pop 2097152 --dim 128 --objective sphere --blocks 2 > $O/probe_sphere_d128_f64.txt 2>&1
python tools/probe_de.py --pop 262144 --dim 4096 --objective rosenbrock --strategy best --x0 4.096 --blocks 3 > $O/probe_cfg4_2p18.txt 2>&1
python tools/probe_de.py --pop 1024 --dim 64 --objective sphere --blocks 3 > $O/probe_sphere_d64_p1k.txt 2>&1
grep -h "gens" $O/probe_*.txt | head -60
