#!/bin/bash
# session 2, call F: coarse-bitmap scan filter (shipped) and the repair's steps in flight for long rows (variants)
set -x
O=gpurun_out/r2_s2f
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -6 $O/pytest.txt
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --F 0.3 --blocks 2
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --F 0.3 --blocks 2
python tools/probe_de.py --pop 65536 --dim 64 --objective sphere --F 0.3 --blocks 2
for rep in 1 2; do
python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 3
NLS_B200_LIB=tools/ab/libnls_b200_ru2.so python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 3
NLS_B200_LIB=tools/ab/libnls_b200_ru4.so python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 3
done
python tools/probe_de.py --pop 262144 --dim 4096 --objective sphere --F 0.2 --strategy best --blocks 2
NLS_B200_LIB=tools/ab/libnls_b200_ru2.so python tools/probe_de.py --pop 262144 --dim 4096 --objective sphere --F 0.2 --strategy best --blocks 2
NLS_B200_LIB=tools/ab/libnls_b200_ru4.so python tools/probe_de.py --pop 262144 --dim 4096 --objective sphere --F 0.2 --strategy best --blocks 2
