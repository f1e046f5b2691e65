#!/bin/bash
set -x
O=gpurun_out/r2_ab2
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -5 $O/pytest.txt
P4="--pop 262144 --dim 4096 --objective rosenbrock --strategy best --x0 4.096 --blocks 2"
NLS_B200_LIB=tools/ab/libnls_r1.so python tools/probe_de.py $P4 > $O/cfg4_r1.txt 2>&1
NLS_DE_BULK=0 python tools/probe_de.py $P4 --tag ldg > $O/cfg4_ldg.txt 2>&1
python tools/probe_de.py $P4 --tag bulk_s2k2 > $O/cfg4_bulk.txt 2>&1
for v in s3k2 s2k3 s2k2b3 s3k3b1; do
  NLS_B200_LIB=tools/ab/libnls_b200_$v.so python tools/probe_de.py $P4 > $O/cfg4_$v.txt 2>&1
done
P2="--pop 1048576 --dim 1000 --objective rastrigin --blocks 2"
NLS_B200_LIB=tools/ab/libnls_r1.so python tools/probe_de.py $P2 > $O/cfg2_r1.txt 2>&1
NLS_DE_BULK=0 python tools/probe_de.py $P2 --tag ldg > $O/cfg2_ldg.txt 2>&1
python tools/probe_de.py $P2 --tag bulk_s2k2 > $O/cfg2_bulk.txt 2>&1
for v in s3k2 s2k3 s2k2b3; do
  NLS_B200_LIB=tools/ab/libnls_b200_$v.so python tools/probe_de.py $P2 > $O/cfg2_$v.txt 2>&1
done
PF="--pop 2097152 --dim 4096 --objective rosenbrock --strategy best --x0 4.096 --blocks 2 --gens 8"
python tools/probe_de.py $PF --tag full_bulk_s2k2 > $O/cfg4full_bulk.txt 2>&1
NLS_DE_BULK=0 python tools/probe_de.py $PF --tag full_ldg > $O/cfg4full_ldg.txt 2>&1
NLS_B200_LIB=tools/ab/libnls_b200_s3k2.so python tools/probe_de.py $PF --tag full_s3k2 > $O/cfg4full_s3k2.txt 2>&1
python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 3 > $O/acc_d1000.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --F 0.3 --blocks 2 > $O/acc_d64_f32.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --blocks 2 > $O/d64_f32.txt 2>&1
python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --blocks 2 > $O/d64_f64.txt 2>&1
grep -h "gens" $O/cfg4_*.txt $O/cfg2_*.txt $O/cfg4full_*.txt $O/acc_*.txt $O/d64_*.txt
for lib in tools/ab/libnls_r1.so nlsolver_b200/libnls_b200.so; do
  NLS_B200_LIB=$lib python tests/tools/quick_time_pso.py 4194304 64 20 0 0 0
  NLS_B200_LIB=$lib python tests/tools/quick_time_pso.py 4194304 64 20 0 0 1
  NLS_B200_LIB=$lib python tests/tools/quick_time_pso.py 4194304 64 20 0 1 0
  NLS_B200_LIB=$lib python tests/tools/quick_time_pso.py 4194304 64 20 0 1 1
  NLS_B200_LIB=$lib python tests/tools/quick_time_pso.py 2097152 256 20 3 1 1
  NLS_B200_LIB=$lib python tests/tools/quick_time_pso.py 2097152 128 20 0 0 1
done 2>&1 | grep -v Warning
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
tail -c 600 $O/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_ab2/bench.json').read().strip().splitlines()[-1])
    print(json.dumps({k:d[k] for k in ('value','ms_per_step','roofline','e2e')}, indent=1)[:1800])
except Exception as e:
    print('bench parse failed', e)
PY
