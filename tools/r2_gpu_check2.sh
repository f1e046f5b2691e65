#!/bin/bash
set -x
O=gpurun_out/r2_check2
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -15 $O/pytest.txt
python tools/bench_small.py > $O/bench_small.json 2> $O/bench_small.err
tail -3 $O/bench_small.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_check2/bench_small.json').read().strip().splitlines()[-1])
    print(json.dumps(d, indent=1))
except Exception as e:
    print('bench_small parse failed', e)
PY
g++ -std=c++17 -O2 -Iinclude examples/example_multi_gpu.cpp -Lnlsolver_b200 -lnls_b200 -Wl,-rpath,$PWD/nlsolver_b200 -o /tmp/example_multi_gpu && /tmp/example_multi_gpu 1
NLS_DE_BULK=0 python tools/probe_de.py --pop 262144 --dim 1000 --objective rastrigin --blocks 1 --gens 6 --tag ldg > $O/plain_ldg.log 2>&1 && NLS_DE_BULK=0 ncu --set full --clock-control none --import-source on -k regex:de_generation_kernel -s 3 -c 1 -f -o $O/de_gen_ldg_d1000 python tools/probe_de.py --pop 262144 --dim 1000 --objective rastrigin --blocks 1 --gens 6 --tag ldg > $O/ncu_ldg.log 2>&1
python tools/probe_de.py --pop 262144 --dim 1000 --objective rastrigin --blocks 1 --gens 6 > $O/plain_bulk.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:de_generation_bulk_kernel -s 3 -c 1 -f -o $O/de_gen_bulk_d1000 python tools/probe_de.py --pop 262144 --dim 1000 --objective rastrigin --blocks 1 --gens 6 > $O/ncu_bulk.log 2>&1
ls -la $O
