#!/bin/bash
# session 2, call E (two GPUs): everything multi-GPU with the library as it stands
set -x
O=gpurun_out/r2_s2e
mkdir -p $O
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -8 $O/pytest.txt
python tools/bench_group.py --gpus 1 > $O/group_n1.json 2> $O/group_n1.err; cat $O/group_n1.json; tail -3 $O/group_n1.err
python tools/bench_group.py --gpus 2 > $O/group_n2.json 2> $O/group_n2.err; cat $O/group_n2.json; tail -3 $O/group_n2.err
g++ -std=c++17 -O2 -Iinclude examples/example_multi_gpu.cpp -Lnlsolver_b200 -lnls_b200 -Wl,-rpath,$PWD/nlsolver_b200 -o /tmp/example_multi_gpu && /tmp/example_multi_gpu 2 > $O/example_multi_gpu.txt 2>&1; cat $O/example_multi_gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/tools/multi_gpu_check.py > $O/multi_gpu_check_n2.txt 2>&1
grep -v Warning $O/multi_gpu_check_n2.txt | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 800 $O/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_s2e/bench_n2.json').read().strip().splitlines()[-1])
    print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}), d['roofline']['frac'], d['e2e']['value'])
    e=d['extra']
    print(json.dumps(e['multi_gpu_parity'],indent=1)[:400])
    c3=e['configs']['config3_pso_accelerated_ackley_d256']
    print('config3', {k:(v['ms_per_generation'] if isinstance(v,dict) else None) for k,v in c3.items() if isinstance(v,dict)})
    for w in e['accepting']['windows']+e['accepting']['high_acceptance']: print({k:w[k] for k in ('window','accepted_fraction','ms_per_generation','k2_ms','k2r_ms','frac_of_measured_hbm')})
except Exception as ex:
    print('bench parse failed', ex)
PY
