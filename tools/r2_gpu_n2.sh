#!/bin/bash
# two GPUs of one box: the multi-GPU tests, the device-group timings, bench.py under torchrun
set -x
O=gpurun_out/r2_n2
mkdir -p $O
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -15 $O/pytest.txt
python tools/bench_small.py > $O/bench_small.json 2> $O/bench_small.err; tail -3 $O/bench_small.err
python -c "
import json
d=json.loads(open('$O/bench_small.json').read().strip().splitlines()[-1]); print(json.dumps(d,indent=1))"
python tools/bench_group.py --gpus 1 > $O/group_n1.json 2> $O/group_n1.err; cat $O/group_n1.json; tail -3 $O/group_n1.err
python tools/bench_group.py --gpus 2 > $O/group_n2.json 2> $O/group_n2.err; cat $O/group_n2.json; tail -3 $O/group_n2.err
g++ -std=c++17 -O2 -Iinclude examples/example_multi_gpu.cpp -Lnlsolver_b200 -lnls_b200 -Wl,-rpath,$PWD/nlsolver_b200 -o /tmp/example_multi_gpu && /tmp/example_multi_gpu 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 1500 $O/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_n2/bench_n2.json').read().strip().splitlines()[-1])
    print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}))
    e=d['extra']
    print(json.dumps(e['multi_gpu_parity'],indent=1))
    print(json.dumps(e['configs']['config3_pso_accelerated_ackley_d256'],indent=1))
    print(json.dumps(e['configs']['config4_island_de_best_rosenbrock_d4096'],indent=1)[:900])
except Exception as ex:
    print('bench parse failed', ex)
PY
