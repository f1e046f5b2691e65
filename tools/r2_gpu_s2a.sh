#!/bin/bash
# session 2, call A (two GPUs): the whole GPU test suite incl. the torchrun multi-GPU check, bench.py at N = 2 and N = 1
set -x
O=gpurun_out/r2_s2a
mkdir -p $O
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -15 $O/pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/tools/multi_gpu_check.py > $O/multi_gpu_check_n2.txt 2>&1
cat $O/multi_gpu_check_n2.txt | grep -v Warning | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 1500 $O/bench_n2.err
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 600 $O/bench_n1.err
python - <<'PY'
import json
for n in (2, 1):
    try:
        d=json.loads(open(f'gpurun_out/r2_s2a/bench_n{n}.json').read().strip().splitlines()[-1])
        print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}), d['roofline']['frac'], d['e2e']['value'])
        e=d['extra']
        if 'multi_gpu_parity' in e: print(json.dumps(e['multi_gpu_parity'],indent=1)[:600])
        c4=e['configs']['config4_island_de_best_rosenbrock_d4096']
        print('config4', {k:c4.get(k) for k in ('ms_per_generation','frac_of_measured_hbm','k2_frac_of_measured_hbm','unavailable')})
        c3=e['configs']['config3_pso_accelerated_ackley_d256']
        print('config3', {k:(v['ms_per_generation'] if isinstance(v,dict) else None) for k,v in c3.items() if isinstance(v,dict)})
    except Exception as ex:
        print('bench parse failed', n, ex)
PY
