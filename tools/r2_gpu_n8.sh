#!/bin/bash
# eight GPUs of one box: the multi-GPU check and bench.py under torchrun at N = 8 and N = 4 (what the driver's scaling run does)
set -x
O=gpurun_out/r2_n8
mkdir -p $O
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tests/tools/multi_gpu_check.py > $O/multi_gpu_check_n8.txt 2>&1
grep -v Warning $O/multi_gpu_check_n8.txt | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 600 $O/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > $O/bench_n4.json 2> $O/bench_n4.err
tail -c 300 $O/bench_n4.err
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 --no-extras --skip-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python tools/bench_group.py --gpus 8 > $O/group_n8.json 2> $O/group_n8.err; tail -2 $O/group_n8.err
python - <<'PY'
import json
for n in (8, 4, 1):
    try:
        d=json.loads(open(f'gpurun_out/r2_n8/bench_n{n}.json').read().strip().splitlines()[-1])
        print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}), d['roofline']['frac'], d['clocks'])
        e=d.get('extra')
        if e:
            print(json.dumps(e['multi_gpu_parity'],indent=1)[:300])
            c3=e['configs']['config3_pso_accelerated_ackley_d256']
            print('config3', {k:(v['ms_per_generation'], v['agent_evals_per_sec']) for k,v in c3.items() if isinstance(v,dict)})
            c4=e['configs']['config4_island_de_best_rosenbrock_d4096']
            print('config4', {k:c4.get(k) for k in ('ms_per_generation','agent_evals_per_sec','frac_of_measured_hbm','unavailable')})
            for r in e['configs']['config5_sweep_d64']: print(r['solver'],r['dtype'],round(r['ms_per_generation'],3),'%.3g'%r['agent_evals_per_sec'],round(r['frac_of_measured_hbm'],3))
    except Exception as ex:
        print('bench parse failed', n, ex)
try:
    d=json.loads(open('gpurun_out/r2_n8/group_n8.json').read().strip().splitlines()[-1])
    print('group n8', [(r['particles_per_gpu'], round(r['us_per_generation'],1)) for r in d['sharded_pso_accelerated_ackley_d256']], [(r['agents_per_island'], r['dim'], round(r['us_per_generation'],1)) for r in d['de_islands']])
except Exception as ex:
    print('group parse failed', ex)
PY
