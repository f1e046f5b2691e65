#!/bin/bash
set -x
O=gpurun_out/r2_check3
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -15 $O/pytest.txt
python tools/bench_small.py > $O/bench_small.json 2> $O/bench_small.err; tail -3 $O/bench_small.err
python -c "
import json
d=json.loads(open('$O/bench_small.json').read().strip().splitlines()[-1]); print(json.dumps(d['config1_readme_de'],indent=1)); print([ (r['solver'],r['pop'],r['dim'],round(r['one_launch_us_per_generation'],1),round(r['graph_replay_us_per_generation'],1)) for r in d['sweep_d64_small']])"
python tools/bench_group.py --gpus 1 > $O/group_n1.json 2> $O/group_n1.err; cat $O/group_n1.json; tail -3 $O/group_n1.err
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
tail -c 600 $O/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_check3/bench.json').read().strip().splitlines()[-1])
    print(json.dumps({k:d[k] for k in ('value','ms_per_step','roofline','e2e')}, indent=1)[:1500])
    for r in d['extra']['configs']['config5_sweep_d64']: print(r['solver'],r['dtype'],round(r['ms_per_generation'],3),round(r['frac_of_measured_hbm'],3))
    print(json.dumps(d['extra']['accepting']['windows'],indent=1)[:1200])
except Exception as e:
    print('bench parse failed', e)
PY
