#!/bin/bash
# session 2, one GPU, the library as committed: tests, the bench lines for
# profiles/ (b200 arm with extras, reference arm), the small-shape and SANN benches, ncu of the changed kernels
set -x
O=gpurun_out/r2_final
N=gpurun_out/r2_final_ncu
mkdir -p $O $N
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -4 $O/pytest.txt
python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.json
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.err
python tools/bench_small.py > $O/bench_small.json 2> $O/bench_small.err; tail -2 $O/bench_small.err
python tools/bench_sann.py > $O/sann.json 2> $O/sann.err; tail -2 $O/sann.err
mkdir -p gpurun_out/r2_sweep; python tools/sweep.py --big > gpurun_out/r2_sweep/sweep.md 2> gpurun_out/r2_sweep/sweep.err; tail -2 gpurun_out/r2_sweep/sweep.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
NCU="ncu --set full --clock-control none --import-source on"
run() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $N/$name.plain.log 2>&1 && $NCU -k regex:"$rx" -s $skip -c $cnt -o $N/$name "$@" > $N/$name.ncu.log 2>&1
  echo "== $name rc=$?"
}
run repair_d64_f32 "de_repair_kernel" 3 1 python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --F 0.3 --blocks 1 --gens 6
run repair_d1000 "de_repair_kernel" 3 1 python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 1 --gens 6
run pso_accel_config3_final "pso_move_kernel" 4 1 python tests/tools/quick_time_pso.py 2097152 256 3 3 1 1
run pso_vanilla_f32_d64_final "pso_move_kernel" 4 1 python tests/tools/quick_time_pso.py 4194304 64 3 0 0 0
for f in $N/*.ncu-rep; do
  n=${f%.ncu-rep}
  ncu -i $f --page raw --csv > $n.raw.csv 2>/dev/null
  ncu -i $f --page source --csv > $n.source.csv 2>/dev/null
  rm -f $f
done
gzip -9f $N/*.source.csv
du -sh gpurun_out
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final/bench_n1.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks','cpu_baseline')})[:900])
print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e'].get('cold_ms'))
e=d['extra']
for w in e['accepting']['windows']+e['accepting']['high_acceptance']: print({k:w[k] for k in ('window','accepted_fraction','ms_per_generation','k2_ms','k2r_ms','frac_of_measured_hbm')})
c=e['configs']
print('config3', c['config3_pso_accelerated_ackley_d256']['none']['ms_per_generation'])
c4=c['config4_island_de_best_rosenbrock_d4096']; print('config4', {k:c4.get(k) for k in ('ms_per_generation','frac_of_measured_hbm','clocks','unavailable')})
for r in c['config5_sweep_d64']: print(r['solver'],r['dtype'],round(r['ms_per_generation'],3),'%.3g'%r['agent_evals_per_sec'],round(r['frac_of_measured_hbm'],3))
s=json.loads(open('gpurun_out/r2_final/bench_small.json').read().strip().splitlines()[-1]); print(json.dumps(s['config1_readme_de'])[:500])
PY
