#!/bin/bash
# session 2, call B (one GPU): ncu evidence for round 2 — launch list of the bench step and one --set full capture of
# each kernel family at its bench shape.  Every command runs plain first (exit 0) before it runs under ncu.
set -x
O=gpurun_out/r2_s2b
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
run() {  # run <name> <kernel regex> <skip> <count> <cmd...>
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/$name.plain.log 2>&1 && $NCU -k regex:"$rx" -s $skip -c $cnt -o $O/$name "$@" > $O/$name.ncu.log 2>&1
  echo "== $name rc=$?"; tail -2 $O/$name.plain.log; grep -E "==PROF==.*(Report|Disconnected)|not profiled" $O/$name.ncu.log | tail -2
}
B="python bench.py --steps 5 --warmup 5 --skip-cpu-baseline --no-extras"
$B > $O/bench_plain.json 2> $O/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $B > $O/launches.out 2>&1
echo "== launches rc=$?"
# K2 at the headline shape (config 2): the TMA-staged generation kernel; 5 warm-up generations skipped
run k2_config2 "de_generation" 6 2 $B
# the accepting regime: short-row K2 (fp32 d=64) and the repair kernel, generation 4 of F=0.3 Sphere
run d64_f32_acc "de_generation_kernel|de_repair_kernel" 6 2 python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --F 0.3 --blocks 1 --gens 6
# config-2 shape with accepted trials (F=0.2 Sphere d=1000): repair kernel on long rows
run d1000_acc "de_repair_kernel" 3 1 python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere --F 0.2 --blocks 1 --gens 6
# config 4 shape (DE-best Rosenbrock d=4096), a quarter of the population
run k2_config4 "de_generation" 3 1 python tools/probe_de.py --pop 262144 --dim 4096 --objective rosenbrock --strategy best --x0 4.096 --blocks 1 --gens 5
# accelerated PSO, config-3 shard shape; vanilla PSO fp32 d=64 (config-5 point)
run pso_accel_config3 "pso_move_kernel" 4 1 python tests/tools/quick_time_pso.py 2097152 256 3 3 1 1
run pso_vanilla_f32_d64 "pso_move_kernel" 4 1 python tests/tools/quick_time_pso.py 4194304 64 3 0 0 0
# gpurun brings back at most 64 MiB: export the pages here, keep only the headline kernel's report
for f in $O/*.ncu-rep; do
  n=${f%.ncu-rep}
  ncu -i $f --page raw --csv > $n.raw.csv 2>/dev/null
  ncu -i $f --page source --csv > $n.source.csv 2>/dev/null
  ncu -i $f --page details > $n.details.txt 2>/dev/null
  case $f in *k2_config2*) ;; *) rm -f $f ;; esac
done
gzip -9 $O/*.source.csv
ls -la $O; du -sh gpurun_out
