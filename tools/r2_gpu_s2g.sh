#!/bin/bash
# session 2, call G: ncu of the mid-size and fp32 short-row kernels
set -x
N=gpurun_out/r2_s2g
mkdir -p $N
NCU="ncu --set full --clock-control none --import-source on"
run() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $N/$name.plain.log 2>&1 && $NCU -k regex:"$rx" -s $skip -c $cnt -o $N/$name "$@" > $N/$name.ncu.log 2>&1
  echo "== $name rc=$?"
}
run k2_p65536_d64_f64 "de_generation_kernel" 3 1 python tools/probe_de.py --pop 65536 --dim 64 --objective sphere --blocks 1 --gens 6
run k2_p2e20_d64_f32 "de_generation_kernel" 3 1 python tools/probe_de.py --pop 1048576 --dim 64 --objective sphere --dtype f32 --blocks 1 --gens 6
run k2_p2e22_d64_f32 "de_generation_kernel" 3 1 python tools/probe_de.py --pop 4194304 --dim 64 --objective sphere --dtype f32 --blocks 1 --gens 6
run pso_vanilla_f32_p2e20 "pso_move_kernel" 4 1 python tests/tools/quick_time_pso.py 1048576 64 3 0 0 0
for f in $N/*.ncu-rep; do
  n=${f%.ncu-rep}
  ncu -i $f --page raw --csv > $n.raw.csv 2>/dev/null
  ncu -i $f --page source --csv > $n.source.csv 2>/dev/null
  rm -f $f
done
gzip -9f $N/*.source.csv
