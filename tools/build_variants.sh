#!/bin/bash
# Builds tuning variants of libnls_b200.so into tools/ab/ (git-ignored; listed in .gpurunignore — take it out of there
# while variants have to travel to the GPU box): usage
#   tools/build_variants.sh "name1:-DFLAG=1 -DOTHER=2" "name2:..."
# Only de_f64.cu is recompiled per variant; the other objects come from the regular build.  Select a variant at run time
# with NLS_B200_LIB=tools/ab/libnls_b200_<name>.so.
set -e
cd "$(dirname "$0")/../nlsolver_b200/csrc"
make -s -j8
OUT=../../tools/ab
mkdir -p $OUT
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v \
      $flags -c de_f64.cu -o $OUT/de_f64_$name.o 2> $OUT/$name.ptxas.log
    nvcc -shared -o $OUT/libnls_b200_$name.so $(ls build/*.o | grep -v "build/de_f64.o") $OUT/de_f64_$name.o -lcudart -ldl 2>/dev/null
    echo "$name: $(grep -A2 "${REPORT:-de_generation_bulk_kernelIdLi1}" $OUT/$name.ptxas.log | grep -E 'registers|spill' | tr '\n' ' ')"
  ) &
done
wait
rm -f $OUT/*.o
