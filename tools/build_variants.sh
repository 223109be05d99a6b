#!/bin/bash
# Dev aid: build tune_libs/<name>.so for the -D variants of gls.cu listed in a spec file
# (one "name<TAB>flags" per line); the other objects are compiled once.  usage: tools/build_variants.sh specs.txt [jobs]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/periodicity_b200/csrc
OBJ=/tmp/var
JOBS=${2:-8}
mkdir -p $OBJ $ROOT/tune_libs
FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
cd $SRC
for f in capi glsm pdm peaks; do
  if [ ! -f $OBJ/$f.o ] || [ $f.cu -nt $OBJ/$f.o ] || [ pdc_common.cuh -nt $OBJ/$f.o ] || [ gls_common.cuh -nt $OBJ/$f.o ]; then
    nvcc $FL -c $f.cu -o $OBJ/$f.o &
  fi
done
wait
n=0
while IFS=$'\t' read -r name defs; do
  ( eval nvcc $FL $defs -c gls.cu -o $OBJ/gls_$name.o && \
    nvcc $FL --cudart static -shared -o $ROOT/tune_libs/$name.so $OBJ/gls_$name.o $OBJ/capi.o $OBJ/glsm.o $OBJ/pdm.o $OBJ/peaks.o ) &
  n=$((n+1))
  if [ $((n % JOBS)) -eq 0 ]; then wait; fi
done < "$1"
wait
ls $ROOT/tune_libs | wc -l
du -sh $ROOT/tune_libs
