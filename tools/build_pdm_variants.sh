#!/bin/bash
# Builds tuning variants of the library into gpurun_variants/ (they travel to the GPU box; the directory is not tracked).
# usage: tools/build_pdm_variants.sh name1="-DFLAG=.. -DFLAG2=.." name2="..."      then: gpurun -- 'bash tools/exp_pdm_variants.sh'
# default set: the level-1 -> level-2 feed variants named in DESIGN.md section 9
set -e
cd "$(dirname "$0")/../periodicity_b200/csrc"
mkdir -p ../../gpurun_variants
[ $# -eq 0 ] && set -- w256="-DPDM_PACK_FLUSH=256" fb4="-DPDM_FLUSH_BINS=4" fb10="-DPDM_FLUSH_BINS=10" fb20="-DPDM_FLUSH_BINS=20" plain="-DPDM_FEED_PLAIN=1" plainfb4="-DPDM_FEED_PLAIN=1 -DPDM_FLUSH_BINS=4" plainfb10="-DPDM_FEED_PLAIN=1 -DPDM_FLUSH_BINS=10" w256fb4="-DPDM_PACK_FLUSH=256 -DPDM_FLUSH_BINS=4"
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  make -j4 OUT=../../gpurun_variants/lib_$name.so OBJDIR=/tmp/pdc_build_$name EXTRA="$flags" > /tmp/build_$name.log 2>&1 &
done
wait
ls -la ../../gpurun_variants
