#!/bin/bash
# usage: tools/ab.sh workload steps frac variant...   (variant "default" = the in-tree library)
wl=$1; steps=$2; frac=$3; shift 3
for v in "$@"; do
  if [ "$v" == "default" ]; then python tools/quick_build_bench.py $wl $steps $frac
  else KS_LIB_PATH=$PWD/kmerseek_b200/variants/lib_$v.so python tools/quick_build_bench.py $wl $steps $frac; fi
done
