#!/bin/bash
# usage: tools/build_variants.sh file.cu name1 "-Dflags1" name2 "-Dflags2" ...  -> kmerseek_b200/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
SRC=$1; shift
mkdir -p kmerseek_b200/variants
python -m kmerseek_b200.build > /dev/null
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc $flags -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -c kmerseek_b200/csrc/$SRC -o /tmp/var_$name.o
  objs=""
  for f in api.cu sketch.cu index_build.cu dense.cu search.cu; do
    if [ "$f" == "$SRC" ]; then objs="$objs /tmp/var_$name.o"; else objs="$objs kmerseek_b200/build/$f.o"; fi
  done
  /usr/local/cuda/bin/nvcc -shared -o kmerseek_b200/variants/lib_$name.so $objs -lz -ldl -cudart static
  echo built $name
done
