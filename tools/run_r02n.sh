#!/bin/bash
# round 2, session n: direct scatter (no staging in window order) in both sketch kernels, 5-bit -> bit streams, fast query scan
mkdir -p gpurun_out
{
for v in default bits5; do
  if [ "$v" == "default" ]; then unset KS_LIB_PATH; else export KS_LIB_PATH=$PWD/kmerseek_b200/variants/lib_$v.so; fi
  echo "== parity $v"; python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
done
unset KS_LIB_PATH
tools/ab.sh c2_swissprot_hp_k24_s1 20 1.0 default nodirect bits5
tools/ab.sh target_100m_dayhoff_k16_s1 20 1.0 default nodirect
tools/ab.sh c2_swissprot_hp_k24_s1 20 0.125 default nodirect bits5
tools/ab.sh target_100m_dayhoff_k16_s1 20 0.125 default nodirect
tools/ab.sh c4_slice_protein_k7_s10 5 1.0 default nodirect
python tools/search_probe.py 2>&1 | tail -8
} > gpurun_out/r02n_ab.log 2>&1
cat gpurun_out/r02n_ab.log
