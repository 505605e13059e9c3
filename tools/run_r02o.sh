#!/bin/bash
# round 2, session o: tie fix-up after the rounds (bin kernel), 6 CTAs per SM for the sketch kernels, new prep kernels; search wall
mkdir -p gpurun_out
{
for v in default tiefix; do
  if [ "$v" == "default" ]; then unset KS_LIB_PATH; else export KS_LIB_PATH=$PWD/kmerseek_b200/variants/lib_$v.so; fi
  echo "== parity $v"; python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
done
unset KS_LIB_PATH
tools/ab.sh c2_swissprot_hp_k24_s1 20 1.0 default ctas6
tools/ab.sh target_100m_dayhoff_k16_s1 20 1.0 default ctas6 tiefix
tools/ab.sh c3_search_dayhoff_k16_s1 20 1.0 default tiefix
tools/ab.sh target_100m_dayhoff_k16_s1 20 0.125 default ctas6 tiefix
tools/ab.sh c2_swissprot_hp_k24_s1 20 0.125 default ctas6
python tools/search_probe.py --workload c3_search_dayhoff_k16_s1 --reps 6 2>&1 | tail -6
} > gpurun_out/r02o_ab.log 2>&1
cat gpurun_out/r02o_ab.log
