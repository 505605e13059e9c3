#!/bin/bash
# round 2, session m: GPU suite, fixed costs of small shards (1/8 of C2 and of the target run), L2-hint A/B of the dense sketch kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc $?" ; tail -3 gpurun_out/r02m_pytest.log
{
for wl in c2_swissprot_hp_k24_s1 target_100m_dayhoff_k16_s1; do
  for f in 1.0 0.125; do python tools/quick_build_bench.py $wl 20 $f; done
done
tools/ab.sh c2_swissprot_hp_k24_s1 20 1.0 default l2hint l2hint_cs cs_only
tools/ab.sh c2_swissprot_hp_k24_s1 20 1.0 default l2hint l2hint_cs cs_only
} > gpurun_out/r02m_ab.log 2>&1
cat gpurun_out/r02m_ab.log
for wl in c2_swissprot_hp_k24_s1 target_100m_dayhoff_k16_s1; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_${wl}_eighth.csv python tools/quick_build_bench.py $wl 5 0.125 > gpurun_out/r02m_ncu_$wl.log 2>&1
done
KS_TIMING=1 python tools/quick_build_bench.py target_100m_dayhoff_k16_s1 5 0.125 2>&1 | tail -12
