#!/bin/bash
# round 2, session p: defaults = direct scatter + tie fix-up + chunk table in shared memory + speculative hash rows + coalesced scans
mkdir -p gpurun_out
{
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unstable or dense or csr_and_search or bucket_sort or search_query or many_targets or tile_boundary or empty or pipelined or golden_manysearch" 2>&1 | tail -3
for wl in c2_swissprot_hp_k24_s1 target_100m_dayhoff_k16_s1 c3_search_dayhoff_k16_s1; do
  for f in 1.0 0.125; do python tools/quick_build_bench.py $wl 20 $f; done
done
python tools/search_probe.py --workload c3_search_dayhoff_k16_s1 --reps 6 2>&1 | tail -4
} > gpurun_out/r02p_ab.log 2>&1
cat gpurun_out/r02p_ab.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02p_launches_search_probe.csv python tools/search_probe.py --workload c3_search_dayhoff_k16_s1 --reps 3 > gpurun_out/r02p_ncu_search.log 2>&1
python profiles/launch_summary.py gpurun_out/r02p_launches_search_probe.csv | head -12
