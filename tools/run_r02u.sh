#!/bin/bash
# round 2, session u: directory entries from the bin offsets, oversize flag read at the head of the bin kernel
mkdir -p gpurun_out
{
python -m pytest tests/test_gpu_parity.py tests/test_gpu_c5_sweep.py -m gpu -x -q -k "unstable or csr_and_search or bucket_sort or medium_scale or c5 or tile_boundary or empty or golden_manysearch or batches" 2>&1 | tail -3
for wl in target_100m_dayhoff_k16_s1 c3_search_dayhoff_k16_s1 c2_swissprot_hp_k24_s1; do python tools/quick_build_bench.py $wl 20 1.0; done
python tools/quick_build_bench.py target_100m_dayhoff_k16_s1 20 0.125
} > gpurun_out/r02u_ab.log 2>&1
cat gpurun_out/r02u_ab.log
