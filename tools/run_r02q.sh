#!/bin/bash
# round 2, session q: earlier head loads in the bucket kernels (default), CTAs per SM of the partition kernels (variants)
mkdir -p gpurun_out
{
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unstable or dense or csr_and_search or bucket_sort" 2>&1 | tail -3
tools/ab.sh c2_swissprot_hp_k24_s1 20 1.0 default dp5 dp6
tools/ab.sh target_100m_dayhoff_k16_s1 20 1.0 default pp5
tools/ab.sh c2_swissprot_hp_k24_s1 20 0.125 default dp5 dp6
tools/ab.sh target_100m_dayhoff_k16_s1 20 0.125 default pp5
} > gpurun_out/r02q_ab.log 2>&1
cat gpurun_out/r02q_ab.log
