#!/bin/bash
# round 2, final tree on one GPU: full GPU suite, smoke, bench line
tag=${1:-r02_w}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $o/${tag}_pytest_gpu.log
python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; echo "smoke rc $?"; tail -4 $o/${tag}_smoke.log
python bench.py > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err; echo "bench rc $?"; tail -c 300 $o/${tag}_bench_c2.err
python - <<P
import json
d=json.loads(open('$o/${tag}_bench_c2.json').read().strip().splitlines()[-1])
print('ms_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'whole', d['roofline']['whole_step']['frac'])
print('stages', d['roofline']['stages']['sketch']['ms'], d['roofline']['stages']['index_build']['ms_by_kernel'])
for k in ('pairs','pairs_and_hits'):
    v=d['search'][k]; print(k, v['ms_per_batch_wall'], v['ms_per_batch_kernels'])
t=d['extra']['target_100m_dayhoff_k16_s1']; print('target', t['ms_per_step'], t['ms_per_step_e2e'], t['whole_step_frac']); print(d['extra']['ingest'])
P
