#!/bin/bash
# one-GPU round-2 evidence run: GPU tests, default bench, ncu launch list, ncu --set full of the dominant kernels
tag=${1:-r02_k}
o=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err; tail -c 400 $o/${tag}_bench_c2.err
python - <<PY
import json
d=json.loads([l for l in open('$o/${tag}_bench_c2.json') if l.startswith('{')][-1])
print('ms_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['stages']['sketch']['ms'], d['roofline']['stages']['index_build']['ms_by_kernel'], 'launches', d['gpu_launches'])
t=d['extra']['target_100m_dayhoff_k16_s1']; print('target', t['ms_per_step'], t['ms_per_step_e2e']); print('ingest', d['extra']['ingest'])
s=d['search']; print('search', s['pairs']['ms_per_batch_wall'], s['pairs']['ms_per_batch_kernels'], s['pairs_and_hits']['ms_per_batch_wall'])
print('cpu', d['cpu_baseline'])
PY
# launch list of the same command (short run; times under ncu are cold-cache and serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $o/${tag}_launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_launch.log 2>&1
# full captures: the C2 build kernels (skip the warm-up builds), the target-run kernels, the search kernels
ncu --set full --clock-control none --import-source on -k regex:'sketch_dense_kernel|dense_partition_kernel|dense_bucket_kernel' \
  --launch-skip 6 -c 4 -o $o/${tag}_c2_build -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'query_kernel|query_scan_kernel|finalize_pairs_kernel|expand_hits_kernel' \
  --launch-skip 8 -c 8 -o $o/${tag}_c3_search -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sketch_quad_kernel|pair_partition_kernel|bucket_sort_bin_kernel|bucket_sort_rep_kernel' \
  --launch-skip 6 -c 4 -o $o/${tag}_target_build -f python bench.py --workload target_100m_dayhoff_k16_s1 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_full3.log 2>&1
ls -la $o/${tag}_*
