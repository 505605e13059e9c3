#!/bin/bash
# round 2, session r: bench line of the current defaults; rank-sort variant of the dense bucket kernel once more
mkdir -p gpurun_out
tools/ab.sh c2_swissprot_hp_k24_s1 20 1.0 default ranksort > gpurun_out/r02r_ab.log 2>&1
cat gpurun_out/r02r_ab.log
python bench.py > gpurun_out/r02r_bench_c2.json 2> gpurun_out/r02r_bench_c2.err; echo "bench rc $?"; tail -c 600 gpurun_out/r02r_bench_c2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02r_bench_c2.json').read().strip().splitlines()[-1])
print('ms_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'whole', d['roofline']['whole_step'])
print('stages', d['roofline']['stages']['sketch']['ms'], d['roofline']['stages']['index_build']['ms_by_kernel'])
print('search', {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if 'ms' in kk}) for k, v in d['search'].items() if k in ('pairs', 'pairs_and_hits')})
print('extra', json.dumps(d.get('extra'))[:1200])
print('clocks', d.get('clocks'))
P
