#!/bin/bash
# round 2, final tree on 8 GPUs: sharded-search parity test and the strong-scaling bench line at N = 8
tag=${1:-r02_x}
o=gpurun_out
mkdir -p $o
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tee $o/${tag}_multi_8gpu_pytest.log | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_c2_strong_n8.json 2> $o/${tag}_bench_n8.err
tail -c 300 $o/${tag}_bench_n8.err
python - <<PY
import json
d=json.loads(open('$o/${tag}_bench_c2_strong_n8.json').read().strip().splitlines()[-1])
s=d['search']
print('N=8 ms_step %.3f e2e %.3f | target %.3f e2e %.3f | search wall %.3f kernels %.3f pairs %d verify %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['extra']['target_100m_dayhoff_k16_s1']['ms_per_step'], d['extra']['target_100m_dayhoff_k16_s1']['ms_per_step_e2e'], s['pairs']['ms_per_batch_wall'], s['pairs']['ms_per_batch_kernels'], s['pairs']['pairs'], s.get('sharded_equals_single_gpu')))
PY
