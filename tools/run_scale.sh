#!/bin/bash
# multi-GPU parity test + strong-scaling bench at N = 1, 2, 4, 8 on one box (outputs under gpurun_out/)
tag=${1:-r02_i}
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tee gpurun_out/${tag}_multi_8gpu_pytest.log | tail -8
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  fi
  tail -c 300 gpurun_out/${tag}_bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_n$n.json').read().strip().splitlines()[-1])
    s=d['search']
    print('N=$n ms_step %.3f value %.1f G/s e2e %.3f ms | target %.3f ms e2e %.3f | search wall %.3f kernels %.3f pairs %d verify %s' % (d['ms_per_step'], d['value']/1e9, d['e2e']['ms_per_step'], d['extra']['target_100m_dayhoff_k16_s1']['ms_per_step'], d['extra']['target_100m_dayhoff_k16_s1']['ms_per_step_e2e'], s['pairs']['ms_per_batch_wall'], s['pairs']['ms_per_batch_kernels'], s['pairs']['pairs'], s['sharded_equals_single_gpu']))
except Exception as e:
    print('N=$n ERR', e)
PY
done
