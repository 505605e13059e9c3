#!/bin/bash
tag=${1:-r02_j}; shift
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 5 --warmup 3 --workload c4_uniref50_protein_k7_s10 --no-cpu-baseline > gpurun_out/${tag}_bench_c4_n$n.json 2> gpurun_out/${tag}_bench_c4_n$n.err
  tail -c 400 gpurun_out/${tag}_bench_c4_n$n.err | grep -v OMP_NUM | grep -v "^\*" | tail -5
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_c4_n$n.json').read().strip().splitlines()[-1])
    print('C4 N=$n ms_step %.3f value %.1f G/s e2e %.3f ms tuples %d stages %s path %s' % (d['ms_per_step'], d['value']/1e9, d['e2e']['ms_per_step'], d['config']['tuples'], d['roofline']['stages_ms'], d['config']['build_path']))
except Exception as e:
    print('C4 N=$n ERR', e)
PY
done
