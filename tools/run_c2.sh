python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense or csr_and_search" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err; tail -c 300 gpurun_out/tmp_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/tmp_bench.json').read().strip().splitlines()[-1])
print('ms_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['stages']['sketch']['ms'], d['roofline']['stages']['index_build']['ms_by_kernel'])
"
