#!/bin/bash
# round 2, final evidence run on one GPU: full GPU suite, smoke, bench line, ncu launch list, ncu --set full captures
# (exported to CSV on the box: gpurun_out/ only carries 64 MiB back)
tag=${1:-r02_s}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $o/${tag}_pytest_gpu.log
python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; echo "smoke rc $?"; tail -4 $o/${tag}_smoke.log
python bench.py > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err; echo "bench rc $?"; tail -c 300 $o/${tag}_bench_c2.err
python bench.py --impl reference --steps 1 --warmup 0 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $o/${tag}_launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_launch.log 2>&1
full() {  # name, kernel regex, launch-skip, count, bench args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip -c $cnt -o $o/${tag}_$name -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra "$@" > $o/${tag}_ncu_$name.log 2>&1
  ncu -i $o/${tag}_$name.ncu-rep --page raw --csv > $o/${tag}_ncu_full_raw_$name.csv 2>/dev/null
  ncu -i $o/${tag}_$name.ncu-rep --page source --csv > $o/${tag}_ncu_source_$name.csv 2>/dev/null
  ls -la $o/${tag}_$name.ncu-rep; rm -f $o/${tag}_$name.ncu-rep
}
full c2_build 'sketch_dense_kernel|dense_partition_kernel|dense_bucket_kernel' 8 4
full c3_search 'query_kernel|query_scan_kernel|finalize_pairs_kernel|expand_hits_kernel' 8 8
full target_build 'sketch_quad_kernel|pair_partition_kernel|bucket_sort_bin_kernel' 6 3 --workload target_100m_dayhoff_k16_s1
du -sh $o
