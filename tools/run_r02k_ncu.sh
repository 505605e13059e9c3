#!/bin/bash
# ncu evidence of one bench command: launch list + --set full of the dominant kernels; reports are exported to CSV on the
# box (raw + source pages) and removed -- gpurun_out/ only carries 64 MiB back
tag=${1:-r02_k}
o=gpurun_out
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
du -sh $o; ls -la $o
