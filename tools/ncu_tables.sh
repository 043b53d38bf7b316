#!/bin/bash
# tuning aid: DRAM traffic / L2 hit rate of the search kernel vs. number of resident warps (= total visited-table footprint)
for c in "$@"; do
  DAB_V2_CTAS_PER_SM=$c timeout 200 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:search_kernel -c 1 --csv python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline 2>/dev/null | grep -E "dram__|lts__|gpu__time|grid_size" | awk -F'","' -v c=$c '{print "ctas/sm=" c, $(NF-2), $(NF-1), $NF}'
done
