mkdir -p gpurun_out
for mode in 0 1 2 3 4 5; do
  BXB200_NVCC_FLAGS="-DRANK_LD_MODE=$mode" python -m bx_python_b200.build > /dev/null 2>&1
  echo "== mode $mode"
  python scratch/ab_count.py 2>&1 | tail -1
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:k_count_ranges_multi -s 2 -c 1 python scratch/ab_count.py 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate" | tr -s ' ' | tr '\n' ' '
  echo
done
