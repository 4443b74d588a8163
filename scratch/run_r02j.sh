for g in 32 64 128; do
  echo "== L2 fetch $g"
  BXB200_DEBUG=1 BXB200_L2_FETCH=$g python scratch/ab_count.py 2>&1 | grep -E "bxb200|count_ranges"
  BXB200_L2_FETCH=$g timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,dram__sectors_read.sum -k regex:k_count_ranges_multi -s 2 -c 1 python scratch/ab_count.py 2>&1 | grep -E "dram__|gpu__time|hit_rate|lts__t_sectors" | tr -s ' ' | tr '\n' ' '
  echo
done
