mkdir -p gpurun_out
python scratch/ab_aggregate.py | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_aggregate_multi -s 2 -c 1 -o gpurun_out/r02r_prof_aggregate -f python scratch/ab_aggregate.py > gpurun_out/r02r_ncu.log 2>&1; tail -1 gpurun_out/r02r_ncu.log
