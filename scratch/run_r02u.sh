# round 2: compute-sanitizer memcheck over the kernels added after r02s (bucketed set kernel, overlapped find pipeline, speculative fill)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "bucketed or single_pass_and_three_pass or set_ranges_many or aggregate_genome or aggregate_random" > gpurun_out/r02u_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02u_memcheck.log | sort | uniq -c
