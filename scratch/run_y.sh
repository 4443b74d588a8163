# run "y": compute-sanitizer over the parity suites (memcheck on everything but the full-size cases, racecheck on the
# shared-memory kernels: staged fill, single-pass look-back, summarize, small-batch find)
mkdir -p gpurun_out
K='not c2_full and not large_properties and not chromosome_scale and not genome_batch and not int32_offsets and not summarize_large'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scores.py tests/test_gpu_abi_client.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_memcheck_y.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_y.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scores.py -m gpu -q -x -k "find_random or single_pass or small_path or edge_sets or summarize_golden or forest_vs or set_ranges_many" > gpurun_out/sanitizer_racecheck_y.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_y.log
