# e2e A/B: chunk size and find mode on the host path
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-bitset --no-cpu 2>/dev/null > gpurun_out/r02v_$name.json
  python -c "
import json
d=json.load(open('gpurun_out/r02v_$name.json'))
print('$name', 'e2e', '%.4g'%d['e2e']['value'], 'i64', '%.4g'%d['extra']['e2e_int64_offsets'], 'serial', '%.4g'%d['extra']['e2e_serial_copies'], 'cnt', '%.4g'%d['e2e']['count_only_value'], 'd2h', round(d['e2e']['d2h_ceiling_gbs_slowest_rank'],1))"
}
run default
run chunk512k BXB200_CHUNK_QUERIES=524288
run chunk2m BXB200_CHUNK_QUERIES=2097152
run mode0 BXB200_FIND_MODE=0
run mode0_512k BXB200_FIND_MODE=0 BXB200_CHUNK_QUERIES=524288
