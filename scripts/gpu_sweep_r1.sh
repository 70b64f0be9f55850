# one gpurun call: A/B runs of the bench at configs[1] (index built once, cached in /dev/shm) + the k-mer workload with the parity check
set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 900 $B > gpurun_out/lv_$tag.json 2>> gpurun_out/lv.err
  echo "== $tag ($*)"; python scripts/show_bench.py gpurun_out/lv_$tag.json
}
run base TXR_RAMP_TAIL=0
run tail TXR_RAMP_TAIL=1
run lv4 TXR_LEVEL_CTAS_PER_SM=4
run lv5 TXR_LEVEL_CTAS_PER_SM=5
run lv6 TXR_LEVEL_CTAS_PER_SM=6
run lv3 TXR_LEVEL_CTAS_PER_SM=3
tail -3 gpurun_out/lv.err
timeout 900 python bench.py --workload kmer --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_r1_kmer.json 2> gpurun_out/bench_r1_kmer.err; tail -2 gpurun_out/bench_r1_kmer.err; python scripts/show_bench.py gpurun_out/bench_r1_kmer.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_kmer.json').read().strip().splitlines()[-1]); print(d['parity_at_scale']); print(d['index_build'])"
