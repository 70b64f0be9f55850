# A/B of the hash/query overlap at configs[1]: CTAs per SM of the query, hash and dedup grids
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
export TAXOR_BENCH_RESIDENT_SLOTS=2
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
run() { # tag overlap query hash dedup
  TXR_OVERLAP=$2 TXR_QUERY_CTAS_PER_SM=$3 TXR_HASH_CTAS_PER_SM=$4 TXR_DEDUP_CTAS_PER_SM=$5 timeout 900 $B > gpurun_out/ov_$1.json 2>> gpurun_out/ov.err
  echo "== $1 (overlap=$2 query=$3 hash=$4 dedup=$5)"; python scripts/show_bench.py gpurun_out/ov_$1.json
}
run off 0 8 8 6
run q5h1d2 1 5 1 2
run q4h2d3 1 4 2 3
run q6h1d1 1 6 1 1
run q5h1d3 1 5 1 3
run q6h1d2 1 6 1 2
run q4h1d2 1 4 1 2
tail -3 gpurun_out/ov.err
