# round-2 run F (1 GPU): probe kernel with ONE step in flight per warp and 40 / 32 registers (48 / 64 warps per SM), alone and
# with the hash + dedup kernels of the next batch beside it (TXR_OVERLAP=1), at configs[1]
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "search_parity or early_exit or overlap" > gpurun_out/r2f_pytest_gpu.log 2>&1; rc=$?; tail -3 gpurun_out/r2f_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
B="timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2f_$tag.json 2> gpurun_out/r2f_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2f_$tag.json; }
run base TXR_X=0
run u1q8 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=8
run u1q10 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=10
run u1q12 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=12
run u1q16 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=16
run u1q6 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=6
run ov_u1q8h1 TXR_OVERLAP=1 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=8 TXR_HASH_CTAS_PER_SM=1 TXR_DEDUP_CTAS_PER_SM=1
run ov_u1q6h2 TXR_OVERLAP=1 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=6 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=1
run ov_u1q7h2 TXR_OVERLAP=1 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=7 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=1
run ov_u1q8h1r5 TXR_OVERLAP=1 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=8 TXR_HASH_CTAS_PER_SM=2 TXR_HASH_REGS=5 TXR_DEDUP_CTAS_PER_SM=1
run ov_u1q10h1 TXR_OVERLAP=1 TXR_QUERY_UNROLL=1 TXR_QUERY_CTAS_PER_SM=10 TXR_HASH_CTAS_PER_SM=1 TXR_DEDUP_CTAS_PER_SM=1
