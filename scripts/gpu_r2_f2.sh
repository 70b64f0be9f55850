# round-2 run F2 (1 GPU): hash + dedup of the next batch beside the 40-register probe kernel, dedup grid sized to fit
set -x
mkdir -p gpurun_out
B="timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2f2_$tag.json 2> gpurun_out/r2f2_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2f2_$tag.json; }
O="TXR_OVERLAP=1 TXR_QUERY_UNROLL=1"
run base TXR_X=0
run q6h2d4 $O TXR_QUERY_CTAS_PER_SM=6 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=4
run q6h2d6 $O TXR_QUERY_CTAS_PER_SM=6 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=6
run q6h2d3 $O TXR_QUERY_CTAS_PER_SM=6 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=3
run q7h2d4 $O TXR_QUERY_CTAS_PER_SM=7 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=4
run q5h2d4 $O TXR_QUERY_CTAS_PER_SM=5 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=4
run q8h1d4 $O TXR_QUERY_CTAS_PER_SM=8 TXR_HASH_CTAS_PER_SM=1 TXR_DEDUP_CTAS_PER_SM=4
run q6h2d4_s4 $O TXR_QUERY_CTAS_PER_SM=6 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=4 TAXOR_BENCH_E2E_SLOTS=4 TAXOR_BENCH_RESIDENT_SLOTS=3
run q6h2d4_l8 $O TXR_QUERY_CTAS_PER_SM=6 TXR_LEVEL_CTAS_PER_SM=8 TXR_HASH_CTAS_PER_SM=2 TXR_DEDUP_CTAS_PER_SM=4
