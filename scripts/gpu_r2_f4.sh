# round-2 run F4 (1 GPU): choosing the overlap shapes for the END-TO-END number (probe CTAs x dedup CTAs x batch ramp), 5 steps each
set -x
mkdir -p gpurun_out
B="timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2f4_$tag.json 2> gpurun_out/r2f4_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2f4_$tag.json; }
run q6d3 TXR_X=0
run q7d3 TXR_QUERY_CTAS_PER_SM=7
run q7d4 TXR_QUERY_CTAS_PER_SM=7 TXR_DEDUP_CTAS_PER_SM=4
run q7d6 TXR_QUERY_CTAS_PER_SM=7 TXR_DEDUP_CTAS_PER_SM=6
run q6d6 TXR_DEDUP_CTAS_PER_SM=6
run q7d6_r15 TXR_QUERY_CTAS_PER_SM=7 TXR_DEDUP_CTAS_PER_SM=6 TXR_RAMP=1.5
run q7d6_r13 TXR_QUERY_CTAS_PER_SM=7 TXR_DEDUP_CTAS_PER_SM=6 TXR_RAMP=1.3
run q7d3_r15 TXR_QUERY_CTAS_PER_SM=7 TXR_RAMP=1.5
run q7d6_h3 TXR_QUERY_CTAS_PER_SM=7 TXR_DEDUP_CTAS_PER_SM=6 TXR_HASH_CTAS_PER_SM=3
run serial TXR_OVERLAP=0
