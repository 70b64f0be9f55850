# round-2 run F3 (1 GPU): the automatic overlap (first batch of a call hashes on the whole GPU, last batch probes alone), against the
# serial schedule and a few shapes
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f3_pytest_gpu.log 2>&1; rc=$?; tail -3 gpurun_out/r2f3_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
B="timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2f3_$tag.json 2> gpurun_out/r2f3_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2f3_$tag.json; }
run auto TXR_X=0
run serial TXR_OVERLAP=0
run auto_d4 TXR_DEDUP_CTAS_PER_SM=4
run auto_d6 TXR_DEDUP_CTAS_PER_SM=6
run auto_r32q8 TXR_QUERY_REGS=32 TXR_QUERY_CTAS_PER_SM=8
run auto_r32q7 TXR_QUERY_REGS=32 TXR_QUERY_CTAS_PER_SM=7
run auto_fused TXR_FUSE_DEDUP=1
run auto_q7 TXR_QUERY_CTAS_PER_SM=7
run auto_s3 TAXOR_BENCH_E2E_SLOTS=3
run auto_s5 TAXOR_BENCH_E2E_SLOTS=5 TAXOR_BENCH_RESIDENT_SLOTS=3
run auto_h3r5 TXR_HASH_REGS=5 TXR_HASH_CTAS_PER_SM=3
