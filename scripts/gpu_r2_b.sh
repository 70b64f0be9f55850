# round-2 run B (1 GPU): fused hash kernel v2 (groups of 128 keys), probe-kernel shape sweeps (CTAs per SM x steps in flight),
# hash-beside-probe overlap with those shapes, the GTDB-shaped workload, ncu captures (traffic of kernel #2, wide kernel)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or syncmer or wide or overlap or scheme or clone or cli" > gpurun_out/r2b_pytest_gpu.log 2>&1; rc=$?; tail -5 gpurun_out/r2b_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
B="timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2"
run() { # tag, env...
  tag=$1; shift
  env "$@" $B > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2b_$tag.json
}
timeout 1500 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; python scripts/show_bench.py gpurun_out/r2b_bench.json
run q8u2 TXR_X=0
run q6u2 TXR_QUERY_CTAS_PER_SM=6
run q5u2 TXR_QUERY_CTAS_PER_SM=5
run q4u2 TXR_QUERY_CTAS_PER_SM=4
run q5u3 TXR_QUERY_CTAS_PER_SM=5 TXR_QUERY_UNROLL=3
run q4u3 TXR_QUERY_CTAS_PER_SM=4 TXR_QUERY_UNROLL=3
run q4u4 TXR_QUERY_CTAS_PER_SM=4 TXR_QUERY_UNROLL=4
run q3u4 TXR_QUERY_CTAS_PER_SM=3 TXR_QUERY_UNROLL=4
run ov_q5u2h1 TXR_OVERLAP=1 TXR_QUERY_CTAS_PER_SM=5 TXR_HASH_CTAS_PER_SM=1
run ov_q4u3h1 TXR_OVERLAP=1 TXR_QUERY_CTAS_PER_SM=4 TXR_QUERY_UNROLL=3 TXR_HASH_CTAS_PER_SM=1
run ov_q3u4h1 TXR_OVERLAP=1 TXR_QUERY_CTAS_PER_SM=3 TXR_QUERY_UNROLL=4 TXR_HASH_CTAS_PER_SM=1
run ov_q4u2h1 TXR_OVERLAP=1 TXR_QUERY_CTAS_PER_SM=4 TXR_HASH_CTAS_PER_SM=1
run ov_q3u3h2 TXR_OVERLAP=1 TXR_QUERY_CTAS_PER_SM=3 TXR_QUERY_UNROLL=3 TXR_HASH_CTAS_PER_SM=2
run ov_q3u4h1_s3 TXR_OVERLAP=1 TXR_QUERY_CTAS_PER_SM=3 TXR_QUERY_UNROLL=4 TXR_HASH_CTAS_PER_SM=1 TAXOR_BENCH_RESIDENT_SLOTS=3
S="python bench.py --reads 262144 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ixf_query_small -s 2 -c 2 -o gpurun_out/r2b_prof_query $S > /dev/null 2> gpurun_out/r2b_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syncmer_kernel -s 1 -c 1 -o gpurun_out/r2b_prof_hash $S > /dev/null 2>> gpurun_out/r2b_ncu.err
rm -rf /dev/shm/taxor_b200_bench/g1000_*
timeout 1800 python bench.py --workload gtdb --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2b_bench_gtdb.json 2> gpurun_out/r2b_bench_gtdb.err; tail -3 gpurun_out/r2b_bench_gtdb.err; python scripts/show_bench.py gpurun_out/r2b_bench_gtdb.json
G="python bench.py --workload gtdb --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ixf_query_large -s 1 -c 1 -o gpurun_out/r2b_prof_large $G > /dev/null 2>> gpurun_out/r2b_ncu.err
ls -la gpurun_out/*.ncu-rep
