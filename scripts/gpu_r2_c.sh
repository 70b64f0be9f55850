# round-2 run C (1 GPU): whole GPU suite, configs[1] with the unfused default + key prefetch + 64-register probe variant, A/B of the
# 7-CTA variant, GTDB-shaped workload with its parity check, ncu capture of the wide kernel, CLI start-up timing (staged upload)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; rc=$?; tail -5 gpurun_out/r2c_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
timeout 1500 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err; python scripts/show_bench.py gpurun_out/r2c_bench.json
B="timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2c_$tag.json; }
run q7 TXR_QUERY_CTAS_PER_SM=7
run q6 TXR_QUERY_CTAS_PER_SM=6
run fused TXR_FUSE_DEDUP=1
run s64_all TXR_L2_SECTOR64=1
run s64_lv TXR_L2_SECTOR64=2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; cut -c1-600 gpurun_out/r2c_bench_ref.json
S="python bench.py --reads 524288 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"kernel" -c 400 --csv --log-file gpurun_out/r2c_launches.csv $S > gpurun_out/r2c_ncu_bench.json 2> gpurun_out/r2c_ncu.err
python scripts/launch_summary.py gpurun_out/r2c_launches.csv
TAXOR_TIMING=1 timeout 900 python scripts/cli_bench.py > gpurun_out/r2c_cli_bench.json 2> gpurun_out/r2c_cli_bench.err; cat gpurun_out/r2c_cli_bench.json; tail -5 gpurun_out/r2c_cli_bench.err
rm -rf /dev/shm/taxor_b200_bench/g1000_* /dev/shm/taxor_b200_cli
timeout 1800 python bench.py --workload gtdb --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2c_bench_gtdb.json 2> gpurun_out/r2c_bench_gtdb.err; tail -3 gpurun_out/r2c_bench_gtdb.err; python scripts/show_bench.py gpurun_out/r2c_bench_gtdb.json
G="python bench.py --workload gtdb --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ixf_query_large -c 4 -o gpurun_out/r2c_prof_large $G > /dev/null 2>> gpurun_out/r2c_ncu.err
ls -la gpurun_out/*.ncu-rep
