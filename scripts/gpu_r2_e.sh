# round-2 run E (1 GPU): whole GPU suite, default bench (dedup second-pass skip, .L2::64B, key prefetch), A/B of the 96-register
# hash variant, the other BASELINE workloads at full size (k-mer mode, three-level hierarchy, GTDB shape), reference arm,
# launch list, CLI bench (plain / gzip / BGZF / bzip2), compute-sanitizer over the new kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; rc=$?; tail -5 gpurun_out/r2e_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
timeout 1500 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err; python scripts/show_bench.py gpurun_out/r2e_bench.json
B="timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2e_$tag.json 2> gpurun_out/r2e_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2e_$tag.json; }
run hashregs5 TXR_HASH_REGS=5
run s64_off TXR_L2_SECTOR64=0
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; cut -c1-300 gpurun_out/r2e_bench_ref.json
S="python bench.py --reads 524288 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"kernel" -c 400 --csv --log-file gpurun_out/r2e_launches.csv $S > gpurun_out/r2e_ncu_bench.json 2> gpurun_out/r2e_ncu.err
python scripts/launch_summary.py gpurun_out/r2e_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ixf_query_small -s 3 -c 3 -o gpurun_out/r2e_prof_query $S > /dev/null 2>> gpurun_out/r2e_ncu.err
timeout 1200 python scripts/cli_bench.py --gz --bgzf --bz2 > gpurun_out/r2e_cli_bench.json 2> gpurun_out/r2e_cli_bench.err; cat gpurun_out/r2e_cli_bench.json; tail -5 gpurun_out/r2e_cli_bench.err
bash scripts/gpu_sanitize.sh
rm -rf /dev/shm/taxor_b200_bench/g1000_* /dev/shm/taxor_b200_cli
timeout 900 python bench.py --workload kmer --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2e_bench_kmer.json 2> gpurun_out/r2e_bench_kmer.err; tail -2 gpurun_out/r2e_bench_kmer.err; python scripts/show_bench.py gpurun_out/r2e_bench_kmer.json
rm -rf /dev/shm/taxor_b200_bench/*
timeout 1500 python bench.py --workload deep --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2e_bench_deep.json 2> gpurun_out/r2e_bench_deep.err; tail -2 gpurun_out/r2e_bench_deep.err; python scripts/show_bench.py gpurun_out/r2e_bench_deep.json
rm -rf /dev/shm/taxor_b200_bench/*
timeout 1500 python bench.py --workload gtdb --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2e_bench_gtdb.json 2> gpurun_out/r2e_bench_gtdb.err; tail -2 gpurun_out/r2e_bench_gtdb.err; python scripts/show_bench.py gpurun_out/r2e_bench_gtdb.json
ls -la gpurun_out/*.ncu-rep | tail -3
