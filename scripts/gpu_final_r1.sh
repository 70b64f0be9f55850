# round-1 evidence run (one gpurun call, 1 GPU): parity suite, sanitizer, bench (both arms), launch list, full ncu capture of
# kernel #2, CLI bench, and the other BASELINE workloads at full size
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; python scripts/show_bench.py gpurun_out/bench_r1_final.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2> gpurun_out/bench_r1_final_ref.err; cut -c1-300 gpurun_out/bench_r1_final_ref.json
S="python bench.py --reads 524288 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"kernel" -c 200 --csv --log-file gpurun_out/launches_final.csv $S > gpurun_out/ncu_final_bench.json 2> gpurun_out/ncu_final.err
python scripts/launch_summary.py gpurun_out/launches_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ixf_query_small -s 3 -c 3 -o gpurun_out/prof_final_query $S > /dev/null 2>> gpurun_out/ncu_final.err
ls -la gpurun_out/*.ncu-rep
timeout 900 python scripts/cli_bench.py --gz > gpurun_out/cli_bench_final.json 2> gpurun_out/cli_bench.err; cat gpurun_out/cli_bench_final.json; tail -3 gpurun_out/cli_bench.err
bash scripts/gpu_sanitize.sh
timeout 900 python bench.py --workload kmer --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_r1_kmer.json 2> gpurun_out/bench_r1_kmer.err; tail -2 gpurun_out/bench_r1_kmer.err; python scripts/show_bench.py gpurun_out/bench_r1_kmer.json
rm -rf /dev/shm/taxor_b200_cli
timeout 1500 python bench.py --workload deep --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_r1_deep.json 2> gpurun_out/bench_r1_deep.err; tail -2 gpurun_out/bench_r1_deep.err; python scripts/show_bench.py gpurun_out/bench_r1_deep.json
