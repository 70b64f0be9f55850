# round-2 run A (one gpurun call, 1 GPU): parity suite with the new kernels (fused distinct set, wide rows, staged upload /
# clone), configs[1] bench with A/B of the fused hash kernel, t_max sweep on a 1 GB index, launch list
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; rc=$?; tail -15 gpurun_out/r2a_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED: skipping the benches"; exit 1; fi
timeout 1500 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; python scripts/show_bench.py gpurun_out/r2a_bench.json
TXR_FUSE_DEDUP=0 timeout 900 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2a_bench_unfused.json 2> gpurun_out/r2a_bench_unfused.err; python scripts/show_bench.py gpurun_out/r2a_bench_unfused.json
S="python bench.py --reads 524288 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"kernel" -c 200 --csv --log-file gpurun_out/r2a_launches.csv $S > /dev/null 2> gpurun_out/r2a_ncu.err
python scripts/launch_summary.py gpurun_out/r2a_launches.csv
: > gpurun_out/r2a_tmax_sweep.jsonl
for T in 64 128 256 512 1024 2048 4096; do
  timeout 600 python bench.py --genome-len 4000000 --t-max $T --reads 200000 --steps 3 --warmup 3 --cpu-seconds 4 >> gpurun_out/r2a_tmax_sweep.jsonl 2> gpurun_out/r2a_tmax_$T.err
  tail -1 gpurun_out/r2a_tmax_sweep.jsonl | python scripts/show_bench.py /dev/stdin
done
