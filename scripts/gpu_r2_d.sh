# round-2 run D (2 GPUs): the multi-GPU paths on hardware -- CLI with two devices (byte-identical result file), index clone over
# NVLink, torchrun bench N=2 with the per-rank parity check and the per-rank e2e / H2D statistics, CLI start-up 1 vs 2 GPUs,
# launch list with the .L2::64B default (DRAM traffic of kernel #2)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus or clone or profile_head or cli_end_to_end" > gpurun_out/r2d_pytest_gpu.log 2>&1; rc=$?; tail -5 gpurun_out/r2d_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
timeout 1500 python bench.py --cpu-seconds 6 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -2 gpurun_out/r2d_bench_n1.err; python scripts/show_bench.py gpurun_out/r2d_bench_n1.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; tail -3 gpurun_out/r2d_bench_n2.err; python scripts/show_bench.py gpurun_out/r2d_bench_n2.json
S="python bench.py --reads 524288 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"kernel" -c 400 --csv --log-file gpurun_out/r2d_launches.csv $S > gpurun_out/r2d_ncu_bench.json 2> gpurun_out/r2d_ncu.err
python scripts/launch_summary.py gpurun_out/r2d_launches.csv
timeout 900 python scripts/cli_bench.py --gpus 1,2 --bgzf > gpurun_out/r2d_cli_bench.json 2> gpurun_out/r2d_cli_bench.err; cat gpurun_out/r2d_cli_bench.json; tail -5 gpurun_out/r2d_cli_bench.err
