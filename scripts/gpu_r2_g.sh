# round-2 run G (8 GPUs): torchrun N=8 bench with the parity check on every rank and the per-rank end-to-end / H2D statistics (the
# host-side ceiling of the end-to-end curve), the same with write-combined pinned read buffers, CLI start-up on 1 vs 8 GPUs
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r2g_topo.txt; free -g >> gpurun_out/r2g_topo.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 1500 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err; tail -3 gpurun_out/r2g_bench_n8.err; python scripts/show_bench.py gpurun_out/r2g_bench_n8.json
TXR_HOST_WC=1 timeout 900 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n8_wc.json 2> gpurun_out/r2g_bench_n8_wc.err; tail -3 gpurun_out/r2g_bench_n8_wc.err; python scripts/show_bench.py gpurun_out/r2g_bench_n8_wc.json
timeout 900 python scripts/cli_bench.py --gpus 1,8 --reads 400000 > gpurun_out/r2g_cli_bench.json 2> gpurun_out/r2g_cli_bench.err; cat gpurun_out/r2g_cli_bench.json; tail -5 gpurun_out/r2g_cli_bench.err
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or clone or sharded" > gpurun_out/r2g_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2g_pytest_gpu.log
