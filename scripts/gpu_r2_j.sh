# round-2 run J (4 GPUs): the end-to-end step under host-side H2D contention -- adaptive grids vs fixed small grids vs the serial schedule
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
timeout 1500 $T bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2j_bench_n4.json 2> gpurun_out/r2j_bench_n4.err; tail -2 gpurun_out/r2j_bench_n4.err; python scripts/show_bench.py gpurun_out/r2j_bench_n4.json
TXR_ADAPTIVE=0 timeout 900 $T bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2j_bench_n4_fixed.json 2> gpurun_out/r2j_bench_n4_fixed.err; python scripts/show_bench.py gpurun_out/r2j_bench_n4_fixed.json
TXR_OVERLAP=0 timeout 900 $T bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2j_bench_n4_serial.json 2> gpurun_out/r2j_bench_n4_serial.err; python scripts/show_bench.py gpurun_out/r2j_bench_n4_serial.json
