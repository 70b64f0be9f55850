# 2-GPU weak-scaling check of bench.py (launched the way the driver does) + the 1-GPU line on the same box
set -x
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err
tail -3 gpurun_out/bench_r1_n2.err; python scripts/show_bench.py gpurun_out/bench_r1_n2.json
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; python scripts/show_bench.py gpurun_out/bench_r1_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_r1_n2_ref.json 2> gpurun_out/bench_r1_n2_ref.err; cut -c1-200 gpurun_out/bench_r1_n2_ref.json
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
