# 2-rank flow check of bench.py launched the way the driver does (torchrun, one rank per GPU).  A 1 GB index keeps the
# call short; what is being checked is the multi-rank plumbing: OpenMP threads under torchrun (OMP_NUM_THREADS=1 is
# exported), rank 0 building while the others wait, max-over-ranks timing, and the reference arm's rank handling.
set -x
G="--genome-len 4000000"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 $G > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err
tail -3 gpurun_out/bench_r1_n2.err; python scripts/show_bench.py gpurun_out/bench_r1_n2.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_n2.json').read().strip().splitlines()[-1]); print(d['index_build'], d['n_gpus'], d['config']['parallelism'])"
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 $G --no-cpu-baseline > gpurun_out/bench_r1_n1_small.json 2> gpurun_out/bench_r1_n1_small.err; python scripts/show_bench.py gpurun_out/bench_r1_n1_small.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 $G > gpurun_out/bench_r1_n2_ref.json 2> gpurun_out/bench_r1_n2_ref.err; cut -c1-200 gpurun_out/bench_r1_n2_ref.json
