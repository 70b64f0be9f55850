# round-2 run L (2 GPUs, short): torchrun plumbing of the final state on a small index (rank 0 builds the native CPU-arm library, the
# other rank loads it; parity on both ranks)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --genome-len 4000000 --reads 200000 --steps 2 --warmup 3 > gpurun_out/r2l_bench_n2_small.json 2> gpurun_out/r2l_bench_n2_small.err; tail -3 gpurun_out/r2l_bench_n2_small.err; python scripts/show_bench.py gpurun_out/r2l_bench_n2_small.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_n2_small.json').read().strip().splitlines()[-1])
print(d["parity_at_scale"], d["e2e"]["per_rank"]["rows"], d["host_threads_per_rank"])
PY
