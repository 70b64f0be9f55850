# one gpurun call: GPU parity suite, then A/B runs of the bench at configs[1] (the index is built once and cached in /dev/shm)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
TXR_L2_HINTS=0 timeout 900 $B > gpurun_out/sw_nohint_128k.json 2> gpurun_out/sw.err
timeout 300 $B > gpurun_out/sw_hint_128k.json 2>> gpurun_out/sw.err
TXR_L2_HINTS=0 timeout 300 $B --batch-reads 262144 > gpurun_out/sw_nohint_256k.json 2>> gpurun_out/sw.err
timeout 300 $B --batch-reads 262144 > gpurun_out/sw_hint_256k.json 2>> gpurun_out/sw.err
tail -5 gpurun_out/sw.err
for f in gpurun_out/sw_*hint*.json; do echo $f; python scripts/show_bench.py $f; done
timeout 900 python scripts/cli_bench.py --gz > gpurun_out/cli_bench_r1.json 2> gpurun_out/cli_bench.err; cat gpurun_out/cli_bench_r1.json; tail -5 gpurun_out/cli_bench.err
