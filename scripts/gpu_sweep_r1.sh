# one gpurun call: GPU parity suite, then A/B runs of the bench at configs[1] (the index is built once and cached in /dev/shm)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
TXR_ROOT_PARTITION=0 timeout 900 $B > gpurun_out/sw_nopart_128k.json 2> gpurun_out/sw.err
timeout 300 $B > gpurun_out/sw_part_128k.json 2>> gpurun_out/sw.err
timeout 300 $B --batch-reads 262144 > gpurun_out/sw_part_256k.json 2>> gpurun_out/sw.err
timeout 300 $B --batch-reads 524288 > gpurun_out/sw_part_512k.json 2>> gpurun_out/sw.err
tail -5 gpurun_out/sw.err
for f in gpurun_out/sw_*part*.json; do echo $f; python scripts/show_bench.py $f; done
