# one gpurun call: GPU parity suite, sanitizer, then A/B runs of the bench at configs[1] (index built once, cached in /dev/shm)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
bash scripts/gpu_sanitize.sh
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
TXR_EARLY_EXIT=0 timeout 900 $B > gpurun_out/sw_noexit.json 2> gpurun_out/sw.err
timeout 300 $B > gpurun_out/sw_exit.json 2>> gpurun_out/sw.err
timeout 300 $B --error-rate 0.05 > gpurun_out/sw_exit_er005.json 2>> gpurun_out/sw.err
TXR_EARLY_EXIT=0 timeout 300 $B --error-rate 0.05 > gpurun_out/sw_noexit_er005.json 2>> gpurun_out/sw.err
tail -5 gpurun_out/sw.err
for f in gpurun_out/sw_*exit*.json; do echo $f; python scripts/show_bench.py $f; done
