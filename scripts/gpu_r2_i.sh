# round-2 run I (1 GPU): adaptive grids with the per-read re-check; whole GPU suite; default bench; A/B against fixed small grids
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; rc=$?; tail -3 gpurun_out/r2i_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
timeout 1500 python bench.py --cpu-seconds 8 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -2 gpurun_out/r2i_bench.err; python scripts/show_bench.py gpurun_out/r2i_bench.json
B="timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2i_$tag.json; }
run adaptive TXR_X=0
run fixed TXR_ADAPTIVE=0
TAXOR_TIMING=1 timeout 600 python scripts/cli_bench.py > gpurun_out/r2i_cli_bench.json 2> gpurun_out/r2i_cli_bench.err; cut -c1-900 gpurun_out/r2i_cli_bench.json
