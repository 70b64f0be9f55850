# round-2 run H (1 GPU): adaptive grids of the hash stage (leave only while probes are running) -- whole GPU suite, default bench,
# A/B against fixed small grids and the serial schedule
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; rc=$?; tail -3 gpurun_out/r2h_pytest_gpu.log
if [ $rc -ne 0 ]; then echo "GPU TESTS FAILED"; exit 1; fi
timeout 1500 python bench.py --cpu-seconds 8 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -2 gpurun_out/r2h_bench.err; python scripts/show_bench.py gpurun_out/r2h_bench.json
B="timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2h_$tag.json 2> gpurun_out/r2h_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2h_$tag.json; }
run adaptive TXR_X=0
run fixed TXR_ADAPTIVE=0
run serial TXR_OVERLAP=0
