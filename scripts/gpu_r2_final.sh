# round-2 final check (1 GPU), what the driver runs at round end: GPU suite, smoke(), default bench line, reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2final_pytest_gpu.log 2>&1; rc=$?; tail -3 gpurun_out/r2final_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; tail -2 gpurun_out/r2final_smoke.log
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2final_bench_ref.json 2> gpurun_out/r2final_bench_ref.err; cut -c1-200 gpurun_out/r2final_bench_ref.json
timeout 1500 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; tail -2 gpurun_out/r2final_bench.err; python scripts/show_bench.py gpurun_out/r2final_bench.json
