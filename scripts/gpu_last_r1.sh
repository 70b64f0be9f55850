# last check of the round on one GPU: full parity suite, smoke(), the default bench line (with the at-scale parity check),
# then the three-level workload with the same check
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -3 gpurun_out/bench_r1_final.err; python scripts/show_bench.py gpurun_out/bench_r1_final.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_final.json').read().strip().splitlines()[-1]); print(d['parity_at_scale']); print(d['index_build'])"
rm -rf /dev/shm/taxor_b200_bench
timeout 1500 python bench.py --workload deep --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_r1_deep.json 2> gpurun_out/bench_r1_deep.err; tail -2 gpurun_out/bench_r1_deep.err; python scripts/show_bench.py gpurun_out/bench_r1_deep.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_deep.json').read().strip().splitlines()[-1]); print(d['parity_at_scale']); print(d['index_build'])"
