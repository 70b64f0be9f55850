# CLI file-to-file timing on a 1 GB index (short call): plain FASTQ, gzip, BGZF; plus the CLI parity tests
set -x
timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --genome-len 4000000 --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-150
timeout 900 python scripts/cli_bench.py --genome-len 4000000 --gz --bgzf > gpurun_out/cli_bench_small.json 2> gpurun_out/cli_bench_small.err; tail -3 gpurun_out/cli_bench_small.err; cat gpurun_out/cli_bench_small.json
