# (1) torchrun 2-rank flow check on ONE GPU's worth of budget is not possible; this script is for a 1-GPU box:
#     parity suite + A/B of the SM-partitioned overlap at configs[1]
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
run() { # tag overlap split
  TXR_OVERLAP=$2 TXR_SM_SPLIT=$3 timeout 900 $B > gpurun_out/sp_$1.json 2>> gpurun_out/sp.err
  echo "== $1 (overlap=$2 split=$3)"; python scripts/show_bench.py gpurun_out/sp_$1.json
}
run off 0 0:0
run split4_1 1 4:1
run split5_1 1 5:1
run split3_1 1 3:1
run split6_1 1 6:1
run split8_3 1 8:3
tail -3 gpurun_out/sp.err
