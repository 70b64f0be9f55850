# round-2 run M (1 GPU): cumulative cap on the batch sizes of a host-fed search (TXR_RAMP_CUM), end-to-end step
set -x
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2m_$tag.json 2> gpurun_out/r2m_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2m_$tag.json; }
run off TXR_X=0
run cum035 TXR_RAMP_CUM=0.35
run cum05 TXR_RAMP_CUM=0.5
