# round-2 run K (1 GPU): the host pass under torchrun's OMP_NUM_THREADS=1 -- bench.py now sets the OpenMP thread count itself
set -x
mkdir -p gpurun_out
B="timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err; echo "== $tag $@"; python scripts/show_bench.py gpurun_out/r2k_$tag.json; }
run default TXR_X=0
run omp1_env_fixed OMP_NUM_THREADS=1
run omp1_forced OMP_NUM_THREADS=1 TAXOR_BENCH_OMP=1
run omp4 TAXOR_BENCH_OMP=4
