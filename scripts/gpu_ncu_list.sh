# launch list (device time per launch) of a short bench run on a 1 GB index; $1 = tag, extra env passes through
set -x
TAG=${1:-x}
B="python bench.py --genome-len 4000000 --reads 262144 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $B > gpurun_out/ncu_${TAG}_plain.json 2> gpurun_out/ncu_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $B > /dev/null 2>> gpurun_out/ncu_${TAG}.err
python scripts/show_bench.py gpurun_out/ncu_${TAG}_plain.json
tail -3 gpurun_out/ncu_${TAG}.err
