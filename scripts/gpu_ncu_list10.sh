# launch list of the query-stage kernels at configs[1] (10 GB index); $1 = tag
set -x
TAG=${1:-x}
B="python bench.py --reads 262144 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 $B > gpurun_out/ncu10_${TAG}_plain.json 2> gpurun_out/ncu10_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"root_part|ixf_query|items_" -c 120 --csv --log-file gpurun_out/launches10_${TAG}.csv $B > /dev/null 2>> gpurun_out/ncu10_${TAG}.err
python scripts/show_bench.py gpurun_out/ncu10_${TAG}_plain.json
python scripts/launch_summary.py gpurun_out/launches10_${TAG}.csv
tail -3 gpurun_out/ncu10_${TAG}.err
