# one gpurun call: GPU parity suite, then A/B runs of the bench at configs[1] (the index is built once and cached in /dev/shm)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
TXR_LEVEL_CTA=0 timeout 900 $B > gpurun_out/sw_warpitem_128k.json 2> gpurun_out/sw.err
timeout 300 $B > gpurun_out/sw_ctaitem_128k.json 2>> gpurun_out/sw.err
TXR_LEVEL_CTA=0 timeout 300 $B --batch-reads 262144 > gpurun_out/sw_warpitem_256k.json 2>> gpurun_out/sw.err
timeout 300 $B --batch-reads 262144 > gpurun_out/sw_ctaitem_256k.json 2>> gpurun_out/sw.err
TXR_L2_HINTS=0 timeout 300 $B > gpurun_out/sw_ctaitem_nohint_128k.json 2>> gpurun_out/sw.err
tail -5 gpurun_out/sw.err
for f in gpurun_out/sw_*item*.json; do echo $f; python scripts/show_bench.py $f; done
S="python bench.py --reads 262144 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"ixf_query" -c 12 --csv --log-file gpurun_out/launches_cta.csv $S > /dev/null 2>> gpurun_out/sw.err
python scripts/launch_summary.py gpurun_out/launches_cta.csv
grep lts__t_sector_hit gpurun_out/launches_cta.csv | cut -d, -f5,15 | head -12
