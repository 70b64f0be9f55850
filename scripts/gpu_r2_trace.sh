# stage timeline of one end-to-end call and one device-resident call at configs[1] (TXR_TRACE=1)
set -x
mkdir -p gpurun_out
TXR_TRACE=1 timeout 900 python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2trace_bench.json 2> gpurun_out/r2trace.err
grep "txr trace" gpurun_out/r2trace.err | tail -40
python scripts/show_bench.py gpurun_out/r2trace_bench.json
