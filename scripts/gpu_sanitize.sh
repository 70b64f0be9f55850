# compute-sanitizer over the parity tests that drive the kernels (memcheck: all selected; racecheck: shared-memory hazards)
set -x
SEL="minimiser or partitioned or early_exit or percentage_and_batching or large_rows or kmer_mode or threshold_zero or fused_distinct or ixf_scheme_variants or index_clone or (wide_rows and 1536)"
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_memcheck.log
timeout 1800 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "minimiser_hash_parity or partitioned or early_exit or syncmer_hash_parity or (wide_rows and 1536) or large_rows" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitize_racecheck.log
grep -c "ERROR SUMMARY" gpurun_out/sanitize_*.log; grep "ERROR SUMMARY" gpurun_out/sanitize_*.log | sort | uniq -c | head
