#!/bin/bash
# round 2, call J: compute-sanitizer memcheck on smoke() and on the byte-counter overflow / replay test (one tool per call)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2j_plain.log 2>&1; echo "plain smoke rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/r2j_memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2j_smoke.out 2>&1; echo "memcheck smoke rc=$?"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/r2j_memcheck_tests.log python -m pytest tests/test_gpu_parity.py tests/test_query_counts.py -m gpu -x -q -k "counter_overflow or sort_regions or packed_reads or fatal_queries or test_query_counts_vs_brute_force or multi_interval" > $OUT/r2j_tests.out 2>&1; echo "memcheck tests rc=$?"
tail -n 5 $OUT/r2j_memcheck_smoke.log $OUT/r2j_smoke.out $OUT/r2j_memcheck_tests.log $OUT/r2j_tests.out
