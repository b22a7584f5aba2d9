#!/bin/bash
# round 2, call G: per-query counts, CLI after the context-thread change
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_query_counts.py tests/test_cli_parity.py -m gpu -x -q > $OUT/r2g_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2g_tests.log
tail -n 30 $OUT/r2g_tests.log
