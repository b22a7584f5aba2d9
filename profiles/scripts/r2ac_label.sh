#!/bin/bash
# round 2, call AC: overlap -label (gtb_index_query_matches), union on the box
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_query_counts.py tests/test_cli_parity.py tests/test_regions_ops.py tests/test_abi_exports.py -m gpu -x -q -k "query or label or subset or union or abi" > $OUT/r2ac_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2ac_tests.log
tail -30 $OUT/r2ac_tests.log
