#!/bin/bash
# round 2, call W: genomic_regions link / inv against the reference binary, gtb_link_regions against the sequential loop
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_regions_ops.py tests/test_abi_exports.py -x -q > $OUT/r2w_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2w_tests.log
tail -15 $OUT/r2w_tests.log
