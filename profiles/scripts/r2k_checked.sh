#!/bin/bash
# round 2, call K: the ABI-level GPU tests against the library built with device-side bounds asserts (make checked)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
export GTB200_LIB=$PWD/ibm-cbc-genomic-tools_b200/lib/libgtb200_checked.so
ls -la $GTB200_LIB > $OUT/r2k_checked_tests.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_query_counts.py "tests/test_baseline_configs.py::test_config0_shipped_example_abi" "tests/test_baseline_configs.py::test_config3_paired_coverage_abi" "tests/test_baseline_configs.py::test_scale_limits_weights" "tests/test_baseline_configs.py::test_config2_scans_chr21_chr22" -m gpu -x -q >> $OUT/r2k_checked_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2k_checked_tests.log
tail -n 8 $OUT/r2k_checked_tests.log | cut -c1-250
