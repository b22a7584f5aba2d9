#!/bin/bash
# round 2, call P: genomic_scans peaks, BAM input -- CLI parity tests (whole CLI suite), scan tests
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_cli_parity.py -m gpu -x -q > $OUT/r2p_cli.log 2>&1
echo "tests rc=$?" >> $OUT/r2p_cli.log
tail -15 $OUT/r2p_cli.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q -k "scan" > $OUT/r2p_scan.log 2>&1
echo "tests rc=$?" >> $OUT/r2p_scan.log
tail -3 $OUT/r2p_scan.log
