#!/bin/bash
# round 2, call H: subset / overlap / gsort command-line parity
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_cli_parity.py -m gpu -x -q -k "subset or gsort or usage" > $OUT/r2h_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2h_tests.log
tail -n 40 $OUT/r2h_tests.log | cut -c1-260
