#!/bin/bash
# round 2, call AG: the drivers leaving by _exit() after the flush (A/B against exit()), then the command-line parity tests with it
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python profiles/scripts/time_cli_exit.py > $OUT/r2ag_cli_exit.json 2> $OUT/r2ag_cli_exit.err; echo "ab rc=$?"; cat $OUT/r2ag_cli_exit.json; tail -3 $OUT/r2ag_cli_exit.err
timeout 420 python -m pytest tests/test_cli_parity.py tests/test_regions_ops.py tests/test_dropin_shim.py tests/test_mgpu.py -m gpu -x -q --durations=5 > $OUT/r2ag_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2ag_tests.log
tail -12 $OUT/r2ag_tests.log
