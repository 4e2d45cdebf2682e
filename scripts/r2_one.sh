#!/bin/bash
# quick one-GPU test of selected test files: scripts/r2_one.sh <tag> <pytest args...>
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest "$@" -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest.txt
