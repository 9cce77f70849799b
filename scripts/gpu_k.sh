#!/bin/bash
# Run the GPU tests matching a -k expression.
TAG=${1:-k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q -k "$2" ) > $OUT/pytest.log 2>&1
tail -40 $OUT/pytest.log
