#!/bin/bash
# flakiness check: the live reference tests and the whole suite several times in fresh processes
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests/test_zz_reference_live.py -m gpu -x -q 2>&1 | tail -1
done
for i in 1 2; do
  timeout 900 python -m pytest tests -m gpu -x -q -p no:randomly 2>&1 | tail -1
done
