#!/bin/bash
# Runs one pytest id with VDB_TRACE=1 (sync + log after every kernel) under a hard timeout; keeps the log tail.
mkdir -p gpurun_out
VDB_TRACE=1 timeout -s KILL ${2:-60} python -m pytest "$1" -q -x -s -p no:cacheprovider > gpurun_out/trace.log 2>&1
echo "exit $?"; grep -E "^\[vdb\]" gpurun_out/trace.log | tail -6; grep -E "^E  |passed|failed|parity\]" gpurun_out/trace.log | head
