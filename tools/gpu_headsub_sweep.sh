#!/bin/bash
for h in 64 32 16 8; do
    echo "HEAD_SUB=$h"
    EPC_HEAD_SUB=$h timeout 200 python tools/overlap_probe.py epc-net 1:128 2:128 3:128 2>&1 | grep streams
done
