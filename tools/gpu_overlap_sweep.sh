#!/bin/bash
# sweep persistent-grid limits under the two-stream overlap (tools/overlap_probe.py)
for h in 148 120 96 72; do
  for b in 148 100; do
    echo "HEAD_CTAS=$h BLOCK_CTAS=$b"
    EPC_HEAD_CTAS=$h EPC_BLOCK_CTAS=$b timeout 200 python tools/overlap_probe.py epc-net 2:128 3:128 2>&1 | grep streams
  done
done
