#!/bin/bash
# determinism experiments: bash tools/repro_suite.sh <ngpus>
G=${1:-4}
tr() { local g=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/repro_determinism.py "$@" 2>&1 | grep -E "DETERMINISM|DIFFERS|Error|error" | head -12; }
N=$((8192 * G))
echo "--- default"; tr $G $N 256 16 10
echo "--- NO_PEER"; FXB200_NO_PEER=1 tr $G $N 256 16 10
echo "--- MAX_BLOCKS=3"; FXB200_MAX_BLOCKS=3 tr $G $N 256 16 10
echo "--- NO_OVERLAP"; FXB200_NO_OVERLAP=1 tr $G $N 256 16 10
echo "--- NO_LAG"; FXB200_NO_LAG=1 tr $G $N 256 16 10
if [ "$G" -ge 4 ]; then echo "--- 2 ranks n=16384"; tr 2 16384 256 16 10; fi
