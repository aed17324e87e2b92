#!/bin/bash
tr() { timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/repro_determinism.py 32768 256 16 $1 2>&1 | grep -E "DETERMINISM|rank 0 fit|Error|error" | cut -c1-300 | tail -6; }
echo "--- default 400"; tr 400
echo "--- NO_PEER 600"; FXB200_NO_PEER=1 tr 600
echo "--- PEER_MODE=4 (proxy fence after the full-barrier wait) 400"; FXB200_PEER_MODE=4 tr 400
