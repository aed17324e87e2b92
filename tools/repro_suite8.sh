#!/bin/bash
# peer-store experiments on 8 GPUs: bash tools/repro_suite8.sh
tr() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/repro_determinism.py 65536 256 16 12 2>&1 | grep -E "DETERMINISM|rank 0 fit|Error|error" | cut -c1-330 | head -16; }
echo "--- default"; tr
echo "--- NO_PEER"; FXB200_NO_PEER=1 tr
echo "--- PEER_MODE=1 plain remote stores"; FXB200_PEER_MODE=1 tr
echo "--- PEER_MODE=2 system fence"; FXB200_PEER_MODE=2 tr
echo "--- PEER_MODE=3 host sync + 2 barriers"; FXB200_PEER_MODE=3 tr
