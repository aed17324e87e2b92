#!/bin/bash
tr() { timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/repro_determinism.py 32768 256 16 $1 > /tmp/repro_out.log 2>&1; grep -E "DETERMINISM|rank 0 fit" /tmp/repro_out.log | cut -c1-300 | tail -5; grep -E "Error|error|FxError" /tmp/repro_out.log | grep -v "error_file\|traceback" | head -5; }
echo "--- PEER_MODE=5 (L1-bypassing remote partial stores) 400"; FXB200_PEER_MODE=5 tr 400
echo "--- PEER_MODE=6 (no remote partial stores; wrong means but deterministic if the trigger is gone) 400"; FXB200_PEER_MODE=6 tr 400
