#!/bin/bash
# Run-to-run determinism of the multi-GPU fit on G GPUs of one box: bash tools/repro_suite.sh <G> [fits]
# (n = 8192 rows per rank; every rank's lambda / trace compared bitwise with its first fit).
# The cross-rank symmetric Gram schedule (FXB200_PEER_SYM=1) must show 0 differing fits before it can be
# the default again: profiles/r02_peer_sym_investigation.md.
G=${1:-4}; FITS=${2:-300}
N=$((8192 * G))
tr() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/repro_determinism.py $N 256 16 $FITS 2>&1 | grep -E "DETERMINISM|DIFFERS|FAILED" | cut -c1-300 | head -12; }
echo "--- default (rank-local Gram)"; tr
echo "--- FXB200_PEER_SYM=1 (cross-rank symmetric Gram)"; FXB200_PEER_SYM=1 tr
