#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== nlist 8192 w 64 graph on"; $TR profiles/diag_sharded.py 8192 64 400000 4000 100 2>&1 | grep -E "step|Error|error" | head -20
echo "== nlist 8192 w 64 graph off"; MMIDX_GRAPH=0 $TR profiles/diag_sharded.py 8192 64 400000 4000 100 2>&1 | grep -E "step|Error|error" | head -20
echo "== nlist 512 w 64 graph on"; $TR profiles/diag_sharded.py 512 64 400000 4000 100 2>&1 | grep -E "step|Error|error" | head -20
echo "== nlist 2048 w 64 timings on"; DIAG_TIMINGS=1 $TR profiles/diag_sharded.py 2048 64 400000 4000 100 2>&1 | grep -E "step|Error|error" | head -20
