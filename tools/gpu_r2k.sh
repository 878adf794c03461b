#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for r in 0 1 2 3; do for m in 0 2; do timeout 60 python tools/us8k_seed.py $r $m 2>&1 | grep -E "US8KSEED|Error" | tail -1; done; done
for i in 1 2 3; do timeout 60 python tools/us8k_seed.py 1 2>&1 | grep -E "US8KSEED|Error" | tail -1; done
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -3
timeout 100 python tools/ktime.py --us8k --tag fix 2>&1 | grep KTIME
