#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_gpu_patch_embed.py -x -q --timeout 300 2>&1 | tail -6
