#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
B200FBANK_LIB=$PWD/tools/build/ws_timing.so timeout 120 python tools/ws_timing.py 2>&1 | tail -3
