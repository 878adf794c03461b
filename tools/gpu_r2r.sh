#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
echo base; timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-extra 2>&1 | tail -1 | cut -c130-170
for v in skipf skipfr skipr; do echo $v; B200FBANK_LIB=$PWD/tools/build/w_$v.so timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-extra 2>&1 | tail -1 | cut -c130-170; done
