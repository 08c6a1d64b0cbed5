#!/bin/bash
# run tools/sweep.py for the default library and every variant under mp-sort_b200/variants
cd "$(dirname "$0")/.."
python tools/sweep.py "$@"
for f in mp-sort_b200/variants/*.so; do MPSORT_LIB=$PWD/$f python tools/sweep.py "$@"; done
