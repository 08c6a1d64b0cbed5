#!/bin/bash
# run tools/sweep.py for the default library and every variant under mp-sort_b200/variants
# (each under its own timeout: an experimental kernel that hangs must not take the call with it)
cd "$(dirname "$0")/.."
timeout 150 python tools/sweep.py "$@"
for f in mp-sort_b200/variants/*.so; do MPSORT_LIB=$PWD/$f timeout 150 python tools/sweep.py "$@" || echo "$f: failed or timed out"; done
