#!/bin/bash
# build in-tree, then run a command on a B200 box:  tools/grun.sh <timeout_s> '<command>'
cd "$(dirname "$0")/.." || exit 1
python -c "from trueno_rag_b200 import build as b; b.build()" || exit 1
make -C oracle -s || exit 1
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
