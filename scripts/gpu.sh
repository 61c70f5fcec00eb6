#!/bin/bash
# Build in-tree (so the .so that travels matches the sources), then run one command on a B200 box.
#   scripts/gpu.sh [--gpus N] <timeout seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
python -c "import __graft_entry__ as g; g.build()" >/dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
