#!/bin/bash
# A/B helper for the GPU box: rebuild the library with extra nvcc flags, then run the given command.
#   tools/ab_build.sh "-DGHB_MINB34=5" python tools/perf_configs.py
extra="$1"; shift
GHB_NVCC_EXTRA="$extra" python -c "import sys; sys.path.insert(0, '.'); import gridaphybrid_b200._lib as m; m.build(force=True)" || exit 1
echo "== build flags: $extra"
"$@"
