#!/bin/bash
# A/B of prebuilt library variants (tools/_bin/libghb_*.so): throughput of the cell-warp kernel on the given shapes
shapes="${1:-34,36}"
echo "== default"; python tools/ab_cw.py 20 "$shapes" 2>&1 | grep "cells:"
for so in tools/_bin/libghb_*.so; do
  echo "== $so"; GHB_LIB_PATH=$so python tools/ab_cw.py 20 "$shapes" 2>&1 | grep "cells:"
done
