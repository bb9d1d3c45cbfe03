#!/bin/bash
# weight-gradient split-K sweep on the ResNet3D-18 layer shapes (batch 16): MMAD_WG_NSPLIT overrides the model's choice
for cfg in "16 16 512 512 3 1 4" "16 16 256 256 3 1 2" "16 16 128 128 3 1 1" "16 16 256 512 3 1 4"; do
  echo "== $cfg"
  MMAD_WG_DEBUG=1 python tools/wgrad_prof.py wgrad $cfg 5 2>&1 | grep -E "nsplit|ms" | tail -2
  for ns in 3 4 5 6 8 10 12; do echo -n "nsplit=$ns "; MMAD_WG_NSPLIT=$ns python tools/wgrad_prof.py wgrad $cfg 5 2>/dev/null | tail -1; done
done
