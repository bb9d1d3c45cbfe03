#!/bin/bash
for rep in 1 2; do for wr in 1 0; do
  echo -n "wres=$wr "; MMAD_CONV_WRES=$wr PYTHONPATH=. python tools/r50_graph.py 18 16 2>&1 | tail -1
  echo -n "wres=$wr "; MMAD_CONV_WRES=$wr python tools/unet_prof.py eval 8 5
done; done
