#!/bin/bash
# ResNet3D-50 / ResNet3D-18 eager steps: old static kernels vs the tile ring (static / dynamic)
for rep in 1 2; do
  for cfg in "OLD" "DYN1" "DYN0"; do
    case $cfg in
      OLD) export MMAD_LIB=multimodal_ad_b200/csrc/libmmad_b200_old.so; unset MMAD_CONV_DYN;;
      DYN1) unset MMAD_LIB; export MMAD_CONV_DYN=1;;
      DYN0) unset MMAD_LIB; export MMAD_CONV_DYN=0;;
    esac
    echo -n "$cfg r50 "; python tools/resnet_bench.py 8 128 5 50 2 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['enqueue_ms'])"
    echo -n "$cfg r18 "; python tools/resnet_bench.py 16 128 5 18 2 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['enqueue_ms'])"
  done
done
