#!/bin/bash
# Debug build of the library with the in-kernel timeline compiled in (-DVIL_TRACE); used by tools/trace_timeline.py only.
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DVIL_TRACE -shared \
  -o vision_longformer_b200/libvil_attn_sm100_trace.so vision_longformer_b200/csrc/vil_attn_api.cu
