#!/bin/bash
# Builds the library with the GNN phase stamps compiled in, runs the trace, rebuilds the product library.
#   bash tools/probes/gnn_trace.sh build    (here)      then on the GPU box: python tools/probes/gnn_trace.py
#   bash tools/probes/gnn_trace.sh restore  (here)
set -e
cd "$(dirname "$0")/../.."
PKG=audio-to-motion-generation_b200
if [ "$1" = build ]; then
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include \
    --expt-relaxed-constexpr -DA2M_GNN_TRACE -c $PKG/csrc/gnn_fused.cu -o $PKG/build/gnn_fused.o
  rm -f $PKG/liba2m_b200.so
  python $PKG/build.py
else
  touch $PKG/csrc/gnn_fused.cu
  python $PKG/build.py
fi
