#!/bin/bash
# Debug build with the in-kernel clock64 timelines (conv_ws tiles, NMS phases) -> build_dbg/libyolo_b200_dbg.so
# (loaded by tools/ws_timeline.py).  The release objects of the other sources are reused: run build() first.
set -e
PKG="$(cd "$(dirname "$0")/.." && pwd)/yolo-compression-and-deployment-in-fpga_b200"
mkdir -p "$PKG/build_dbg"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
nvcc $FLAGS -DYB_WS_TIMELINE -c "$PKG/csrc/conv_ws.cu" -o "$PKG/build_dbg/conv_ws.o" &
nvcc $FLAGS -DYB_WS_TIMELINE ${YB_RP_EXP:+-DYB_RP_EXP=$YB_RP_EXP} -c "$PKG/csrc/conv_rp.cu" -o "$PKG/build_dbg/conv_rp.o" &
nvcc $FLAGS -DYB_WS_TIMELINE -c "$PKG/csrc/conv_wsp.cu" -o "$PKG/build_dbg/conv_wsp.o" &
nvcc $FLAGS -DYB_WS_TIMELINE -c "$PKG/csrc/conv_ws2.cu" -o "$PKG/build_dbg/conv_ws2.o" &
nvcc $FLAGS -DYB_NMS_TIMELINE -c "$PKG/csrc/head.cu" -o "$PKG/build_dbg/head.o" &
nvcc $FLAGS -DYB_FS_TIMELINE -c "$PKG/csrc/conv_fs.cu" -o "$PKG/build_dbg/conv_fs.o" &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$PKG/build_dbg/libyolo_b200_dbg${YB_RP_EXP}.so" "$PKG/build/yolo_b200.o" "$PKG/build/conv_direct.o" \
  "$PKG/build/conv_first.o" "$PKG/build_dbg/conv_fs.o" "$PKG/build/resize.o" "$PKG/build/graph.o" "$PKG/build/conv_umma.o" "$PKG/build_dbg/conv_ws.o" "$PKG/build_dbg/conv_rp.o" "$PKG/build_dbg/conv_wsp.o" "$PKG/build_dbg/conv_ws2.o" "$PKG/build/conv_ws3.o" "$PKG/build/quantize.o" "$PKG/build_dbg/head.o" -cudart static
echo built "$PKG/build_dbg/libyolo_b200_dbg${YB_RP_EXP}.so"
