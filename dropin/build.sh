#!/bin/bash
# dropin/build.sh -- builds the reference's UNMODIFIED GPU drivers (nets/*/main.cu + net.cu, compiled from where they lie
# under /root/reference) against this repo's facade, with the reference Makefile's own three nvcc command lines
# (nets/*/Makefile:26-29).  A shadow tree of symlinks stands in for the reference checkout so that the sources' relative
# includes ("../../../lib/GPU/Layer.cuh") and the Makefile's LIB_DIR=../../../lib resolve to dropin/lib/GPU, and
# `-lredcufhe` / "REDcuFHE/redcufhe_gpu.cuh" resolve to dropin/_build/lib/libredcufhe.so / dropin/include through
# LIBRARY_PATH / CPATH, exactly as an installed (RED)cuFHE would.  Outputs only under dropin/_build (git-ignored).
#
# Trees (each one a stand-in for a reference checkout whose lib/GPU was replaced):
#   tree            Layer-level facade: this repo's IntLayer / BinLayer (dropin/lib/GPU/{Int,Bin}Layer.cuh), 1 GPU, all nets
#   tree_func       Func-level facade: the REFERENCE's own lib/GPU/{Bin,Int}Layer.{cu,cuh}, unmodified, compiled against this
#                   repo's BinFunc_gpu.cuh / IntFunc_gpu.cuh / BinOps_gpu.cuh / IntOps_gpu.cuh / gates.cuh / Layer.cuh
#   tree_g<N>       Layer-level facade with -DNUM_GPUS=<N> (the reference edits `#define NUM_GPUS` in lib/GPU/Layer.cuh:15)
#   tree_func_g<N>  Func-level facade with -DNUM_GPUS=<N>
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REDSEC_REF:-/root/reference}
B=$ROOT/dropin/_build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CCBIN=""; [ -x /usr/bin/g++ ] && CCBIN="-ccbin /usr/bin/g++"
[ -d "$REF/nets" ] || { echo "reference tree not found at $REF"; exit 0; }
[ -f "$ROOT/redsec_b200/libredsec_b200.so" ] || { echo "build libredsec_b200.so first"; exit 1; }
rm -rf "$B"; mkdir -p "$B/lib"
export CPATH="$ROOT/dropin/include:$ROOT/include:$ROOT/redsec_b200/host"
export LIBRARY_PATH="$B/lib:$ROOT/redsec_b200"
# -lredcufhe
$NVCC $CCBIN -O2 -std=c++17 -shared -Xcompiler -fPIC -o "$B/lib/libredcufhe.so" "$ROOT/dropin/src/facade.cpp" -lredsec_b200

# build_tree <tree name> <layer|func> <num gpus> <nets...>
build_tree() {
  local T=$B/$1 MODE=$2 NG=$3; shift 3
  local DEF="-DNUM_GPUS=$NG"
  mkdir -p "$T/lib/GPU" "$T/client"
  # $(LIB_DIR)/GPU: headers + the objects the reference link line globs ($(LIB_DIR)/GPU/*.o)
  for h in Layer.cuh gates.cuh BinOps_gpu.cuh IntOps_gpu.cuh BinFunc_gpu.cuh IntFunc_gpu.cuh; do ln -sf "$ROOT/dropin/lib/GPU/$h" "$T/lib/GPU/$h"; done
  local SHIMS="layer_util ops_shim func_shim"
  if [ "$MODE" = layer ]; then
    for h in IntLayer.cuh BinLayer.cuh; do ln -sf "$ROOT/dropin/lib/GPU/$h" "$T/lib/GPU/$h"; done
    SHIMS="$SHIMS layers_shim"
  else
    # the reference's own layer sequencing code, from where it lies, over this repo's Func classes
    for f in IntLayer.cuh BinLayer.cuh IntLayer.cu BinLayer.cu; do ln -sf "$REF/lib/GPU/$f" "$T/lib/GPU/$f"; done
    for f in IntLayer BinLayer; do
      ( cd "$T/lib/GPU" && $NVCC $CCBIN -c -o $f.o $f.cu -g -I/usr/local/include $DEF -Xcompiler -fopenmp )   # lib/Makefile:19
    done
  fi
  for s in $SHIMS; do
    $NVCC $CCBIN -O2 -std=c++17 -x cu -c -o "$T/lib/GPU/$s.o" "$ROOT/dropin/src/$s.cpp" -I"$T/lib/GPU" $DEF -Xcompiler -fopenmp
  done
  for net in "$@"; do
    local set_=$(dirname "$net")
    mkdir -p "$T/nets/$net"
    for f in main.cu net.cu net.cuh net.h; do [ -f "$REF/nets/$net/$f" ] && ln -sf "$REF/nets/$net/$f" "$T/nets/$net/$f"; done
    for f in "$REF/nets/$set_"/*.h; do ln -sf "$f" "$T/nets/$set_/$(basename "$f")"; done
    ln -sf "$ROOT/data/nets/$net/var_prep.dat" "$T/nets/$net/var_prep.dat"
    ( cd "$T/nets/$net"
      LIB_DIR=../../../lib
      JOIN_FLAGS_GPU="-I. -I$LIB_DIR -g $DEF"
      # the three commands of the gpu-encrypt target, verbatim (plus -ccbin for this image's host compiler, the NUM_GPUS
      # define the reference sets by editing its header, and the engine library)
      $NVCC $CCBIN -c -o net_gpu.o net.cu $JOIN_FLAGS_GPU -lredcufhe -Xcompiler -fopenmp -Xcompiler -Wall -Xcompiler -DGPU_ENC
      $NVCC $CCBIN -c -o main_gpu.o main.cu $JOIN_FLAGS_GPU -lredcufhe -Xcompiler -fopenmp -Xcompiler -Wall -Xcompiler -DGPU_ENC
      $NVCC $CCBIN -o gpu-encrypt.out $LIB_DIR/GPU/*.o net_gpu.o main_gpu.o $JOIN_FLAGS_GPU -lredcufhe -lredsec_b200 -Xcompiler -fopenmp -Xcompiler -Wall -Xcompiler -DGPU_ENC )
    # the source symlinks were only needed while compiling: drop them so that nothing under the repo resolves to reference code
    rm -f "$T/nets/$net/main.cu" "$T/nets/$net/net.cu" "$T/nets/$net/net.cuh" "$T/nets/$net/net.h"
    echo "built $T/nets/$net/gpu-encrypt.out"
  done
  find "$T/nets" -maxdepth 2 -name '*.h' -type l -delete
  rm -f "$T/lib/GPU/IntLayer.cu" "$T/lib/GPU/BinLayer.cu"
  [ "$MODE" = func ] && rm -f "$T/lib/GPU/IntLayer.cuh" "$T/lib/GPU/BinLayer.cuh"
  return 0
}

# a caller of the Ops-level surface (gates.cuh / BinOps:: / IntOps::), linked like a net driver
build_ops_check() {
  local T=$B/$1
  mkdir -p "$T/opscheck"
  ( cd "$T/opscheck"
    $NVCC $CCBIN -O2 -std=c++17 -x cu -c -o ops_check.o "$ROOT/dropin/src/ops_check.cpp" -I"$T/lib/GPU" -Xcompiler -fopenmp
    $NVCC $CCBIN -o ops_check.out "$T"/lib/GPU/*.o ops_check.o -lredcufhe -lredsec_b200 -Xcompiler -fopenmp )
  echo "built $T/opscheck/ops_check.out"
}

if [ $# -gt 0 ]; then
  build_tree tree layer 1 "$@"
  build_ops_check tree
else
  build_tree tree layer 1 mnist/sign1024x1 mnist/sign1024x2 mnist/sign1024x3 cifar/binarynet cifar/binarynet_small
  build_ops_check tree
  build_tree tree_func func 1 mnist/sign1024x1 cifar/binarynet_small
  build_tree tree_g2 layer 2 mnist/sign1024x1 cifar/binarynet_small
  build_tree tree_func_g2 func 2 cifar/binarynet_small
  build_tree tree_g8 layer 8 cifar/binarynet
fi
