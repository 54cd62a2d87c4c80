#!/bin/bash
# dropin/build.sh -- builds the reference's UNMODIFIED GPU drivers (nets/*/main.cu + net.cu, compiled from where they lie
# under /root/reference) against this repo's facade, with the reference Makefile's own three nvcc command lines
# (nets/*/Makefile:26-29).  A shadow tree of symlinks stands in for the reference checkout so that the sources' relative
# includes ("../../../lib/GPU/Layer.cuh") and the Makefile's LIB_DIR=../../../lib resolve to dropin/lib/GPU, and
# `-lredcufhe` / "REDcuFHE/redcufhe_gpu.cuh" resolve to dropin/_build/lib/libredcufhe.so / dropin/include through
# LIBRARY_PATH / CPATH, exactly as an installed (RED)cuFHE would.  Outputs only under dropin/_build (git-ignored).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REDSEC_REF:-/root/reference}
B=$ROOT/dropin/_build
T=$B/tree
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CCBIN=""; [ -x /usr/bin/g++ ] && CCBIN="-ccbin /usr/bin/g++"
NETS=${@:-"mnist/sign1024x1 mnist/sign1024x2 mnist/sign1024x3 cifar/binarynet cifar/binarynet_small"}
[ -d "$REF/nets" ] || { echo "reference tree not found at $REF"; exit 0; }
[ -f "$ROOT/redsec_b200/libredsec_b200.so" ] || { echo "build libredsec_b200.so first"; exit 1; }
rm -rf "$B"; mkdir -p "$T/lib/GPU" "$T/client" "$B/lib"
export CPATH="$ROOT/dropin/include:$ROOT/include:$ROOT/redsec_b200/host"
export LIBRARY_PATH="$B/lib:$ROOT/redsec_b200"
# -lredcufhe
$NVCC $CCBIN -O2 -std=c++17 -shared -Xcompiler -fPIC -o "$B/lib/libredcufhe.so" "$ROOT/dropin/src/facade.cpp" -lredsec_b200
# $(LIB_DIR)/GPU: headers + the object the reference link line globs
for h in Layer.cuh IntLayer.cuh BinLayer.cuh; do ln -s "$ROOT/dropin/lib/GPU/$h" "$T/lib/GPU/$h"; done
$NVCC $CCBIN -O2 -std=c++17 -x cu -c -o "$T/lib/GPU/layers_shim.o" "$ROOT/dropin/src/layers_shim.cpp" -I"$ROOT/dropin/lib/GPU" -Xcompiler -fopenmp
for net in $NETS; do
  set_=$(dirname "$net")
  mkdir -p "$T/nets/$net"
  for f in main.cu net.cu net.cuh net.h; do [ -f "$REF/nets/$net/$f" ] && ln -sf "$REF/nets/$net/$f" "$T/nets/$net/$f"; done
  for f in "$REF/nets/$set_"/*.h; do ln -sf "$f" "$T/nets/$set_/$(basename "$f")"; done
  ln -sf "../../../../../../data/nets/$net/var_prep.dat" "$T/nets/$net/var_prep.dat"
  ( cd "$T/nets/$net"
    LIB_DIR=../../../lib
    JOIN_FLAGS_GPU="-I. -I$LIB_DIR -g"
    # the three commands of the gpu-encrypt target, verbatim (plus -ccbin for this image's host compiler and the engine library)
    $NVCC $CCBIN -c -o net_gpu.o net.cu $JOIN_FLAGS_GPU -lredcufhe -Xcompiler -fopenmp -Xcompiler -Wall -Xcompiler -DGPU_ENC
    $NVCC $CCBIN -c -o main_gpu.o main.cu $JOIN_FLAGS_GPU -lredcufhe -Xcompiler -fopenmp -Xcompiler -Wall -Xcompiler -DGPU_ENC
    $NVCC $CCBIN -o gpu-encrypt.out $LIB_DIR/GPU/*.o net_gpu.o main_gpu.o $JOIN_FLAGS_GPU -lredcufhe -lredsec_b200 -Xcompiler -fopenmp -Xcompiler -Wall -Xcompiler -DGPU_ENC )
  # the source symlinks were only needed while compiling: drop them so that nothing under the repo resolves to reference code
  rm -f "$T/nets/$net/main.cu" "$T/nets/$net/net.cu" "$T/nets/$net/net.cuh" "$T/nets/$net/net.h"
  echo "built $T/nets/$net/gpu-encrypt.out"
done
find "$T/nets" -maxdepth 2 -name '*.h' -type l -delete
