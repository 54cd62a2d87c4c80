#!/bin/bash
# oracle/build_ref.sh -- compiles the reference's own PLAINTEXT twin of lib/ (no TFHE needed) from the
# sources where they lie under /root/reference, together with oracle/ref_harness.cpp, into oracle/_ref/ptxt_<net>.
# Only runs where /root/reference exists (this container); nothing is copied into the repo.  The encrypted path of
# the reference cannot be built: tfhe/tfhe.h, libtfhe-spqlios-fma and libredcufhe are absent.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/lib" ] || { echo "no reference tree at $REF; skipping oracle/_ref"; exit 0; }
mkdir -p "$OUT"
CXX=/usr/bin/g++
LIBSRC="BinFunc.cpp BinLayer.cpp BinOps.cpp IntFunc.cpp IntLayer.cpp IntOps.cpp Layer.cpp"
for net in mnist/sign1024x1 mnist/sign1024x2 mnist/sign1024x3 cifar/binarynet cifar/binarynet_small \
           mnist/relu1024x1 mnist/relu1024x2 mnist/relu1024x3; do
  name=$(echo $net | tr '/' '_')
  [ -x "$OUT/ptxt_$name" ] && [ "$OUT/ptxt_$name" -nt "$HERE/ref_harness.cpp" ] && continue
  srcs=""; for f in $LIBSRC; do srcs="$srcs $REF/lib/$f"; done
  # no -fopenmp: the OpenMP pragmas of lib/BinFunc.cpp:1056 privatise an uninitialised flag; serial build is well defined
  $CXX -O2 -w -I"$REF/lib" -I"$REF/nets/$net" -I"$REF/nets/$(dirname $net)" \
      $srcs "$REF/nets/$net/net.cpp" "$HERE/ref_harness.cpp" -o "$OUT/ptxt_$name"
  echo "built $OUT/ptxt_$name"
done
