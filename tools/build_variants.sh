#!/bin/bash
# Builds kernel variants of librumdeed_b200.so into tools/variants/ (for tools/variant_bench.py).
set -e
cd "$(dirname "$0")/../rumdeed_b200/csrc"
rm -rf ../../tools/variants
mkdir -p ../../tools/variants
build() { # name, extra flags
  name=$1; shift
  out=../../tools/variants/$name
  mkdir -p $out
  make -s OUT=$out EXTRA="$*" >/dev/null
  echo "built $name: $(cuobjdump -res-usage $out/_obj/rb2_pair_sym.o 2>/dev/null | grep -A1 'pair_symILi1ELi2' | grep -o 'REG:[0-9]*')"
}
build base
# round 2: cost of the close-pair flag (one compare per pair) -- integer compare of the high word (base), FP64 compare, none; slow path inlined
build closefp -DRB2_CLOSE_INT=0
build exinline -DRB2_EXACT_INLINE=1
build closeoff -DRB2_CLOSE_OFF
