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
build m3u1 -DRB2_SYM_MINB2=3 -DRB2_SYM_UNROLL=1
build m3u2 -DRB2_SYM_MINB2=3 -DRB2_SYM_UNROLL=2
build m3u4 -DRB2_SYM_MINB2=3 -DRB2_SYM_UNROLL=4
build u8 -DRB2_SYM_UNROLL=8
