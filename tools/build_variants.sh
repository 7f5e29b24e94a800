#!/bin/bash
# Builds kernel variants of librumdeed_b200.so into tools/variants/ (for tools/variant_bench.py).
set -e
cd "$(dirname "$0")/../rumdeed_b200/csrc"
mkdir -p ../../tools/variants
build() { # name, extra flags
  name=$1; shift
  out=../../tools/variants/$name
  mkdir -p $out
  make -s OUT=$out EXTRA="$*" >/dev/null
  echo "built $name: $(cuobjdump -res-usage $out/_obj/rb2_pair.o 2>/dev/null | grep -A1 'pairILi1ELi1ELb0' | grep -o 'REG:[0-9]*')"
}
build base
build minb5 -DRB2_MINB=5
build minb6 -DRB2_MINB=6
build minb3 -DRB2_MINB=3
build unroll1 -DRB2_UNROLL=1
build unroll4 -DRB2_UNROLL=4
build b256 -DRB2_BLOCK=256 -DRB2_MINB=2
build b64 -DRB2_BLOCK=64 -DRB2_MINB=8
build tj256 -DRB2_TJ=256
