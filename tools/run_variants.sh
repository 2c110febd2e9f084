#!/bin/bash
# On the GPU box: run a command once per variant library in tools/build/v (see build_variant.sh)
#   tools/run_variants.sh "a b c" python tools/prof_driver.py gram 32768
names=$1; shift
cp probnum_b200/csrc/libpncov.so /tmp/libpncov_main.so
for n in $names; do
  echo "== variant $n"
  cp tools/build/v/lib_$n.so probnum_b200/csrc/libpncov.so
  "$@" 2>&1 | tail -${TAIL:-4}
done
cp /tmp/libpncov_main.so probnum_b200/csrc/libpncov.so
