#!/bin/bash
# Build an experimental variant of libpncov.so into tools/build/v/lib_<name>.so:
#   tools/build_variant.sh <name> <source.cu> [-Dmacro ...]
# Only <source.cu> is recompiled (with the extra macros); the other objects come from csrc/build.
set -e
name=$1; src=$2; shift 2
cd /root/repo/probnum_b200/csrc
mkdir -p /root/repo/tools/build/v
objs=""
for f in pncov_api pncov_gram pncov_matvec pncov_extra pncov_chol; do
  if [ "$f.cu" == "$src" ]; then
    nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c $src -o /tmp/${f}_$name.o 2>&1 | grep -i "error" || true
    objs="$objs /tmp/${f}_$name.o"
  else
    objs="$objs build/$f.o"
  fi
done
nvcc -shared -o /root/repo/tools/build/v/lib_$name.so $objs -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo "built lib_$name.so"
