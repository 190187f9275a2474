#!/bin/bash
# Build A/B variants of libgrace_b200.so with different -D knobs into grace-devel_b200/variants/.
# usage: scripts/dev/ab_variants.sh name1:"-DX=1 -DY=2" name2:"..."
set -e
cd "$(dirname "$0")/../../grace-devel_b200"
mkdir -p variants
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  (
  mkdir -p variants/$name
  for f in csrc/*.cu; do
    b=$(basename $f .cu)
    if [ "$b" = "${AB_SRC:-trace}" ] || [ ! -f build/$b.o ]; then
      $NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../include -Icsrc --expt-relaxed-constexpr $defs -c $f -o variants/$name/$b.o
    else
      cp build/$b.o variants/$name/$b.o
    fi
  done
  $NVCC $ARCH -shared -o variants/libgrace_b200_$name.so variants/$name/*.o
  rm -rf variants/$name
  echo built variants/libgrace_b200_$name.so
  ) &
done
wait
