#!/bin/bash
# usage: build_variant_file.sh FILE NAME -DFOO=1 ...   -> grace-devel_b200/build/var_NAME.so (csrc/FILE.cu rebuilt with the flags)
set -e
cd "$(dirname "$0")/../../grace-devel_b200"
FILE=$1; NAME=$2; shift 2
mkdir -p /tmp/var_$NAME
for f in csrc/*.cu; do b=$(basename $f .cu); [ $b = $FILE ] || cp build/$b.o /tmp/var_$NAME/$b.o; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../include -Icsrc --expt-relaxed-constexpr "$@" -c csrc/$FILE.cu -o /tmp/var_$NAME/$FILE.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/var_$NAME.so /tmp/var_$NAME/*.o
echo built build/var_$NAME.so
