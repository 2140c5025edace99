#!/bin/bash
# build_variant.sh NAME FILE.cu [-D...]: tools/ab/liblgx_NAME.so = the current objects with FILE.cu recompiled with the extra flags
set -e
cd "$(dirname "$0")/../cylinder-pose-estimation_b200"
name=$1; src=$2; shift 2
tmp=$(mktemp -d)
cp build/*.o $tmp/
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -I ../include -I csrc "$@" -c csrc/$src -o $tmp/${src%.cu}.o 2>/dev/null
mkdir -p ../tools/ab
/usr/local/cuda/bin/nvcc -shared -o ../tools/ab/liblgx_$name.so $tmp/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf $tmp
echo built tools/ab/liblgx_$name.so
