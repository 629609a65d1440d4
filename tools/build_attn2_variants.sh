#!/bin/bash
# Builds side-loadable variants of the library that differ only in the -D switches of vsum_attn2_tc05.cu
# (tools/variants/libvsum_<name>.so; git-ignored, they travel to the GPU box with the snapshot).
#   tools/build_attn2_variants.sh name1 "-DFOO=1 -DBAR=2" name2 "-D..." ...
set -e
cd "$(dirname "$0")/../video-summarization_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/variants /tmp/a2v
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../../include \
       $flags -c vsum_attn2_tc05.cu -o /tmp/a2v/$name.o
  objs=$(ls build/*.o | grep -v vsum_attn2_tc05.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/variants/libvsum_$name.so $objs /tmp/a2v/$name.o
  echo "built tools/variants/libvsum_$name.so ($flags)"
done
