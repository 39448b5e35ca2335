#!/bin/bash
# Debug build of the library with the globaltimer stamps compiled in (-DVDB_TIMELINE) -> tools/bin/libvdb200_tl.so
# (git-ignored, but it travels to the GPU box).  Use it with  VDB200_LIB=$PWD/tools/bin/libvdb200_tl.so python tools/*_timeline.py
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin /tmp/vdb_tlbuild
for f in host_util igemm attention elementwise; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -DVDB_TIMELINE \
       -c versatile-diffusion_b200/csrc/$f.cu -o /tmp/vdb_tlbuild/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/libvdb200_tl.so /tmp/vdb_tlbuild/{host_util,igemm,attention,elementwise}.o
ls -la tools/bin/libvdb200_tl.so
