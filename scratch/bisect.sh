#!/bin/bash
# usage: bisect.sh "<extra nvcc flags>"  -> builds lib variant into scratch/lib_$tag.so
set -e
cd /root/repo/dysgu_b200/csrc
for f in engine sw ed; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --diag-suppress 177 $1 -c $f.cu -o /tmp/bis_$f.o & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -cudart static -o /root/repo/scratch/lib_$2.so /tmp/bis_engine.o /tmp/bis_sw.o /tmp/bis_ed.o
