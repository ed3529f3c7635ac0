#!/bin/bash
# Debug aid: a variant build of the library with extra -D flags on one source (A/B on one box):  tools/ab_build.sh name file.cu -DFLAG ...
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p tools/_ab
python -c "import sys; sys.path.insert(0,'.'); from isp_tts_b200 import build; build.build()"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2 --ftz=false --prec-div=true --prec-sqrt=true --fmad=true "$@" -c isp-tts_b200/csrc/$src -o tools/_ab/$name.o
objs=$(ls isp-tts_b200/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -gencode arch=compute_100a,code=sm_100a -o tools/_ab/lib_$name.so $objs tools/_ab/$name.o
echo tools/_ab/lib_$name.so
