#!/bin/bash
# build a variant of libfps_b200.so with extra -D flags for kdline_stream.cu / kdline_warp.cu (experiments):
#   scripts/build_variant.sh NAME "-DS_THREADS_DEF=256 -DS_PF_DEF=1"  -> fpsample_b200/variants/libfps_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; files=${3:-kdline_stream.cu}
mkdir -p fpsample_b200/variants fpsample_b200/build/var_$name
objs=""
for o in fpsample_b200/build/*.o; do
  b=$(basename $o .o)
  if [[ " $files " == *" $b.cu "* ]]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags -c fpsample_b200/csrc/$b.cu -o fpsample_b200/build/var_$name/$b.o
    objs="$objs fpsample_b200/build/var_$name/$b.o"
  else objs="$objs $o"; fi
done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o fpsample_b200/variants/libfps_$name.so $objs -Xcompiler -fPIC -ldl
echo built fpsample_b200/variants/libfps_$name.so
