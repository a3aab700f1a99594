#!/bin/bash
# Another build of the library with extra compiler flags, next to the product library:
#   tools/build_variant_lib.sh NAME "-DFLAG ..."  ->  mssvt_b200/libmssvt_b200_NAME.so
# (-DMSSVT_TRACE: per-phase clock64 timelines printed by the FFN / tile kernels).  Select it with
# MSSVT_B200_LIB=mssvt_b200/libmssvt_b200_NAME.so (tools/kernel_times.py, tools/profile_forward.py).
set -e
name="$1"; extra="$2"
cd "$(dirname "$0")/../mssvt_b200/csrc"
make -j8 OBJD="$PWD/build_$name" OUT="$PWD/../libmssvt_b200_$name.so" \
  FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=default -cudart static $extra"
