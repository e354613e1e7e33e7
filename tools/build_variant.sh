#!/bin/bash
# build_variant.sh <name> <nvcc -D flags...>: a second copy of the product library with other compile-time switches, for A/B timing
# (loaded with HP_B200_LIB=3d-point-clouds-autocomplete_b200/lib/variants/libhp_b200_<name>.so)
set -e
name=$1; shift
pkg=3d-point-clouds-autocomplete_b200
mkdir -p $pkg/lib/variants /tmp/hpv_$name
for f in $pkg/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    -Iinclude -I$pkg/csrc -DHP_BUILDING_LIBRARY --expt-relaxed-constexpr "$@" -c $f -o /tmp/hpv_$name/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $pkg/lib/variants/libhp_b200_$name.so /tmp/hpv_$name/*.o -cudart static
echo built $pkg/lib/variants/libhp_b200_$name.so
