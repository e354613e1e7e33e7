#!/usr/bin/env bash
# Proof of the native binding: the reference's UNMODIFIED pybind glue
# utils/pytorch_structural_losses/structural_loss.cpp, compiled where it lies under /root/reference, + hp_b200_shim.cpp,
# linked against libhp_b200.so -> baseline/_ref_native/StructuralLossesBackend<ext>.so (git-ignored, shipped by gpurun).
# No reference source is copied; where /root/reference is absent (the GPU box) the prebuilt module is used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${HP_REFERENCE_ROOT:-/root/reference}/utils/pytorch_structural_losses"
OUT="$REPO/baseline/_ref_native"
LIBDIR="$REPO/3d-point-clouds-autocomplete_b200/lib"
if [ ! -f "$REF/structural_loss.cpp" ]; then
  echo "[build_shim] $REF/structural_loss.cpp not present (GPU box?) - using the prebuilt module in $OUT if any"
  exit 0
fi
mkdir -p "$OUT/obj"
PY="${PYTHON:-python}"
read -r TORCH_INC TORCH_LIB PY_INC EXT_SUFFIX CXX11 <<<"$($PY - <<'PYEOF'
import os, sysconfig, torch
d = os.path.dirname(torch.__file__)
print(os.path.join(d, "include"), os.path.join(d, "lib"), sysconfig.get_paths()["include"],
      sysconfig.get_config_var("EXT_SUFFIX"), int(torch._C._GLIBCXX_USE_CXX11_ABI))
PYEOF
)"
SO="$OUT/StructuralLossesBackend${EXT_SUFFIX}"
if [ -f "$SO" ] && [ "$SO" -nt "$REF/structural_loss.cpp" ] && [ "$SO" -nt "$HERE/hp_b200_shim.cpp" ] && [ "$SO" -nt "$REPO/include/hp_b200.h" ]; then
  echo "[build_shim] up to date: $SO"; exit 0
fi
COMMON=(-I"$TORCH_INC" -I"$TORCH_INC/torch/csrc/api/include" -I"$PY_INC" -I/usr/local/cuda/include -I"$REPO/include"
        -DTORCH_EXTENSION_NAME=StructuralLossesBackend -DTORCH_API_INCLUDE_EXTENSION_H
        -D_GLIBCXX_USE_CXX11_ABI="$CXX11" -std=c++17 -O2 -fPIC -w)
g++ "${COMMON[@]}" -c "$REF/structural_loss.cpp" -o "$OUT/obj/structural_loss.o" &
g++ "${COMMON[@]}" -c "$HERE/hp_b200_shim.cpp" -o "$OUT/obj/hp_b200_shim.o" &
wait
# rpath relative to the module, so the pair keeps working wherever the repository snapshot lands
g++ -shared -o "$SO" "$OUT/obj/structural_loss.o" "$OUT/obj/hp_b200_shim.o" \
    -L"$LIBDIR" -lhp_b200 -Wl,-rpath,'$ORIGIN/../../3d-point-clouds-autocomplete_b200/lib' \
    -L"$TORCH_LIB" -Wl,-rpath,"$TORCH_LIB" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python \
    -L/usr/local/cuda/lib64 -lcudart
echo "[build_shim] built $SO"
