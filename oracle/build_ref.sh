#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/): builds the UNMODIFIED reference CUDA extension
# (utils/pytorch_structural_losses/{structural_loss.cpp,nndistance.cu,approxmatch.cu})
# from the sources where they lie under /root/reference into oracle/_ref/.
# Outputs only into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).
# No reference source is copied into this repo. The resulting module is the parity
# oracle for nn_distance / approx_match / match_cost on the GPU (tests/ -m gpu) and a
# secondary comparator in bench.py; it is never on the product path.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${HP_REFERENCE_ROOT:-/root/reference}/utils/pytorch_structural_losses"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "[oracle/build_ref] $REF not present (GPU box?) - using prebuilt files in $OUT if any"
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
if [ -f "$SO" ] && [ "$SO" -nt "$REF/approxmatch.cu" ] && [ "$SO" -nt "$REF/nndistance.cu" ] && [ "$SO" -nt "$REF/structural_loss.cpp" ]; then
  echo "[oracle/build_ref] up to date: $SO"; exit 0
fi
COMMON=(-I"$TORCH_INC" -I"$TORCH_INC/torch/csrc/api/include" -I"$PY_INC"
        -DTORCH_EXTENSION_NAME=StructuralLossesBackend -DTORCH_API_INCLUDE_EXTENSION_H
        -D_GLIBCXX_USE_CXX11_ABI="$CXX11" -std=c++17 -O3)
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# same default flags torch's BuildExtension passes; arch = plain sm_100 (the reference has no arch flags)
NVFLAGS=(-gencode arch=compute_100,code=sm_100 --expt-relaxed-constexpr -Xcompiler -fPIC -w
         -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__)
"$NVCC" "${COMMON[@]}" "${NVFLAGS[@]}" -c "$REF/nndistance.cu"  -o "$OUT/obj/nndistance.o" &
"$NVCC" "${COMMON[@]}" "${NVFLAGS[@]}" -c "$REF/approxmatch.cu" -o "$OUT/obj/approxmatch.o" &
g++ "${COMMON[@]}" -fPIC -w -I/usr/local/cuda/include -c "$REF/structural_loss.cpp" -o "$OUT/obj/structural_loss.o" &
wait
g++ -shared -o "$SO" "$OUT/obj/structural_loss.o" "$OUT/obj/nndistance.o" "$OUT/obj/approxmatch.o" \
    -L"$TORCH_LIB" -Wl,-rpath,"$TORCH_LIB" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python \
    -L/usr/local/cuda/lib64 -lcudart
echo "[oracle/build_ref] built $SO"
