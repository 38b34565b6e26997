#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/ from the reference sources WHERE THEY LIE
# (/root/reference, read-only).  Outputs (.so only) go to the git-ignored oracle/_ref/, which
# still travels to the GPU box with gpurun.  No reference source is copied into the repo: the
# line-range extracts are temporary files deleted at the end of this script.
#
# Why an extract and not `nvcc <reference file>`: the reference .cu includes THC/THC.h and uses
# THCudaMalloc/THCCeilDiv/AT_CHECK (removed from torch >= 1.11), so the file as a whole cannot be
# compiled against torch 2.11.  Its device code (:19-308) is self-contained CUDA C.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}"
SRC="$REF/utils/nms/src/rotate_polygon_nms_kernel.cu"
OUT="$HERE/_ref"
if [ ! -f "$SRC" ]; then
  echo "build_ref.sh: $SRC not present (GPU box?) -- using prebuilt oracle/_ref if any" >&2
  exit 0
fi
mkdir -p "$OUT"
# device helpers + kernel: from '#define DIVUP' up to (not including) 'void _set_device'
awk '/^#define DIVUP/{on=1} /^void _set_device/{on=0} on' "$SRC" > "$OUT/_ref_device_cuda.inc"
# device helpers only (no kernel): stop before '__global__ void rotate_nms_kernel'
awk '/^#define DIVUP/{on=1} /^__global__ void rotate_nms_kernel/{on=0} on' "$SRC" > "$OUT/_ref_device_host.inc"
trap 'rm -f "$OUT/_ref_device_cuda.inc" "$OUT/_ref_device_host.inc"' EXIT

# (1) host build: no FMA contraction (x86 baseline), OpenMP for the timing baseline
g++ -O2 -fPIC -shared -fopenmp -ffp-contract=off -std=c++17 \
    -DREF_EXTRACT_HOST="\"$OUT/_ref_device_host.inc\"" \
    "$HERE/ref_host_wrap.cpp" -o "$OUT/libref_rnms_host.so"
# (1b) host build WITH fma contraction (sensitivity probe)
g++ -O2 -fPIC -shared -fopenmp -mfma -ffp-contract=fast -std=c++17 \
    -DREF_EXTRACT_HOST="\"$OUT/_ref_device_host.inc\"" \
    "$HERE/ref_host_wrap.cpp" -o "$OUT/libref_rnms_host_fma.so"
# (2) CUDA build: nvcc DEFAULT math flags like the reference's setup.py; only the arch is added
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
    -DREF_EXTRACT_CUDA="\"$OUT/_ref_device_cuda.inc\"" \
    "$HERE/ref_cuda_wrap.cu" -o "$OUT/libref_rnms_cuda.so"
# keep SASS/PTX of the reference kernel for the arithmetic-pinning study
"$NVCC" -gencode arch=compute_100a,code=sm_100a -cubin \
    -DREF_EXTRACT_CUDA="\"$OUT/_ref_device_cuda.inc\"" \
    "$HERE/ref_cuda_wrap.cu" -o "$OUT/ref_rnms_sm100a.cubin"
echo "built: $(ls "$OUT")"
