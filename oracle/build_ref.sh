#!/bin/bash
# oracle/build_ref.sh — build the part of the REFERENCE itself that this image can build: cubins of Hpt's own CUDA
# kernels, compiled from the sources where they lie under /root/reference (nothing is copied into the repository),
# outputs only into oracle/_ref/ (git-ignored; it travels to the GPU box like the other built artefacts).
#
# TEST INFRASTRUCTURE ONLY: tests/test_reference_kernels_gpu.py loads these cubins through the CUDA driver API and
# compares reference kernel / oracle / this library three ways.  Nothing under hpt_b200/ knows they exist.
#
# What builds: the reference compiles every .cu to PTX with plain nvcc (hpt-cudakernels/build.rs:233-247), so no Rust is
# needed for the device code.  With this image's nvcc 12.9 + gcc 13:
#   reduce/argmax.cu, reduce/argmin.cu, strided_copy.cu      compile as they are  → built here
#   reduce/{sum,max,mean,…}.cu, normalization/softmax.cu     do NOT compile: reduce_classes.cuh:52,194,626,703 name the
#                                                            dependent type `FloatOutBinaryPromote<T, T>::Output` without
#                                                            `typename` as a template argument (accepted by MSVC, the
#                                                            reference's CI compiler; -std=c++20 relaxes the alias
#                                                            declarations at :55 but not these)
#   binary/{add,sub,mul,rem}.cu (the NormalBinOps, §8 a1)   and binary/{div,bitand,bitor,bitxor,shl,shr,cmp}.cu (§8 f1 riders; cmp.cu alone
#                                                            is 6 × 169 × 6 kernels)
#                                                            compile once the TOOLCHAIN supplies what the sources assume:
#                                                            utils/type_cast.cuh uses __half / __nv_bfloat16 without
#                                                            including their headers → `-include cuda_fp16.h -include
#                                                            cuda_bf16.h` on the command line (a recipe flag, the sources
#                                                            stay as they lie); 169 dtype pairs × 6 kernels each, ≈ 2.5 min
#                                                            and 9 MB per op, built in parallel
#   unary/*.cu                                               do NOT compile: utils/make_vec.cuh:25-43 specialises a member
#                                                            template in class scope (MSVC-only)
# The sources are not patched (that would no longer be the reference); for those ops the oracle stays pinned by the
# promotion tables, the reference's stored test vectors and its own test oracle (oracle/hpt_oracle.py header).
# The reference kernels are also WRONG beyond ~9.7 M elements (grid-size caps): the tests stay below.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
if [ ! -d "$REF/hpt-cudakernels/src" ]; then
  echo "build_ref: $REF/hpt-cudakernels not present (GPU box): using the prebuilt oracle/_ref/*.cubin"
  exit 0
fi
mkdir -p "$OUT"
cd "$REF/hpt-cudakernels"
for f in reduce/argmax reduce/argmin strided_copy; do
  n=$(basename $f)
  if [ ! -f "$OUT/$n.cubin" ] || [ "src/$f.cu" -nt "$OUT/$n.cubin" ]; then
    TMPDIR=${TMPDIR:-/tmp} $NVCC -std=c++17 -cubin -O3 -arch=sm_100a --extended-lambda --diag-suppress=20054 -Isrc/cutlass \
      "src/$f.cu" -o "$OUT/$n.cubin"
    echo "build_ref: $OUT/$n.cubin"
  fi
done
pids=""
for n in add sub mul rem div bitand bitor bitxor shl shr cmp; do
  if [ ! -f "$OUT/binary_$n.cubin" ] || [ "src/binary/$n.cu" -nt "$OUT/binary_$n.cubin" ]; then
    ( TMPDIR=${TMPDIR:-/tmp} $NVCC -std=c++20 -cubin -O3 -arch=sm_100a --extended-lambda --diag-suppress=20054 \
        -include cuda_fp16.h -include cuda_bf16.h -Isrc/cutlass "src/binary/$n.cu" -o "$OUT/binary_$n.cubin.tmp" \
        && mv "$OUT/binary_$n.cubin.tmp" "$OUT/binary_$n.cubin" && echo "build_ref: $OUT/binary_$n.cubin" ) &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
