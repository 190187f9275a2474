#!/bin/sh
# Builds the reference-derived checkers into oracle/_ref/ (git-ignored, shipped to the GPU box
# with the snapshot).  Sources stay under /root/reference; the CUDA-12 compatibility patch is
# applied to a scratch copy outside the repository.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
SCRATCH=${TMPDIR:-/tmp}/grace_ref_patched
mkdir -p "$OUT"
python "$HERE/patch_ref.py" "$REF" "$SCRATCH" > /dev/null
# 1. the reference's CUDA implementation behind our driver
/usr/local/cuda/bin/nvcc -arch=sm_100 -O3 -std=c++17 -w -Xcompiler -fopenmp \
    -I "$SCRATCH/include" -I "$SCRATCH/tests" -I "$SCRATCH/include/grace/external/sgpu" \
    "$HERE/ref_driver.cu" -o "$OUT/ref_driver" -lcurand
/usr/local/cuda/bin/nvcc -arch=sm_100 -O3 -std=c++17 -w -Xcompiler -fopenmp \
    -I "$SCRATCH/include" -I "$SCRATCH/tests" -I "$SCRATCH/include/grace/external/sgpu" \
    "$HERE/ref_gen_driver.cu" -o "$OUT/ref_gen_driver" -lcurand
/usr/local/cuda/bin/nvcc -arch=sm_100 -O3 -std=c++17 -w -I "$REF/include" -I "$REF/tests" \
    "$HERE/ref_gadget_driver.cu" -o "$OUT/ref_gadget_driver"
/usr/local/cuda/bin/nvcc -arch=sm_100 -O3 -std=c++17 -w -Xcompiler -fopenmp \
    -I "$SCRATCH/include" -I "$SCRATCH/tests" -I "$SCRATCH/include/grace/external/sgpu" -I "$HERE/../tests/cpp" \
    "$HERE/ref_generic_driver.cu" -o "$OUT/ref_generic_driver" -lcurand
/usr/local/cuda/bin/nvcc -arch=sm_100 -O3 -std=c++17 -w -Xcompiler -fopenmp \
    -I "$SCRATCH/include" -I "$SCRATCH/tests" -I "$SCRATCH/include/grace/external/sgpu" -I "$HERE/../tests/cpp" \
    "$HERE/ref_sph_double_driver.cu" -o "$OUT/ref_sph_double_driver" -lcurand
# 1b. the bench workload through the reference's API, as shipped and with the grid cap lifted to fill a
#     B200 ("tuned reference").  The synthetic-snapshot generator is linked as object
#     files of the product library (workload generation only; libgrace_b200.so is not loaded).
GEN_OBJS="$HERE/../grace-devel_b200/build/synth.o $HERE/../grace-devel_b200/build/context.o $HERE/../grace-devel_b200/build/radix_sort.o"
if [ -f "$HERE/../grace-devel_b200/build/synth.o" ]; then
  /usr/local/cuda/bin/nvcc -arch=sm_100 -O3 -std=c++17 -w -Xcompiler -fopenmp \
      -I "$SCRATCH/include" -I "$SCRATCH/tests" -I "$SCRATCH/include/grace/external/sgpu" -I "$HERE/../include" \
      "$HERE/ref_bench.cu" $GEN_OBJS -o "$OUT/ref_bench" -lcurand
  TUNED=${TMPDIR:-/tmp}/grace_ref_patched_tuned
  python "$HERE/patch_ref.py" "$REF" "$TUNED" $((148 * 8)) > /dev/null
  /usr/local/cuda/bin/nvcc -arch=sm_100 -O3 -std=c++17 -w -Xcompiler -fopenmp \
      -I "$TUNED/include" -I "$TUNED/tests" -I "$TUNED/include/grace/external/sgpu" -I "$HERE/../include" \
      "$HERE/ref_bench.cu" $GEN_OBJS -o "$OUT/ref_bench_tuned" -lcurand
else
  echo "note: build grace-devel_b200 first (ref_bench needs build/synth.o)"
fi
# 2. the reference's host-callable code (unpatched headers)
/usr/bin/g++ -O3 -fPIC -shared -fopenmp -ffp-contract=off -fvisibility=hidden \
    -I "$REF/include" -I /usr/local/cuda/include "$HERE/ref_cpu.cpp" -o "$OUT/libgrace_ref_cpu.so"
# 3. chealpix (pix2vec_nest) for the HEALPix generator
/usr/bin/gcc -O2 -fPIC -shared -I "$REF/RayVectorGeneration/src/chealpix" \
    "$REF/RayVectorGeneration/src/chealpix/chealpix.c" -o "$OUT/libchealpix.so" -lm
echo "built: $(ls "$OUT")"
