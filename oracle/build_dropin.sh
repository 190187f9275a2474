#!/bin/sh
# Compiles the REFERENCE's own test / profiling programs, unmodified, from /root/reference/tests against
# THIS repository's drop-in headers (include/grace/**) and links them with libgrace_b200.so: the proof
# that a program written against GRACE builds and runs on the B200 path without source changes.
# Outputs go to oracle/_ref/dropin/ (git-ignored, shipped to the GPU box); tests/test_gpu_dropin.py runs
# them there.  The reference sources are read where they lie and never copied into the repository.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$HERE/..
OUT=$HERE/_ref/dropin
mkdir -p "$OUT"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -w -Xcompiler -fopenmp -I $ROOT/include -I $REF/tests"
LIBS="-L $ROOT/grace-devel_b200 -lgrace_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../grace-devel_b200 -lcurand -lcuda"
PROGS="hitcounts/hitcounts tree_traversal/tree_traversal distance_sort/distance_sort integrate/integrate \
integrate_gadget/integrate_gadget project_gadget/project_gadget profile_tree/profile_tree \
profile_tree_gadget/profile_tree_gadget profile_trace_gadget/profile_trace_gadget \
profile_project_gadget/profile_project_gadget profile_one_to_many_rays_gadget/profile_one_to_many_rays_gadget \
morton_key/30bit_key morton_key/63bit_key morton_key_kernel/30bit_keys morton_key_kernel/63bit_keys"
fail=0
for p in $PROGS; do
  name=$(echo $p | tr '/' '_')
  ( $NVCC $FLAGS "$REF/tests/$p.cu" -o "$OUT/$name" $LIBS > "$OUT/$name.log" 2>&1 && rm -f "$OUT/$name.log" ) &
done
wait
for p in $PROGS; do
  name=$(echo $p | tr '/' '_')
  if [ ! -x "$OUT/$name" ]; then echo "dropin: FAILED to build $name"; head -5 "$OUT/$name.log"; fail=1; fi
done
# The same two programs against the REFERENCE's own headers (CUDA-12 patched copy): their host loops assume a
# unit box while they pass (-1, 1)^3, so they report a mismatch with the reference's own implementation too;
# the test compares the two builds' outputs instead of expecting PASSED.
SCRATCH=${TMPDIR:-/tmp}/grace_ref_patched
[ -d "$SCRATCH/include" ] || python "$HERE/patch_ref.py" "$REF" "$SCRATCH" > /dev/null
# ... and the reference's three self-checking programs against its own headers: the gate BASELINE.md 2b puts
# before the patched reference build is used as a comparator (tests/test_gpu_dropin.py keeps their output)
for p in morton_key_kernel/30bit_keys morton_key_kernel/63bit_keys tree_traversal/tree_traversal distance_sort/distance_sort integrate/integrate; do
  name=ref_$(echo $p | tr '/' '_')
  $NVCC -arch=sm_100 -O2 -std=c++17 -w -Xcompiler -fopenmp -I "$SCRATCH/include" -I "$SCRATCH/tests" \
      -I "$SCRATCH/include/grace/external/sgpu" "$REF/tests/$p.cu" -o "$OUT/$name" -lcurand > "$OUT/$name.log" 2>&1 \
      && rm -f "$OUT/$name.log" || { echo "dropin: FAILED to build $name"; fail=1; }
done
echo "dropin: built $(ls "$OUT" | grep -vc '\.log$') programs in $OUT"
exit $fail
