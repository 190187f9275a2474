#!/usr/bin/env python
"""Make a scratch copy of the reference's include tree compile with CUDA 12.9 for sm_100.

TEST INFRASTRUCTURE.  Reads /root/reference (never modified), writes the patched copy to a
scratch directory OUTSIDE the repo (argv[2]); only the compiled binaries end up under
oracle/_ref/.  The edits are the compatibility patch of SURVEY.md Appendix B: removed CUDA
APIs are replaced by their successors, arithmetic is untouched.

  1. texture references (removed in CUDA 12) -> __ldg on the node / leaf pointers the kernel
     already receives (cuda/kernels/bintree_trace.cuh:24,37-38,242-252,282-283) and a plain
     pointer for the primitives (trace_texref, :319-343)
  2. __any(x) -> __any_sync(full, x)                              (:151,156)
  3. __syncwarp() after the shared-memory staging loop and after the test loop (:178-191):
     the kernel relied on pre-Volta lock-step execution
  4. sgpu: __shfl_up / shfl.up PTX / __ballot -> *_sync forms
     (external/sgpu/device/intrinsics.cuh:116-222, ctascan.cuh:154, ctasegscan.cuh:58,63)

  patch_ref.py <reference> <scratch copy> [MAX_BLOCKS]
"""
import os
import re
import shutil
import sys

src = sys.argv[1]
dst = sys.argv[2]
if os.path.exists(dst):
    shutil.rmtree(dst)
shutil.copytree(os.path.join(src, "include"), os.path.join(dst, "include"))
os.makedirs(os.path.join(dst, "tests"), exist_ok=True)
shutil.copytree(os.path.join(src, "tests", "helper"), os.path.join(dst, "tests", "helper"))


def edit(rel, fn):
    p = os.path.join(dst, rel)
    s = open(p).read()
    t = fn(s)
    assert t != s, "patch did not apply: " + rel
    open(p, "w").write(t)


def trace(s):
    s = s.replace("#define FETCH_NODE(nodes, i) tex1Dfetch(nodes##_tex, i)",
                  "#define FETCH_NODE(nodes, i) __ldg(&nodes[i])")
    s = s.replace('#include "grace/cuda/util/texref_iter.cuh"\n', "")
    s = re.sub(r"texture<float4, cudaTextureType1D, cudaReadModeElementType> nodes_tex;\n", "", s)
    s = re.sub(r"texture<int4, cudaTextureType1D, cudaReadModeElementType> leaves_tex;\n", "", s)
    s = s.replace("__any(lr_hit & 1u)", "__any_sync(0xffffffffu, lr_hit & 1u)")
    s = s.replace("__any(lr_hit >= 2)", "__any_sync(0xffffffffu, lr_hit >= 2)")
    # warp-synchronous staging needs explicit barriers on Volta+
    s = s.replace("""                    sm_prims[max_per_leaf * wid + i] = primitives[node.x + i];
                }
""", """                    sm_prims[max_per_leaf * wid + i] = primitives[node.x + i];
                }
                __syncwarp();
""")
    s = s.replace("""                               sm_iter_usr);
                    }
                }
            }
""", """                               sm_iter_usr);
                    }
                }
                __syncwarp();
            }
""")
    # texture binds / unbinds
    s = re.sub(r"    cudaError_t cuerr;\n\n    cuerr = cudaBindTexture\(.*?GRACE_CUDA_CHECK\(cuerr\);\n\n    cuerr = cudaBindTexture\(.*?GRACE_CUDA_CHECK\(cuerr\);\n",
               "", s, flags=re.S)
    s = s.replace("    GRACE_CUDA_CHECK(cudaUnbindTexture(gpu::nodes_tex));\n", "")
    s = s.replace("    GRACE_CUDA_CHECK(cudaUnbindTexture(gpu::leaves_tex));\n", "")
    # trace_texref: forward the raw primitive pointer
    s = re.sub(r"    TexRefIter<TPrimitive, PRIMITIVE_TEX_UID> prims_iter;\n\n    cudaError_t cuerr\n        = prims_iter.bind\(d_primitives, N_primitives \* sizeof\(TPrimitive\)\);\n    GRACE_CUDA_CHECK\(cuerr\);\n\n    trace<RayData>\(d_rays_iter, N_rays, prims_iter,",
               "    trace<RayData>(d_rays_iter, N_rays, d_primitives,", s)
    s = s.replace("    GRACE_CUDA_CHECK(prims_iter.unbind());\n", "")
    return s


edit("include/grace/cuda/kernels/bintree_trace.cuh", trace)


def sgpu_intr(s):
    s = re.sub(r"__shfl_up\(([^,]+), offset, width\)", r"__shfl_up_sync(0xffffffffu, \1, offset, width)", s)
    s = s.replace('"shfl.up.b32 r0|p, %1, %2, %3;"', '"shfl.sync.up.b32 r0|p, %1, %2, %3, 0xffffffff;"')
    return s


edit("include/grace/external/sgpu/device/intrinsics.cuh", sgpu_intr)
edit("include/grace/external/sgpu/device/ctascan.cuh",
     lambda s: s.replace("__ballot(x)", "__ballot_sync(0xffffffffu, x)"))
edit("include/grace/external/sgpu/device/ctasegscan.cuh",
     lambda s: s.replace("__ballot(flag)", "__ballot_sync(0xffffffffu, flag)")
                .replace("__ballot(0 != delta_shared[tid])", "__ballot_sync(__activemask(), 0 != delta_shared[tid])"))
# Optional "tuned reference" (BASELINE.md 2b): the shipped grid cap is sized for a 7-SMX Kepler
# (kernel_config.h:11, MAX_BLOCKS = 112 = 7 x 16); argv[3] replaces it (e.g. 148 SMs x 8 resident
# 256-thread blocks) so the reference fills a B200.  Nothing else changes.
if len(sys.argv) > 3:
    mb = int(sys.argv[3])
    edit("include/grace/cuda/kernel_config.h",
         lambda s: re.sub(r"const int MAX_BLOCKS = 112;", "const int MAX_BLOCKS = %d;" % mb, s))
print("patched copy written to", dst)
