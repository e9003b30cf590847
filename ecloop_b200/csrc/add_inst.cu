// add_inst.cu — one instance of the fused add kernel per translation unit so that the twelve of them compile in
// parallel: -DADD_VARIANT = (A33 ? 1 : 0) | (A65 ? 2 : 0) | (ENDO ? 4 : 0) (1, 2, 3, 5, 6, 7: ecl_api.cu pick_add_kernel),
// each with -DADD_HBM=0 (filter in shared memory, inline probes) and -DADD_HBM=1 (filter in HBM, probe pipe).
#include <cuda_runtime.h>

// Branch-free field arithmetic in the pipelined instances (fp.cuh fe_fix_branchfree): their point is that hash and field
// work share basic blocks. The two-phase instances keep the branchy form (4 instructions instead of 17 per product).
#ifndef ECL_FE_BRANCHFREE
#if defined(ADD_HBM) && ADD_HBM
#define ECL_FE_BRANCHFREE_WANT (((ECL_SP_BF_HBM_DEFAULT) >> ADD_VARIANT) & 1)
#else
#define ECL_FE_BRANCHFREE_WANT (((ECL_SP_BF_DEFAULT) >> ADD_VARIANT) & 1)
#endif
#ifndef ECL_SP_BF_DEFAULT
#define ECL_SP_BF_DEFAULT 0x06
#endif
#ifndef ECL_SP_BF_HBM_DEFAULT
#define ECL_SP_BF_HBM_DEFAULT 0x02
#endif
#if ECL_FE_BRANCHFREE_WANT
#define ECL_FE_BRANCHFREE 1
#else
#define ECL_FE_BRANCHFREE 0
#endif
#endif

#include "add_kernel.cuh"

// Which instances run the software-pipelined kernel (bit v = ADD_VARIANT v; only variants without the endomorphism
// exist in that form), and which two-phase instances hash one point at a time (NW = 1: half the loop body).
// The choice per variant is a measurement (profiles/r02_*_add_variants.txt), not a principle.
// Round-2 measurements (profiles/r02_a_add_variants.txt, M base keys/s, filter in shared memory | 4 GiB filter in HBM):
//   variant 2 (addr65)        two-phase NW=2 4077 | 3248   NW=1 4300 | 3956   pipelined 4411 | 3138
//   variant 3 (both)          two-phase NW=2 2767 | 2206   NW=1 2976 | ....   pipelined 2888 | 1260
//   variant 5 (addr33 + endo) two-phase NW=2 1300 | 1225   NW=1 1307 | ....
//   variant 7 (both + endo)   two-phase NW=2  557 |  428   NW=1  561 | ....
// A loop body beyond ~128 KB (NW=2 with two hash kinds, the pipelined form with both) streams its instructions from an
// L2 that the random probe traffic keeps busy: with a filter in HBM the smaller body wins by up to 20 %.
#if ADD_HBM
#ifndef ECL_SP_MASK_HBM
#define ECL_SP_MASK_HBM 0x02  // variant 1
#endif
#ifndef ECL_NW1_MASK_HBM
#define ECL_NW1_MASK_HBM 0xCC  // variants 2, 3, 6, 7 (variant 5: NW=2 1225, NW=1 1181)
#endif
#define ECL_SP_MASK_ ECL_SP_MASK_HBM
#define ECL_NW1_MASK_ ECL_NW1_MASK_HBM
#else
#ifndef ECL_SP_MASK
#define ECL_SP_MASK 0x06  // variants 1 (addr33) and 2 (addr65)
#endif
#ifndef ECL_NW1_MASK
#define ECL_NW1_MASK 0xE8  // variants 3, 5, 6, 7
#endif
#define ECL_SP_MASK_ ECL_SP_MASK
#define ECL_NW1_MASK_ ECL_NW1_MASK
#endif

#ifndef ADD_VARIANT
#error "compile with -DADD_VARIANT=<flags 1..3 or 5..7>"
#endif
#define V_A33 ((ADD_VARIANT & 1) != 0)
#define V_A65 ((ADD_VARIANT & 2) != 0)
#define V_ENDO ((ADD_VARIANT & 4) != 0)
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

// -DADD_HBM=1 builds the instance for filters in HBM (asynchronous probe pipe, probe_pipe.cuh)
#ifndef ADD_HBM
#define ADD_HBM 0
#endif
#if ADD_HBM
#define LAUNCH_NAME CAT(ecl_add_launch_hbm_, ADD_VARIANT)
#else
#define LAUNCH_NAME CAT(ecl_add_launch_, ADD_VARIANT)
#endif

cudaError_t LAUNCH_NAME(const AddParams &p, unsigned grid, unsigned smem, cudaStream_t stream) {
#if ((ECL_SP_MASK_ >> ADD_VARIANT) & 1) && !V_ENDO
  auto fn = add_kernel_sp<ADD_H, V_A33, V_A65, ADD_HBM != 0>;
#elif (ECL_NW1_MASK_ >> ADD_VARIANT) & 1
  auto fn = add_kernel<ADD_H, V_A33, V_A65, V_ENDO, ADD_HBM != 0, 1>;
#else
  auto fn = add_kernel<ADD_H, V_A33, V_A65, V_ENDO, ADD_HBM != 0, 2>;
#endif
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fn<<<grid, ADD_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}
