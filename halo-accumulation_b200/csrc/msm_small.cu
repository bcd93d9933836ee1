// msm_small.cu -- second translation unit of msm.cu: only k_accumulate_quad and k_reduce_slabs_quad (and their launchers), with
// the field multiplication as an out-of-line call instead of ~240 inlined instructions per use.
// Why: these two kernels run few warps that are out of step with each other (one bucket or slab per quad, Poisson fills), and
// their bodies inline 20-40 multiplications: ncu on the 2^16-point MSM shows "no instruction" (instruction-cache misses) as
// their largest stall (2.9 of ~7.5 stall cycles per issue in k_accumulate_quad, issue rate 0.27 per scheduler).  With the
// multiplication out of line the code of a kernel is a few KiB: 2^13 / 2^14 / 2^16-point MSMs 0.318 / 0.399 / 0.782 -> 0.297 /
// 0.376 / 0.711 ms, open at 2^20 45.2 -> 44.5 ms (profiles/r02_fp_mul_call_small_kernels_ab.jsonl).  The large kernels keep
// the inlined multiplication: their many warps run the same code and the call costs them 1-3 %.
#define HALO_FP_MUL_CALL 1
#define HALO_MSM_SMALL_TU 1
#include "msm.cu"
