// msm.cuh -- internal interface of the MSM engine (msm.cu).
#pragma once
#include <functional>

#include "common.cuh"

namespace halo {
// One MSM problem, all pointers device resident: sum_{i<n} scalars[i] * bases[i] plus an optional short "tail"
// of extra (base, scalar) pairs held elsewhere in memory (used for the H' term of the IPA rounds, pcdl.rs:204,208).
struct MsmInput {
    const affine_t* bases = nullptr;
    const fr_t* scalars = nullptr;
    uint32_t n = 0;
    const affine_t* tail_bases = nullptr;
    const fr_t* tail_scalars = nullptr;
    uint32_t n_tail = 0;
    // FIXED-base mode: bases = table of precomputed multiples, table[w * fixed_stride + fixed_first + i]
    uint32_t fixed_stride = 0;
    uint32_t fixed_first = 0;
};
MsmPlan msm_make_fixed_plan(uint64_t n, int force_c);
// Counting sort of an MSM ahead of its accumulation: own stream (high priority), own sort buffers, throttled grid.
struct SortAhead {
    MsmWorkspace* ws;
    cudaStream_t stream;
    cudaEvent_t sorted;
    int ctas_per_sm;  // 0: full grid
};
// Device part of one MSM: 3 partial points per window into d_out (see msm.cu).
// (`plan` is the caller's copy: the reduction geometry in it is settled here and read back by msm_finish_host)
void msm_enqueue(halo_ctx* ctx, const MsmInput& in, MsmPlan& plan, xyzz_t* d_out, int lane = 0, const SortAhead* ahead = nullptr);
// First `passes` levels of the bucket sums as flat pairwise affine additions with batched inversion (msm_pairs.cu).
const affine_t* pair_tree_enqueue(halo_ctx* ctx, MsmWorkspace& ws, cudaStream_t st, const MsmInput& in, const uint32_t* entries,
                                  const uint32_t* total_slots, uint64_t slots_max, int passes);
// vals[i] <- 1 / vals[i] for n non-zero elements of Fq, in place (Montgomery's trick over a product hierarchy, msm_pairs.cu)
void batch_invert(halo_ctx* ctx, cudaStream_t st, fq_t* vals, uint32_t n, DevBuf& scratch);
// Builds the table of precomputed multiples for the resident generators (FIXED-base mode).
void msm_precompute_tables(halo_ctx* ctx, int force_c);
// MSM over resident generators G_first.., FIXED-base when available.
void msm_gens_device(halo_ctx* ctx, const fr_t* d_scalars, uint64_t first, uint64_t n, xyzz_t& out);
// Up to 4 MSMs enqueued back to back with a single synchronisation; results on the host.
void msm_batch(halo_ctx* ctx, const MsmInput* ins, int count, xyzz_t* outs, const std::function<void()>* while_running = nullptr);
void msm_finish_host(const xyzz_t* wsums, const MsmPlan& plan, xyzz_t& out);
// Full MSM with device-resident inputs; synchronises the context stream and returns the point on the host.
void msm_device(halo_ctx* ctx, const affine_t* d_bases, const fr_t* d_scalars, uint64_t n, xyzz_t& out);
// Kernels of the second translation unit of msm.cu (msm_small.cu: out-of-line field multiplication): four (or two) lanes per
// bucket for the accumulation of small MSMs, and the quad-cooperative bucket reduction.
void launch_accumulate_quad(cudaStream_t st, int lanes, int minb, const affine_t* bases, uint32_t n, const affine_t* tail_bases,
                            const uint32_t* offsets, const uint32_t* entries, uint32_t NB, xyzz_t* buckets, uint32_t split_len);
void launch_reduce_slabs_quad(cudaStream_t st, dim3 grid, int T, int log_s, const xyzz_t* in, const xyzz_t* extra, size_t in_stride,
                              xyzz_t* outA, xyzz_t* outR, xyzz_t* outE, int out_stride);
}  // namespace halo
