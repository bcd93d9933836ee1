// msm.cu -- K2: Pippenger multi-scalar multiplication on sm_100a.
//
// Replaces `VariableBaseMSM::msm_unchecked` as called from group.rs:20 and group.rs:25 (callers:
// pedersen.rs:14, pcdl.rs:109,204,208,338, acc.rs:153,178,195).  Pipeline (all on one stream):
//   1. k_digits<COUNT>   scalars (Montgomery) -> canonical -> signed radix-2^c digits; histogram per bucket
//   2. exclusive scan    bucket offsets
//   3. k_digits<SCATTER> counting-sort scatter of (point index | sign) into bucket order
//   4. k_accumulate      one thread per bucket: XYZZ accumulator in registers, mixed adds of gathered bases
//   5. k_bucket_reduce   one CTA per window: running sums + block suffix scan -> sum_k k * B_k
//   6. host              Horner over the W window sums (255 doublings; O(1) in n) and Jacobian output
// Integer pipe bound (IMAD); see DESIGN.md for the roofline accounting.
#include "common.cuh"
#include "msm.cuh"

namespace halo {

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
MsmPlan msm_make_plan(uint64_t n, int force_c) {
    int best_c = 4;
    double best = 1e300;
    for (int c = 4; c <= 17; c++) {
        int W = 255 / c + 1;
        double M = (double)(1u << (c - 1)) * (255.0 / (W * c) > 0.97 ? 1.0 : 0.75);  // narrower windows use half their slots
        // bucket accumulation (10 modmul per mixed add) + bucket reduction (2 full adds per bucket, poorly
        // parallel -> weighted) ; tuned on B200, see profiles/
        double cost = (double)W * ((double)n * 10.0 + M * 28.0 * 6.0);
        if (cost < best) {
            best = cost;
            best_c = c;
        }
    }
    int c = force_c ? force_c : best_c;
    if (c < 4) c = 4;  // W = 255 / c + 1 <= MSM_MAX_WINDOWS
    if (c > 20) c = 20;
    MsmPlan p;
    p.c = c;
    p.W = 255 / c + 1;  // c * W >= 256 > 255: the top window absorbs the final carry
    p.M = 1u << (c - 1);
    p.NB = (uint32_t)p.W * p.M;
    // near-equal widths summing to 255; the low windows take the remainder so the top window is the narrow one
    int base = 255 / p.W, rem = 255 - base * p.W;
    for (int w = 0; w < MSM_MAX_WINDOWS; w++) p.widths.w[w] = w < p.W ? (uint8_t)(base + (w < rem ? 1 : 0)) : 0;
    return p;
}

// ------------------------------------------------------------------------------------------------
// 1/3. digit extraction: count or scatter
// ------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(const fr_t* __restrict__ scalars, uint32_t n,
                                                const fr_t* __restrict__ tail_scalars, uint32_t n_tail, const MsmWidths widths,
                                                int W, uint32_t M, uint32_t* __restrict__ counts,
                                                const uint32_t* __restrict__ offsets, uint32_t* __restrict__ entries) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + n_tail) return;
    fr_t s = i < n ? scalars[i] : tail_scalars[i - n];
    uint32_t k[8];
    fp_to_canon(k, s);  // arkworks `into_bigint`
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        const uint32_t c = widths.w[w];
        uint32_t d = (k[0] & ((1u << c) - 1u)) + carry;
#pragma unroll
        for (int j = 0; j < 7; j++) k[j] = __funnelshift_r(k[j], k[j + 1], c);
        k[7] >>= c;
        carry = 0;
        uint32_t neg = 0;
        if (d > (1u << (c - 1))) {  // digit in (-2^(c-1), 2^(c-1)]; never taken in the top window (scalar < 2^255)
            d = (1u << c) - d;
            neg = 1;
            carry = 1;
        }
        if (d != 0) {
            uint32_t b = (uint32_t)w * M + (d - 1);
            uint32_t slot = atomicAdd(&counts[b], 1u);
            if (SCATTER) entries[offsets[b] + slot] = i | (neg << 31);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 2. exclusive scan of u32 (three small kernels; NB <= 4096 * 2048)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += y;
        }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    uint32_t base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[(blockDim.x >> 5) - 1];
    return base + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_local(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                             uint32_t n, uint32_t* __restrict__ tile_sums) {
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = (base + j < n) ? in[base + j] : 0;
        sum += v[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = ex;
        ex += v[j];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// single CTA: exclusive scan of up to 2048 tile sums, also writes the grand total to out_total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(uint32_t* tile_sums, uint32_t ntiles, uint32_t* out_total) {
    constexpr int PER = 8;  // 256 * 8 = 2048 tiles
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        uint32_t idx = threadIdx.x * PER + j;
        v[j] = idx < ntiles ? tile_sums[idx] : 0;
        sum += v[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int j = 0; j < PER; j++) {
        uint32_t idx = threadIdx.x * PER + j;
        if (idx < ntiles) tile_sums[idx] = ex;
        ex += v[j];
    }
    if (threadIdx.x == 0) *out_total = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, uint32_t n,
                                                           const uint32_t* __restrict__ tile_sums) {
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) out[base + j] += add;
}

// offsets[0..n] = exclusive scan of counts[0..n) with offsets[n] = total
static void exclusive_scan(const uint32_t* counts, uint32_t* offsets, uint32_t n, uint32_t* tmp, cudaStream_t st,
                           uint64_t* launches) {
    uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_local<<<ntiles, SCAN_THREADS, 0, st>>>(counts, offsets, n, tmp);
    k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(tmp, ntiles, offsets + n);
    k_scan_add<<<ntiles, SCAN_THREADS, 0, st>>>(offsets, n, tmp);
    *launches += 3;
}

// ------------------------------------------------------------------------------------------------
// 4. bucket accumulation: one thread per bucket, accumulator in registers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_accumulate(const affine_t* __restrict__ bases, uint32_t n,
                                                    const affine_t* __restrict__ tail_bases,
                                                    const uint32_t* __restrict__ offsets,
                                                    const uint32_t* __restrict__ entries, uint32_t NB,
                                                    xyzz_t* __restrict__ buckets) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    uint32_t beg = offsets[b], end = offsets[b + 1];
    xyzz_t acc;
    xyzz_set_inf(acc);
    for (uint32_t e = beg; e < end; e++) {
        uint32_t ent = entries[e];
        uint32_t idx = ent & 0x7fffffffu;
        affine_t p = idx < n ? bases[idx] : tail_bases[idx - n];
        xyzz_madd(acc, p, (ent >> 31) != 0);
    }
    buckets[b] = acc;
}

// ------------------------------------------------------------------------------------------------
// 5. bucket reduction: one CTA per window computes S_w = sum_{k=1..M} k * B_{w,k-1}
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void xyzz_add_nl(xyzz_t& acc, const xyzz_t& q) { xyzz_add(acc, q); }
__device__ __noinline__ void xyzz_dbl_nl(xyzz_t& acc) { xyzz_dbl(acc, acc); }

constexpr int REDUCE_THREADS = 512;

__global__ void __launch_bounds__(REDUCE_THREADS) k_bucket_reduce(const xyzz_t* __restrict__ buckets, uint32_t M, int T,
                                                                  int log_s, xyzz_t* __restrict__ wsums) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    xyzz_t* sm = reinterpret_cast<xyzz_t*>(smem_raw);
    const int j = threadIdx.x;
    const uint32_t s = 1u << log_s;  // buckets per thread; T * s == M
    const xyzz_t* B = buckets + (size_t)blockIdx.x * M + (size_t)j * s;
    // running sums from the top: R = sum B_t, A = sum (t + 1) * B_t
    xyzz_t R, A;
    xyzz_set_inf(R);
    xyzz_set_inf(A);
    for (int t = (int)s - 1; t >= 0; t--) {
        xyzz_t q = B[t];
        xyzz_add_nl(R, q);
        xyzz_add_nl(A, R);
    }
    // inclusive suffix scan of R over the CTA: Suf_j = sum_{i >= j} R_i
    sm[j] = R;
    __syncthreads();
    for (int stride = 1; stride < T; stride <<= 1) {
        bool has = j + stride < T;
        xyzz_t other;
        if (has) other = sm[j + stride];
        __syncthreads();
        if (has) xyzz_add_nl(R, other);
        sm[j] = R;
        __syncthreads();
    }
    // sum_{j >= 1} Suf_j = sum_j j * R_j  -> tree sum (thread 0 contributes nothing)
    if (j == 0) xyzz_set_inf(R);
    sm[j] = R;
    __syncthreads();
    for (int stride = T >> 1; stride >= 1; stride >>= 1) {
        if (j < stride) {
            xyzz_t other = sm[j + stride];
            xyzz_add_nl(R, other);
            sm[j] = R;
        }
        __syncthreads();
    }
    // R (thread 0) = sum_j j * R_j ; scale by s = 2^log_s
    if (j == 0)
        for (int t = 0; t < log_s; t++) xyzz_dbl_nl(R);
    __syncthreads();
    // tree sum of A_j
    sm[j] = A;
    __syncthreads();
    for (int stride = T >> 1; stride >= 1; stride >>= 1) {
        if (j < stride) {
            xyzz_t other = sm[j + stride];
            xyzz_add_nl(A, other);
            sm[j] = A;
        }
        __syncthreads();
    }
    if (j == 0) {
        xyzz_add_nl(A, R);
        wsums[blockIdx.x] = A;
    }
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
static inline bool is_pow2_u32(uint32_t x) { return x && !(x & (x - 1)); }

void msm_window_sums(halo_ctx* ctx, const MsmInput& in, const MsmPlan& plan, xyzz_t* d_wsums_out) {
    const affine_t* d_bases = in.bases;
    const fr_t* d_scalars = in.scalars;
    const uint32_t n = in.n;
    const uint32_t ntot = in.n + in.n_tail;
    MsmWorkspace& ws = ctx->ws;
    cudaStream_t st = ctx->stream;
    const uint32_t NB = plan.NB;
    ws.counts.reserve((size_t)(NB + 1) * 4);
    ws.offsets.reserve((size_t)(NB + 1) * 4);
    ws.entries.reserve((size_t)ntot * plan.W * 4);
    ws.buckets.reserve((size_t)NB * sizeof(xyzz_t));
    ws.scan_tmp.reserve(4096 * 4);
    if ((NB + SCAN_TILE - 1) / SCAN_TILE > 2048) throw CudaError{cudaErrorInvalidValue, "bucket count too large for scan", __FILE__, __LINE__};

    uint32_t* counts = ws.counts.as<uint32_t>();
    uint32_t* offsets = ws.offsets.as<uint32_t>();
    uint32_t* entries = ws.entries.as<uint32_t>();
    xyzz_t* buckets = ws.buckets.as<xyzz_t>();
    const bool prof = ctx->profile;
    auto mark = [&](int i) {
        if (prof) HALO_CUDA(cudaEventRecord(ctx->ev[i], st));
    };

    mark(0);
    HALO_CUDA(cudaMemsetAsync(counts, 0, (size_t)(NB + 1) * 4, st));
    const int TPB = 256;
    uint32_t grid = (ntot + TPB - 1) / TPB;
    k_digits<false><<<grid, TPB, 0, st>>>(d_scalars, n, in.tail_scalars, in.n_tail, plan.widths, plan.W, plan.M, counts, nullptr, nullptr);
    mark(1);
    exclusive_scan(counts, offsets, NB, ws.scan_tmp.as<uint32_t>(), st, &ctx->kernel_launches);
    HALO_CUDA(cudaMemsetAsync(counts, 0, (size_t)(NB + 1) * 4, st));
    mark(2);
    k_digits<true><<<grid, TPB, 0, st>>>(d_scalars, n, in.tail_scalars, in.n_tail, plan.widths, plan.W, plan.M, counts, offsets, entries);
    mark(3);
    k_accumulate<<<(NB + 127) / 128, 128, 0, st>>>(d_bases, n, in.tail_bases, offsets, entries, NB, buckets);
    mark(4);
    int T = plan.M < (uint32_t)REDUCE_THREADS ? (int)plan.M : REDUCE_THREADS;
    int log_s = 0;
    while (((uint32_t)T << log_s) < plan.M) log_s++;
    size_t smem = (size_t)T * sizeof(xyzz_t);
    HALO_CUDA(cudaFuncSetAttribute(k_bucket_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   REDUCE_THREADS * (int)sizeof(xyzz_t)));
    k_bucket_reduce<<<plan.W, T, smem, st>>>(buckets, plan.M, T, log_s, d_wsums_out);
    mark(5);
    ctx->kernel_launches += 4;
    HALO_CUDA(cudaGetLastError());
}

// Horner over window sums (host, O(255) doublings independent of n): total = sum_w 2^(c w) S_w
void msm_finish_host(const xyzz_t* wsums, const MsmPlan& plan, xyzz_t& out) {
    xyzz_t total;
    xyzz_set_inf(total);
    for (int w = plan.W - 1; w >= 0; w--) {
        xyzz_add(total, wsums[w]);
        if (w > 0)
            for (int k = 0; k < plan.widths.w[w - 1]; k++) xyzz_dbl(total, total);
    }
    out = total;
}

// Enqueue `count` MSMs back to back on the context stream (they share the workspace, so they serialise on the
// stream), then one D2H of all window sums, one synchronisation, and the host Horner per MSM.
void msm_batch(halo_ctx* ctx, const MsmInput* ins, int count, xyzz_t* outs) {
    if (count > 4) throw CudaError{cudaErrorInvalidValue, "msm_batch: count > 4", __FILE__, __LINE__};
    MsmPlan plans[4];
    ctx->ws.wsums.reserve(4 * MSM_MAX_WINDOWS * sizeof(xyzz_t));
    xyzz_t* d_wsums = ctx->ws.wsums.as<xyzz_t>();
    if (!ctx->pinned) {
        HALO_CUDA(cudaMallocHost(&ctx->pinned, 4 * MSM_MAX_WINDOWS * sizeof(xyzz_t)));
        ctx->pinned_cap = 4 * MSM_MAX_WINDOWS * sizeof(xyzz_t);
    }
    xyzz_t* h_wsums = reinterpret_cast<xyzz_t*>(ctx->pinned);
    bool any = false;
    for (int k = 0; k < count; k++) {
        if (ins[k].n + ins[k].n_tail == 0) continue;
        plans[k] = msm_make_plan(ins[k].n + ins[k].n_tail, ctx->force_c);
        msm_window_sums(ctx, ins[k], plans[k], d_wsums + k * MSM_MAX_WINDOWS);
        HALO_CUDA(cudaMemcpyAsync(h_wsums + k * MSM_MAX_WINDOWS, d_wsums + k * MSM_MAX_WINDOWS, plans[k].W * sizeof(xyzz_t),
                                  cudaMemcpyDeviceToHost, ctx->stream));
        any = true;
    }
    if (any) HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (any && ctx->profile) {
        float t[5];
        for (int i = 0; i < 5; i++) HALO_CUDA(cudaEventElapsedTime(&t[i], ctx->ev[i], ctx->ev[i + 1]));
        ctx->last.digits_ms = t[0];
        ctx->last.scan_ms = t[1];
        ctx->last.scatter_ms = t[2];
        ctx->last.accumulate_ms = t[3];
        ctx->last.reduce_ms = t[4];
        HALO_CUDA(cudaEventElapsedTime(&ctx->last.total_ms, ctx->ev[0], ctx->ev[5]));
    }
    for (int k = 0; k < count; k++) {
        if (ins[k].n + ins[k].n_tail == 0)
            xyzz_set_inf(outs[k]);
        else
            msm_finish_host(h_wsums + k * MSM_MAX_WINDOWS, plans[k], outs[k]);
    }
}

void msm_device(halo_ctx* ctx, const affine_t* d_bases, const fr_t* d_scalars, uint64_t n, xyzz_t& out) {
    MsmInput in;
    in.bases = d_bases;
    in.scalars = d_scalars;
    in.n = (uint32_t)n;
    msm_batch(ctx, &in, 1, &out);
}

}  // namespace halo
