// capi.cu -- extern "C" entry points of libhalo_b200.so (include/halo_b200.h).
#include <cstdio>
#include <cstring>
#include <system_error>
#include <thread>

#include "../../include/halo_b200.h"
#include "common.cuh"
#include "msm.cuh"
#include "params.cuh"
#include "vec.cuh"

using namespace halo;

namespace halo {
// Registry behind DevBuf's canaries: one process-wide list, guarded by a mutex (a context is single-threaded like the
// reference, but several contexts may live on several host threads).
std::vector<DevBuf*>& devbuf_registry() {
    static std::vector<DevBuf*> r;
    return r;
}
std::mutex& devbuf_registry_mutex() {
    static std::mutex m;
    return m;
}

__global__ void __launch_bounds__(256) k_mark_infinity(affine_t* __restrict__ bases, const uint8_t* __restrict__ inf, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && inf[i]) affine_set_inf(bases[i]);
}

__global__ void __launch_bounds__(128) k_jac_to_affine(const jac_t* __restrict__ in, affine_t* __restrict__ out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    jac_t p = in[i];
    xyzz_t q;
    jac_to_xyzz(q, p);
    affine_t a;
    xyzz_to_affine(a, q);
    out[i] = a;
}

int set_error(halo_ctx* ctx, int code, const char* fmt, const char* a, const char* b, int line) {
    if (ctx) {
        char buf[512];
        snprintf(buf, sizeof buf, fmt, a, b, line);
        ctx->last_error = buf;
    }
    return code;
}

}  // namespace halo

namespace halo {
void h2d_copy(halo_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return;
    constexpr size_t CH = halo_ctx::STAGE_CHUNK;
    bool pageable = false;
    if (ctx->tune_stage_pageable && bytes >= 4 * CH) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess)
            pageable = at.type == cudaMemoryTypeUnregistered;
        else
            cudaGetLastError();
    }
    if (!pageable) {
        HALO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return;
    }
    constexpr int TMAX = halo_ctx::STAGE_THREADS, S = halo_ctx::STAGE_SLOTS;
    const int T = ctx->tune_stage_threads < 1 ? 1 : ctx->tune_stage_threads > TMAX ? TMAX : ctx->tune_stage_threads;
    for (int i = 0; i < T * S; i++)
        if (!ctx->stage_pinned[i]) {
            HALO_CUDA(cudaMallocHost(&ctx->stage_pinned[i], CH));
            HALO_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
        }
    const size_t nchunks = (bytes + CH - 1) / CH;
    cudaError_t errs[TMAX];
    auto work = [&](int t) {
        errs[t] = cudaSetDevice(ctx->device);
        size_t use = 0;
        for (size_t k = (size_t)t; k < nchunks && errs[t] == cudaSuccess; k += T, use++) {
            const int slot = t * S + (int)(use % S);
            if (use >= (size_t)S) errs[t] = cudaEventSynchronize(ctx->stage_ev[slot]);  // the DMA out of this slot two chunks ago
            const size_t off = k * CH, len = bytes - off < CH ? bytes - off : CH;
            memcpy(ctx->stage_pinned[slot], static_cast<const char*>(src) + off, len);
            if (errs[t] == cudaSuccess)
                errs[t] = cudaMemcpyAsync(static_cast<char*>(dst) + off, ctx->stage_pinned[slot], len, cudaMemcpyHostToDevice, st);
            if (errs[t] == cudaSuccess) errs[t] = cudaEventRecord(ctx->stage_ev[slot], st);
        }
    };
    // the ring may still be draining from the previous staged copy (a different stream): wait before refilling it
    for (int i = 0; i < TMAX * S; i++)
        if (ctx->stage_ev[i]) HALO_CUDA(cudaEventSynchronize(ctx->stage_ev[i]));
    std::thread th[TMAX];
    int spawned = 0;
    for (int t = 1; t < T; t++) {
        try {
            th[t] = std::thread(work, t);
            spawned |= 1 << t;
        } catch (const std::system_error&) {
        }
    }
    work(0);
    for (int t = 1; t < T; t++) {
        if (spawned & (1 << t))
            th[t].join();
        else
            work(t);  // no thread available: this thread takes the chunks as well
    }
    for (int t = 0; t < T; t++)
        if (errs[t] != cudaSuccess) throw CudaError{errs[t], "staged host-to-device copy", __FILE__, __LINE__};
}

void async_init(halo_ctx* ctx) {
    if (ctx->copy_stream) return;
    HALO_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    int prio_lo = 0, prio_hi = 0;
    HALO_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    HALO_CUDA(cudaStreamCreateWithPriority(&ctx->sort_stream, cudaStreamNonBlocking, prio_hi));
    for (auto& s : ctx->slots) {
        HALO_CUDA(cudaEventCreateWithFlags(&s.sorted, cudaEventDisableTiming));
        HALO_CUDA(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        HALO_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        HALO_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s.h_parts), 3 * MSM_MAX_WINDOWS * sizeof(xyzz_t)));
    }
}
}  // namespace halo

#define HALO_TRY(ctx) \
    try {             \
        HALO_CUDA(cudaSetDevice((ctx)->device));
#define HALO_CATCH(ctx)                                                                                        \
    }                                                                                                          \
    catch (const halo::CudaError& e) {                                                                         \
        return halo::set_error(ctx, HALO_ECUDA, "CUDA error: %s at %s:%d", cudaGetErrorString(e.err), e.file, e.line); \
    }                                                                                                          \
    catch (const std::bad_alloc&) {                                                                            \
        return halo::set_error(ctx, HALO_ENOMEM, "out of host memory%s%s%d", "", "", 0);                        \
    }                                                                                                          \
    catch (...) { /* nothing unwinds across the C ABI */                                                       \
        return halo::set_error(ctx, HALO_ECUDA, "unexpected internal exception%s%s%d", "", "", 0);              \
    }                                                                                                          \
    return HALO_OK;

static int fail(halo_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}

static void out_jac_from_xyzz(const xyzz_t& p, uint64_t out[12]) {
    jac_t j;
    xyzz_to_jac(j, p);
    memcpy(out, &j, 96);
}

extern "C" {

int halo_ctx_create(int device, uint64_t max_n, halo_ctx** out) {
    if (!out) return HALO_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return HALO_ECUDA;
    halo_ctx* ctx = new halo_ctx();
    ctx->device = device;
    ctx->max_n = max_n;
    try {
        HALO_CUDA(cudaSetDevice(device));
        HALO_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        HALO_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
        {  // staging threads for pageable host buffers: the host's hardware threads shared among the visible GPUs, 4 .. 8
            int ndev = 1;
            if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) ndev = 1;
            const int hw = (int)std::thread::hardware_concurrency();
            const int t = hw / ndev;
            ctx->tune_stage_threads = t < 4 ? 4 : t > halo_ctx::STAGE_THREADS ? halo_ctx::STAGE_THREADS : t;
        }
        HALO_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        HALO_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        HALO_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        for (auto& e : ctx->ev) HALO_CUDA(cudaEventCreate(&e));
    } catch (const halo::CudaError&) {
        delete ctx;
        return HALO_ECUDA;
    }
    *out = ctx;
    return HALO_OK;
}

void halo_ctx_destroy(halo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->gens.release();
    ctx->fixed_table.release();
    ctx->gens_pre.release();
    ctx->stage_scalars.release();
    ctx->stage_bases.release();
    ctx->stage_misc.release();
    ctx->poly_dev.release();
    for (halo::DevBuf* b : {&ctx->ipa_G, &ctx->ipa_cs, &ctx->ipa_zs, &ctx->ipa_pbar, &ctx->ipa_tail, &ctx->ipa_frozen,
                           &ctx->ipa_sums, &ctx->ipa_den, &ctx->ipa_inv_scratch, &ctx->ipa_bx, &ctx->ipa_diff, &ctx->ipa_den2, &ctx->ipa_ops}) b->release();
    for (MsmWorkspace* wsp : {&ctx->ws, &ctx->ws2}) {
    MsmWorkspace& ws = *wsp;
    for (DevBuf* b : {&ws.counts, &ws.offsets, &ws.cursor, &ws.entries, &ws.buckets, &ws.wsums, &ws.scan_tmp,
                      &ws.task_bucket, &ws.task_partial, &ws.split_ctrl, &ws.split_tasks, &ws.split_buckets, &ws.split_partials,
                      &ws.pt_a, &ws.pt_b, &ws.pt_prefix, &ws.pt_levels, &ws.sort_tmp, &ws.sort_coarse, &ws.sort_fine})
        b->release();
    }
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (void* p : ctx->stage_pinned)
        if (p) cudaFreeHost(p);
    for (cudaEvent_t e : ctx->stage_ev)
        if (e) cudaEventDestroy(e);
    for (auto& s : ctx->slots) {
        s.scalars.release();
        for (DevBuf* b : {&s.sort_ws.counts, &s.sort_ws.offsets, &s.sort_ws.entries, &s.sort_ws.scan_tmp, &s.sort_ws.sort_tmp, &s.sort_ws.sort_coarse, &s.sort_ws.sort_fine}) b->release();
        if (s.sorted) cudaEventDestroy(s.sorted);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.done) cudaEventDestroy(s.done);
        if (s.h_parts) cudaFreeHost(s.h_parts);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->sort_stream) cudaStreamDestroy(ctx->sort_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* halo_last_error(halo_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
const char* halo_curve_name(void) { return HALO_CURVE_NAME; }
uint64_t halo_kernel_launches(halo_ctx* ctx) { return ctx ? ctx->kernel_launches : 0; }
uint64_t halo_num_generators(halo_ctx* ctx) { return ctx ? ctx->n_gens : 0; }
int halo_set_msm_window(halo_ctx* ctx, int c) {
    if (!ctx || c < 0 || c > 20) return HALO_EINVAL;
    ctx->force_c = c;
    return HALO_OK;
}
int halo_set_tuning(halo_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return HALO_EINVAL;
    if (!strcmp(key, "acc_static")) ctx->tune_acc_static = value;
    else if (!strcmp(key, "acc_blocks_per_sm")) ctx->tune_acc_blocks_per_sm = value;
    else if (!strcmp(key, "acc_quad")) ctx->tune_acc_quad = value;
    else if (!strcmp(key, "acc_quad_max_buckets")) ctx->tune_acc_quad_max_buckets = value;
    else if (!strcmp(key, "acc_quad_lanes")) ctx->tune_acc_quad_lanes = value;
    else if (!strcmp(key, "acc_quad_blocks")) ctx->tune_acc_quad_blocks = value;
    else if (!strcmp(key, "pair_passes")) ctx->tune_pair_passes = value;
    else if (!strcmp(key, "split_blocking")) ctx->tune_split_blocking = value;
    else if (!strcmp(key, "split_first_16ths")) ctx->tune_split_first_16ths = value < 1 ? 1 : value > 15 ? 15 : value;
    else if (!strcmp(key, "split_second_16ths")) ctx->tune_split_second_16ths = value < 0 ? 0 : value > 14 ? 14 : value;
    else if (!strcmp(key, "sort_ahead")) ctx->tune_sort_ahead = value;
    else if (!strcmp(key, "stage_pageable")) ctx->tune_stage_pageable = value;
    else if (!strcmp(key, "stage_threads")) ctx->tune_stage_threads = value;
    else if (!strcmp(key, "reduce_quad")) ctx->tune_reduce_quad = value;
    else if (!strcmp(key, "sort2")) ctx->tune_sort2 = value;
    else if (!strcmp(key, "sort2_min_lg")) ctx->tune_sort2_min_lg = value;
    else if (!strcmp(key, "pair_bwd_async")) ctx->tune_pair_bwd_async = value;
    else if (!strcmp(key, "ipa_defer_rounds")) ctx->tune_ipa_defer = value;
    else if (!strcmp(key, "ipa_fold_call_min_lg")) ctx->tune_fold_call_min_lg = value;
    else if (!strcmp(key, "ipa_defer2_rounds")) ctx->tune_ipa_defer2 = value;
    else if (!strcmp(key, "ipa_two_lanes")) ctx->tune_ipa_two_lanes = value;
    else if (!strcmp(key, "ipa_freeze_len")) ctx->tune_ipa_freeze_len = value;
    else if (!strcmp(key, "ipa_frozen_c")) ctx->tune_ipa_frozen_c = value;
    else return fail(ctx, HALO_EINVAL, "halo_set_tuning: unknown key");
    return HALO_OK;
}
int halo_set_profiling(halo_ctx* ctx, int on) {
    if (!ctx) return HALO_EINVAL;
    ctx->profile = on != 0;
    return HALO_OK;
}
int halo_last_msm_timings(halo_ctx* ctx, float out_ms[6]) {
    if (!ctx) return HALO_EINVAL;
    const Timings& t = ctx->last;
    out_ms[0] = t.digits_ms;
    out_ms[1] = t.scan_ms;
    out_ms[2] = t.scatter_ms;
    out_ms[3] = t.accumulate_ms;
    out_ms[4] = t.reduce_ms;
    out_ms[5] = t.total_ms;
    return HALO_OK;
}

int halo_derive_generators(halo_ctx* ctx, uint64_t n) { return halo_derive_generators_range(ctx, 0, n); }

int halo_points_sum(const uint64_t* points_jac, uint64_t g, uint64_t out_jac[12]) {
    if (!out_jac || (!points_jac && g)) return HALO_EINVAL;
    xyzz_t acc;
    xyzz_set_inf(acc);
    for (uint64_t i = 0; i < g; i++) {
        jac_t j;
        memcpy(&j, points_jac + 12 * i, 96);
        xyzz_t q;
        jac_to_xyzz(q, j);
        xyzz_add(acc, q);
    }
    out_jac_from_xyzz(acc, out_jac);
    return HALO_OK;
}

int halo_points_equal(const uint64_t a_jac[12], const uint64_t b_jac[12]) {
    jac_t a, b;
    memcpy(&a, a_jac, 96);
    memcpy(&b, b_jac, 96);
    bool ia = fp_is_zero(a.z), ib = fp_is_zero(b.z);
    if (ia || ib) return ia && ib;
    fq_t za2, zb2, l, r;
    fp_sqr(za2, a.z);
    fp_sqr(zb2, b.z);
    fp_mul(l, a.x, zb2);
    fp_mul(r, b.x, za2);
    if (!fp_eq(l, r)) return 0;
    fp_mul(l, a.y, zb2);
    fp_mul(l, l, b.z);
    fp_mul(r, b.y, za2);
    fp_mul(r, r, a.z);
    return fp_eq(l, r) ? 1 : 0;
}

int halo_timer_start(halo_ctx* ctx) {
    if (!ctx) return HALO_EINVAL;
    HALO_TRY(ctx)
    HALO_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
    HALO_CATCH(ctx)
}
int halo_timer_stop(halo_ctx* ctx, float* elapsed_ms) {
    if (!ctx || !elapsed_ms) return HALO_EINVAL;
    HALO_TRY(ctx)
    HALO_CUDA(cudaEventRecord(ctx->ev[7], ctx->stream));
    HALO_CUDA(cudaEventSynchronize(ctx->ev[7]));
    HALO_CUDA(cudaEventElapsedTime(elapsed_ms, ctx->ev[6], ctx->ev[7]));
    HALO_CATCH(ctx)
}

int halo_derive_generators_range(halo_ctx* ctx, uint64_t first, uint64_t n) {
    if (!ctx) return HALO_EINVAL;
    if (n == 0 || n > ctx->max_n) return fail(ctx, HALO_EINVAL, "halo_derive_generators: n exceeds max_n");
    HALO_TRY(ctx)
    // P_0 = S, P_1 = H land in a small staging buffer; G_i = P_{i+2} directly in the resident array
    ctx->gens.reserve(n * sizeof(affine_t));
    ctx->stage_misc.reserve(2 * sizeof(affine_t));
    params_derive_points(ctx, 0, 2, ctx->stage_misc.as<affine_t>());
    params_derive_points(ctx, 2 + first, n, ctx->gens.as<affine_t>());
    affine_t sh[2];
    HALO_CUDA(cudaMemcpyAsync(sh, ctx->stage_misc.p, sizeof sh, cudaMemcpyDeviceToHost, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->S = sh[0];
    ctx->H = sh[1];
    ctx->have_SH = true;
    ctx->n_gens = n;
    ctx->pre_n = 0;  // tables of precomputed multiples are stale
    HALO_CATCH(ctx)
}

int halo_precompute_generators(halo_ctx* ctx, int c) {
    if (!ctx) return HALO_EINVAL;
    if (ctx->n_gens == 0) return fail(ctx, HALO_ESTATE, "halo_precompute_generators: no generators resident");
    if (c != 0 && (c < 8 || c > 24)) return fail(ctx, HALO_EINVAL, "halo_precompute_generators: window must be 0 (auto) or 8..24");
    HALO_TRY(ctx)
    msm_precompute_tables(ctx, c);
    HALO_CATCH(ctx)
}

int halo_set_fixed_base(halo_ctx* ctx, int on) {
    if (!ctx) return HALO_EINVAL;
    ctx->use_fixed = on != 0;
    return HALO_OK;
}

int halo_load_generators(halo_ctx* ctx, const uint64_t S_jac[12], const uint64_t H_jac[12], const uint64_t* gs_affine,
                         uint64_t n) {
    if (!ctx || !gs_affine) return HALO_EINVAL;
    if (n == 0 || n > ctx->max_n) return fail(ctx, HALO_EINVAL, "halo_load_generators: n exceeds max_n");
    HALO_TRY(ctx)
    ctx->gens.reserve(n * sizeof(affine_t));
    HALO_CUDA(cudaMemcpyAsync(ctx->gens.p, gs_affine, n * sizeof(affine_t), cudaMemcpyHostToDevice, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (S_jac && H_jac) {
        jac_t s, h;
        memcpy(&s, S_jac, 96);
        memcpy(&h, H_jac, 96);
        xyzz_t t;
        jac_to_xyzz(t, s);
        xyzz_to_affine(ctx->S, t);
        jac_to_xyzz(t, h);
        xyzz_to_affine(ctx->H, t);
        ctx->have_SH = true;
    }
    ctx->n_gens = n;
    ctx->pre_n = 0;
    HALO_CATCH(ctx)
}

// ---- generator store ---------------------------------------------------------------------------------
// The reference keeps its public parameters as generated source (main.rs:47-67 writes them, consts.rs holds 16 386 of
// them, and report.md:2081-2086 names that as the limit on n).  Here they are a flat file of the device records:
//   header (64 B): magic "HALOGEN1", curve name (8 B, zero padded), n, record size (64), checksum, 24 B reserved
//   S, H           2 x 64 B Montgomery affine
//   G_0 .. G_{n-1} n x 64 B Montgomery affine (x | y, little-endian limbs: the bytes of consts.rs' mk_aff! arguments)
// Loading streams the file through two pinned buffers (the read of chunk k+1 overlaps the copy of chunk k), verifies
// the checksum and checks on the device that every record is a canonical point on the curve.
namespace {
constexpr char STORE_MAGIC[8] = {'H', 'A', 'L', 'O', 'G', 'E', 'N', '1'};
constexpr size_t STORE_CHUNK = 32u << 20;
struct StoreHeader {
    char magic[8];
    char curve[8];
    uint64_t n;
    uint64_t record_bytes;
    uint64_t checksum;
    uint64_t reserved[3];
};
static_assert(sizeof(StoreHeader) == 64, "store header layout");
// 64-bit multiply-rotate checksum over the u64 words of S, H and the records, in file order
inline uint64_t store_mix(uint64_t h, const void* data, size_t bytes) {
    const uint64_t* w = static_cast<const uint64_t*>(data);
    for (size_t i = 0; i < bytes / 8; i++) {
        h ^= w[i];
        h *= 0x9e3779b97f4a7c15ull;
        h = (h << 29) | (h >> 35);
    }
    return h;
}
struct FileCloser {
    FILE* f;
    ~FileCloser() {
        if (f) fclose(f);
    }
};
struct PinnedPair {
    void* p[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~PinnedPair() {
        for (int i = 0; i < 2; i++) {
            if (p[i]) cudaFreeHost(p[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
    }
};
}  // namespace

int halo_save_generators(halo_ctx* ctx, const char* path) {
    if (!ctx || !path) return HALO_EINVAL;
    if (ctx->n_gens == 0 || !ctx->have_SH) return fail(ctx, HALO_ESTATE, "halo_save_generators: no generators resident");
    HALO_TRY(ctx)
    FileCloser fc{fopen(path, "wb")};
    if (!fc.f) return fail(ctx, HALO_EIO, "halo_save_generators: cannot open the file for writing");
    StoreHeader hd = {};
    memcpy(hd.magic, STORE_MAGIC, 8);
    strncpy(hd.curve, HALO_CURVE_NAME, 8);
    hd.n = ctx->n_gens;
    hd.record_bytes = sizeof(affine_t);
    affine_t sh[2] = {ctx->S, ctx->H};
    uint64_t sum = store_mix(0, sh, sizeof sh);
    if (fwrite(&hd, sizeof hd, 1, fc.f) != 1 || fwrite(sh, sizeof sh, 1, fc.f) != 1)
        return fail(ctx, HALO_EIO, "halo_save_generators: write failed");
    PinnedPair pp;
    for (int i = 0; i < 2; i++) {
        HALO_CUDA(cudaMallocHost(&pp.p[i], STORE_CHUNK));
        HALO_CUDA(cudaEventCreateWithFlags(&pp.ev[i], cudaEventDisableTiming));
    }
    const size_t total = ctx->n_gens * sizeof(affine_t);
    const char* src = ctx->gens.as<char>();
    size_t issued = 0, written = 0;
    int k = 0;
    auto issue = [&](int b) {
        size_t len = total - issued < STORE_CHUNK ? total - issued : STORE_CHUNK;
        HALO_CUDA(cudaMemcpyAsync(pp.p[b], src + issued, len, cudaMemcpyDeviceToHost, ctx->stream));
        HALO_CUDA(cudaEventRecord(pp.ev[b], ctx->stream));
        issued += len;
    };
    if (issued < total) issue(0);
    while (written < total) {  // the device-to-host copy of chunk k+1 runs while chunk k is hashed and written
        int b = k & 1;
        if (issued < total) issue(b ^ 1);
        HALO_CUDA(cudaEventSynchronize(pp.ev[b]));
        size_t len = total - written < STORE_CHUNK ? total - written : STORE_CHUNK;
        sum = store_mix(sum, pp.p[b], len);
        if (fwrite(pp.p[b], 1, len, fc.f) != len) return fail(ctx, HALO_EIO, "halo_save_generators: write failed");
        written += len;
        k++;
    }
    hd.checksum = sum;
    if (fseek(fc.f, 0, SEEK_SET) != 0 || fwrite(&hd, sizeof hd, 1, fc.f) != 1 || fflush(fc.f) != 0)
        return fail(ctx, HALO_EIO, "halo_save_generators: write failed");
    HALO_CATCH(ctx)
}

int halo_load_generators_file(halo_ctx* ctx, const char* path, uint64_t n) {
    if (!ctx || !path) return HALO_EINVAL;
    HALO_TRY(ctx)
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) return fail(ctx, HALO_EIO, "halo_load_generators_file: cannot open the file");
    StoreHeader hd;
    affine_t sh[2];
    if (fread(&hd, sizeof hd, 1, fc.f) != 1 || fread(sh, sizeof sh, 1, fc.f) != 1)
        return fail(ctx, HALO_EIO, "halo_load_generators_file: truncated header");
    char curve[8] = {};
    strncpy(curve, HALO_CURVE_NAME, 8);
    if (memcmp(hd.magic, STORE_MAGIC, 8) != 0 || hd.record_bytes != sizeof(affine_t))
        return fail(ctx, HALO_EINVAL, "halo_load_generators_file: not a generator store");
    if (memcmp(hd.curve, curve, 8) != 0)
        return fail(ctx, HALO_EINVAL, "halo_load_generators_file: the store holds points of the other curve");
    if (n == 0) n = hd.n;
    if (n > hd.n) return fail(ctx, HALO_EINVAL, "halo_load_generators_file: the store holds fewer generators than requested");
    if (n > ctx->max_n) return fail(ctx, HALO_EINVAL, "halo_load_generators_file: n exceeds max_n");
    ctx->n_gens = 0;  // the resident set is replaced; a failure below leaves the context without generators
    ctx->pre_n = 0;
    ctx->gens.reserve(n * sizeof(affine_t));
    PinnedPair pp;
    for (int i = 0; i < 2; i++) {
        HALO_CUDA(cudaMallocHost(&pp.p[i], STORE_CHUNK));
        HALO_CUDA(cudaEventCreateWithFlags(&pp.ev[i], cudaEventDisableTiming));
    }
    // a prefix of the store is a valid parameter set (G_0..G_{n-1}); the checksum covers the whole file, so it is
    // verified only when all of it is read -- the on-curve check below covers every record either way
    const size_t total = n * sizeof(affine_t);
    uint64_t sum = store_mix(0, sh, sizeof sh);
    size_t done = 0;
    for (int k = 0; done < total; k++) {
        int b = k & 1;
        if (k >= 2) HALO_CUDA(cudaEventSynchronize(pp.ev[b]));  // the copy out of this buffer two chunks ago
        size_t len = total - done < STORE_CHUNK ? total - done : STORE_CHUNK;
        if (fread(pp.p[b], 1, len, fc.f) != len) return fail(ctx, HALO_EIO, "halo_load_generators_file: truncated file");
        sum = store_mix(sum, pp.p[b], len);
        HALO_CUDA(cudaMemcpyAsync(ctx->gens.as<char>() + done, pp.p[b], len, cudaMemcpyHostToDevice, ctx->stream));
        HALO_CUDA(cudaEventRecord(pp.ev[b], ctx->stream));
        done += len;
    }
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n == hd.n && sum != hd.checksum) return fail(ctx, HALO_EINVAL, "halo_load_generators_file: checksum mismatch");
    ctx->stage_bases.reserve(sizeof sh);
    HALO_CUDA(cudaMemcpyAsync(ctx->stage_bases.p, sh, sizeof sh, cudaMemcpyHostToDevice, ctx->stream));
    if (params_count_off_curve(ctx, ctx->stage_bases.as<affine_t>(), 2) != 0 ||
        params_count_off_curve(ctx, ctx->gens.as<affine_t>(), n) != 0)
        return fail(ctx, HALO_EINVAL, "halo_load_generators_file: the store holds records that are not points on the curve");
    ctx->S = sh[0];
    ctx->H = sh[1];
    ctx->have_SH = true;
    ctx->n_gens = n;
    HALO_CATCH(ctx)
}

int halo_get_generators(halo_ctx* ctx, uint64_t off, uint64_t n, uint64_t* out_affine) {
    if (!ctx || !out_affine) return HALO_EINVAL;
    if (off + n > ctx->n_gens) return fail(ctx, HALO_EINVAL, "halo_get_generators: range exceeds resident generators");
    HALO_TRY(ctx)
    HALO_CUDA(cudaMemcpyAsync(out_affine, ctx->gens.as<affine_t>() + off, n * sizeof(affine_t), cudaMemcpyDeviceToHost,
                              ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    HALO_CATCH(ctx)
}

int halo_get_SH(halo_ctx* ctx, uint64_t S_jac[12], uint64_t H_jac[12]) {
    if (!ctx || !ctx->have_SH) return fail(ctx, HALO_ESTATE, "halo_get_SH: parameters not loaded");
    xyzz_t t;
    xyzz_from_affine(t, ctx->S);
    out_jac_from_xyzz(t, S_jac);
    xyzz_from_affine(t, ctx->H);
    out_jac_from_xyzz(t, H_jac);
    return HALO_OK;
}

int halo_derive_points(halo_ctx* ctx, uint64_t start, uint64_t count, uint64_t* out_affine) {
    if (!ctx || !out_affine) return HALO_EINVAL;
    HALO_TRY(ctx)
    ctx->stage_bases.reserve(count * sizeof(affine_t));
    params_derive_points(ctx, start, count, ctx->stage_bases.as<affine_t>());
    HALO_CUDA(cudaMemcpyAsync(out_affine, ctx->stage_bases.p, count * sizeof(affine_t), cudaMemcpyDeviceToHost, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    HALO_CATCH(ctx)
}

int halo_msm_gens_resident(halo_ctx* ctx, const void* d_scalars, uint64_t off, uint64_t n, uint64_t out_jac[12]) {
    if (!ctx || !out_jac || (!d_scalars && n)) return HALO_EINVAL;
    if (off + n > ctx->n_gens) return fail(ctx, HALO_ESTATE, "halo_msm_gens: range exceeds resident generators");
    HALO_TRY(ctx)
    xyzz_t r;
    msm_gens_device(ctx, reinterpret_cast<const fr_t*>(d_scalars), off, n, r);
    out_jac_from_xyzz(r, out_jac);
    HALO_CATCH(ctx)
}

int halo_msm_gens(halo_ctx* ctx, const uint64_t* scalars, uint64_t off, uint64_t n, uint64_t out_jac[12]) {
    if (!ctx || !out_jac || (!scalars && n)) return HALO_EINVAL;
    if (off + n > ctx->n_gens) return fail(ctx, HALO_ESTATE, "halo_msm_gens: range exceeds resident generators");
    // Large calls are split into point slices that go through the two pipeline slots: the host-to-device copies of the later
    // slices overlap the kernels of the earlier ones (2^24 scalars from pinned memory, round 1: 45.4 ms unsplit, 41.1 ms with
    // two halves, 38.2 ms with 5/16 + 11/16; round 2: 37.0 ms with 5/16 + 11/16, 36.05 ms with 2/16 + 5/16 + 9/16; from
    // pageable memory 41.9 -> 39.8 ms); the partial sums are added on the host.
    if (ctx->tune_split_blocking > 0 && n >= ((uint64_t)1 << ctx->tune_split_blocking) && !ctx->slots[0].active && !ctx->slots[1].active) {
        // slices in sixteenths of n: the first slice's copy is the exposed one; with a second cut the third slice is submitted
        // into the first slice's slot as soon as that one is collected (two slots in flight at any time)
        const int a = ctx->tune_split_first_16ths, b = ctx->tune_split_second_16ths;
        uint64_t cut[4] = {0, n / 16 * (uint64_t)a, 0, n};
        int ns = 2;
        if (b > 0 && a + b < 16) {
            cut[2] = n / 16 * (uint64_t)(a + b);
            ns = 3;
        } else {
            cut[2] = n;
        }
        int tk[3];
        uint64_t parts[36];
        int rc = 0, submitted = 0, collected = 0;
        for (int i = 0; i < ns && !rc; i++) {
            if (i == 2) {
                rc = halo_msm_gens_collect(ctx, tk[0], parts);
                collected = 1;
                if (rc) break;
            }
            rc = halo_msm_gens_submit(ctx, scalars + 4 * cut[i], off + cut[i], cut[i + 1] - cut[i], &tk[i]);
            if (!rc) submitted = i + 1;
        }
        for (int i = collected; i < submitted; i++) {  // on failure this drains what is in flight; the first error is the one returned
            const int rci = halo_msm_gens_collect(ctx, tk[i], parts + 12 * i);
            if (!rc) rc = rci;
        }
        if (rc) return rc;
        return halo_points_sum(parts, (uint64_t)ns, out_jac);
    }
    HALO_TRY(ctx)
    ctx->stage_scalars.reserve((n ? n : 1) * sizeof(fr_t));
    if (n) h2d_copy(ctx, ctx->stage_scalars.p, scalars, n * sizeof(fr_t), ctx->stream);
    xyzz_t r;
    msm_gens_device(ctx, ctx->stage_scalars.as<fr_t>(), off, n, r);
    out_jac_from_xyzz(r, out_jac);
    HALO_CATCH(ctx)
}

// Shared body of the two submit entry points: host scalars are copied on the copy stream first; device-resident
// scalars (resident = true) are used in place and must stay untouched until the ticket is collected.
static int submit_impl(halo_ctx* ctx, const uint64_t* scalars, bool resident, uint64_t off, uint64_t n, int* ticket) {
    if (!ctx || !ticket || (!scalars && n)) return HALO_EINVAL;
    if (off + n > ctx->n_gens) return fail(ctx, HALO_ESTATE, "halo_msm_gens_submit: range exceeds resident generators");
    HALO_TRY(ctx)
    async_init(ctx);
    const int si = ctx->next_slot;
    halo_ctx::AsyncSlot& s = ctx->slots[si];
    if (s.active) return fail(ctx, HALO_ESTATE, "halo_msm_gens_submit: both pipeline slots are in flight; collect one first");
    s.empty = n == 0;
    if (!s.empty) {
        const bool ahead_on = ctx->tune_sort_ahead > 0 && n >= ((uint64_t)1 << 22);
        // (delaying the sort until the HBM-bound pass-0 gathers of the MSM in front are over measured slower: the
        // throttled sort then no longer fits under what is left of that MSM)
        MsmInput in;
        if (resident) {
            in.scalars = reinterpret_cast<const fr_t*>(scalars);
            // the caller's producer of the scalars ran on a stream we do not know: like halo_msm_gens_resident, the
            // contract is that they are complete when this call is made
        } else {
            s.scalars.reserve(n * sizeof(fr_t));
            h2d_copy(ctx, s.scalars.p, scalars, n * sizeof(fr_t), ctx->copy_stream);
            HALO_CUDA(cudaEventRecord(s.copied, ctx->copy_stream));
            HALO_CUDA(cudaStreamWaitEvent(ahead_on ? ctx->sort_stream : ctx->stream, s.copied, 0));
            in.scalars = s.scalars.as<fr_t>();
        }
        in.n = (uint32_t)n;
        const bool fixed = ctx->use_fixed && ctx->pre_n == ctx->n_gens && ctx->gens_pre.p && n >= (1u << 17) && n * 8 >= ctx->pre_n;
        if (fixed) {
            in.bases = ctx->gens_pre.as<affine_t>();
            in.fixed_stride = (uint32_t)ctx->pre_n;
            in.fixed_first = (uint32_t)off;
            s.plan = ctx->pre_plan;
        } else {
            in.bases = ctx->gens.as<affine_t>() + off;
            s.plan = msm_make_plan(n, ctx->force_c);
        }
        ctx->ws.wsums.reserve((size_t)6 * 3 * MSM_MAX_WINDOWS * sizeof(xyzz_t));
        xyzz_t* d_parts = ctx->ws.wsums.as<xyzz_t>() + (size_t)(4 + si) * 3 * MSM_MAX_WINDOWS;  // slots 0-3 belong to msm_batch
        // the sort buffers of this slot were last read by the accumulation of the MSM two submits ago, which the caller
        // has collected (slot free) -- so the sort may start as soon as the scalars have arrived
        SortAhead ahead{&s.sort_ws, ctx->sort_stream, s.sorted, ctx->tune_sort_ahead};
        msm_enqueue(ctx, in, s.plan, d_parts, 0, ahead_on ? &ahead : nullptr);
        const int nwin = s.plan.fixed ? 1 : s.plan.W;
        HALO_CUDA(cudaMemcpyAsync(s.h_parts, d_parts, (size_t)3 * nwin * sizeof(xyzz_t), cudaMemcpyDeviceToHost, ctx->stream));
        HALO_CUDA(cudaEventRecord(s.done, ctx->stream));
    }
    s.active = true;
    *ticket = si;
    ctx->next_slot = si ^ 1;
    HALO_CATCH(ctx)
}

int halo_msm_gens_submit(halo_ctx* ctx, const uint64_t* scalars, uint64_t off, uint64_t n, int* ticket) {
    return submit_impl(ctx, scalars, false, off, n, ticket);
}

int halo_msm_gens_submit_resident(halo_ctx* ctx, const void* d_scalars, uint64_t off, uint64_t n, int* ticket) {
    return submit_impl(ctx, reinterpret_cast<const uint64_t*>(d_scalars), true, off, n, ticket);
}

int halo_msm_gens_collect(halo_ctx* ctx, int ticket, uint64_t out_jac[12]) {
    if (!ctx || !out_jac || ticket < 0 || ticket > 1) return HALO_EINVAL;
    halo_ctx::AsyncSlot& s = ctx->slots[ticket];
    if (!s.active) return fail(ctx, HALO_ESTATE, "halo_msm_gens_collect: ticket not in flight");
    HALO_TRY(ctx)
    xyzz_t r;
    if (s.empty) {
        xyzz_set_inf(r);
    } else {
        HALO_CUDA(cudaEventSynchronize(s.done));
        msm_finish_host(s.h_parts, s.plan, r);
    }
    s.active = false;
    out_jac_from_xyzz(r, out_jac);
    HALO_CATCH(ctx)
}

int halo_msm(halo_ctx* ctx, const uint64_t* bases_affine, const uint8_t* inf_flags, const uint64_t* scalars, uint64_t n,
             uint64_t out_jac[12]) {
    if (!ctx || !out_jac || ((!scalars || !bases_affine) && n)) return HALO_EINVAL;
    if (n > ctx->max_n) return fail(ctx, HALO_EINVAL, "halo_msm: n exceeds max_n");
    HALO_TRY(ctx)
    ctx->stage_scalars.reserve((n ? n : 1) * sizeof(fr_t));
    ctx->stage_bases.reserve((n ? n : 1) * sizeof(affine_t));
    if (n) {
        h2d_copy(ctx, ctx->stage_scalars.p, scalars, n * sizeof(fr_t), ctx->stream);
        h2d_copy(ctx, ctx->stage_bases.p, bases_affine, n * sizeof(affine_t), ctx->stream);
        if (inf_flags) {
            ctx->stage_misc.reserve(n);
            HALO_CUDA(cudaMemcpyAsync(ctx->stage_misc.p, inf_flags, n, cudaMemcpyHostToDevice, ctx->stream));
            k_mark_infinity<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->stage_bases.as<affine_t>(),
                                                                                 ctx->stage_misc.as<uint8_t>(), n);
            ctx->kernel_launches++;
        }
    }
    xyzz_t r;
    msm_device(ctx, ctx->stage_bases.as<affine_t>(), ctx->stage_scalars.as<fr_t>(), n, r);
    out_jac_from_xyzz(r, out_jac);
    HALO_CATCH(ctx)
}

int halo_msm_jac(halo_ctx* ctx, const uint64_t* bases_jac, const uint64_t* scalars, uint64_t n, uint64_t out_jac[12]) {
    if (!ctx || !out_jac || ((!scalars || !bases_jac) && n)) return HALO_EINVAL;
    if (n > ctx->max_n) return fail(ctx, HALO_EINVAL, "halo_msm_jac: n exceeds max_n");
    HALO_TRY(ctx)
    ctx->stage_scalars.reserve((n ? n : 1) * sizeof(fr_t));
    ctx->stage_bases.reserve((n ? n : 1) * sizeof(affine_t));
    ctx->stage_misc.reserve((n ? n : 1) * sizeof(jac_t));
    if (n) {
        HALO_CUDA(cudaMemcpyAsync(ctx->stage_scalars.p, scalars, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        HALO_CUDA(cudaMemcpyAsync(ctx->stage_misc.p, bases_jac, n * sizeof(jac_t), cudaMemcpyHostToDevice, ctx->stream));
        k_jac_to_affine<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->stage_misc.as<jac_t>(),
                                                                             ctx->stage_bases.as<affine_t>(), n);
        ctx->kernel_launches++;
    }
    xyzz_t r;
    msm_device(ctx, ctx->stage_bases.as<affine_t>(), ctx->stage_scalars.as<fr_t>(), n, r);
    out_jac_from_xyzz(r, out_jac);
    HALO_CATCH(ctx)
}


// ---- scalar vectors (group.rs:13-15, :29-37) ----------------------------------------------------------
int halo_scalar_dot(halo_ctx* ctx, const uint64_t* xs, const uint64_t* ys, uint64_t n, uint64_t out[4]) {
    if (!ctx || !out || ((!xs || !ys) && n)) return HALO_EINVAL;
    HALO_TRY(ctx)
    ctx->stage_scalars.reserve((2 * (n ? n : 1) + 1 + VEC_DOT_MAX_BLOCKS) * sizeof(fr_t));
    fr_t* a = ctx->stage_scalars.as<fr_t>();
    fr_t* b = a + n;
    fr_t* res = b + n;
    if (n) {
        HALO_CUDA(cudaMemcpyAsync(a, xs, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        HALO_CUDA(cudaMemcpyAsync(b, ys, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    vec_dot(ctx, a, b, n, res + 1, res);
    HALO_CUDA(cudaMemcpyAsync(out, res, 32, cudaMemcpyDeviceToHost, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    HALO_CATCH(ctx)
}

int halo_construct_powers(halo_ctx* ctx, const uint64_t z[4], uint64_t n, uint64_t* out) {
    if (!ctx || !z || (!out && n)) return HALO_EINVAL;
    HALO_TRY(ctx)
    if (n) {
        ctx->stage_scalars.reserve(n * sizeof(fr_t));
        fr_t zz;
        memcpy(&zz, z, 32);
        vec_powers(ctx, zz, n, ctx->stage_scalars.as<fr_t>());
        HALO_CUDA(cudaMemcpyAsync(out, ctx->stage_scalars.p, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
        HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    HALO_CATCH(ctx)
}

// ---- K5: h(X) (pcdl.rs:56-77, :338; acc.rs:85-94) --------------------------------------------------------
static int check_lg(halo_ctx* ctx, uint32_t lg_n, bool need_gens) {
    if (lg_n > 30 || ((uint64_t)1 << lg_n) > ctx->max_n) return fail(ctx, HALO_EINVAL, "h: 2^lg_n exceeds max_n");
    if (need_gens && ((uint64_t)1 << lg_n) > ctx->n_gens) return fail(ctx, HALO_ESTATE, "h: 2^lg_n exceeds resident generators");
    return HALO_OK;
}

int halo_h_expand(halo_ctx* ctx, const uint64_t* xis, uint32_t lg_n, uint64_t* out) {
    if (!ctx || !xis || !out) return HALO_EINVAL;
    if (int rc = check_lg(ctx, lg_n, false)) return rc;
    HALO_TRY(ctx)
    uint64_t n = (uint64_t)1 << lg_n;
    ctx->stage_scalars.reserve(n * sizeof(fr_t));
    fr_t one;
    fp_one(one);
    vec_h_expand(ctx, reinterpret_cast<const fr_t*>(xis), (int)lg_n, one, false, ctx->stage_scalars.as<fr_t>());
    HALO_CUDA(cudaMemcpyAsync(out, ctx->stage_scalars.p, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    HALO_CATCH(ctx)
}

int halo_h_msm(halo_ctx* ctx, const uint64_t* xis, uint32_t lg_n, uint64_t out_jac[12]) {
    if (!ctx || !xis || !out_jac) return HALO_EINVAL;
    if (int rc = check_lg(ctx, lg_n, true)) return rc;
    HALO_TRY(ctx)
    uint64_t n = (uint64_t)1 << lg_n;
    ctx->stage_scalars.reserve(n * sizeof(fr_t));
    fr_t one;
    fp_one(one);
    vec_h_expand(ctx, reinterpret_cast<const fr_t*>(xis), (int)lg_n, one, false, ctx->stage_scalars.as<fr_t>());
    xyzz_t r;
    msm_gens_device(ctx, ctx->stage_scalars.as<fr_t>(), 0, n, r);
    out_jac_from_xyzz(r, out_jac);
    HALO_CATCH(ctx)
}

int halo_h_msm_with(halo_ctx* ctx, const uint64_t* xis, uint32_t lg_n, const uint64_t* bases_affine, const uint8_t* inf_flags,
                    const uint64_t* scalars, uint64_t k, uint64_t out_h_jac[12], uint64_t out_small_jac[12]) {
    if (!ctx || !xis || !out_h_jac || !out_small_jac || ((!bases_affine || !scalars) && k)) return HALO_EINVAL;
    if (int rc = check_lg(ctx, lg_n, true)) return rc;
    if (k > 4096) return fail(ctx, HALO_EINVAL, "halo_h_msm_with: the companion MSM is meant to be small (k <= 4096)");
    HALO_TRY(ctx)
    uint64_t n = (uint64_t)1 << lg_n;
    ctx->stage_scalars.reserve(n * sizeof(fr_t));
    fr_t one;
    fp_one(one);
    vec_h_expand(ctx, reinterpret_cast<const fr_t*>(xis), (int)lg_n, one, false, ctx->stage_scalars.as<fr_t>());
    MsmInput in[2];
    in[0].scalars = ctx->stage_scalars.as<fr_t>();
    in[0].n = (uint32_t)n;
    if (ctx->use_fixed && ctx->pre_n == ctx->n_gens && ctx->gens_pre.p && n >= (1u << 17) && n * 8 >= ctx->pre_n) {
        in[0].bases = ctx->gens_pre.as<affine_t>();
        in[0].fixed_stride = (uint32_t)ctx->pre_n;
        in[0].fixed_first = 0;
    } else {
        in[0].bases = ctx->gens.as<affine_t>();
    }
    // companion: bases | scalars | infinity flags staged in stage_misc
    const size_t kk = k ? k : 1;
    ctx->stage_misc.reserve(kk * (sizeof(affine_t) + sizeof(fr_t) + 1));
    affine_t* d_b = ctx->stage_misc.as<affine_t>();
    fr_t* d_s = reinterpret_cast<fr_t*>(d_b + kk);
    uint8_t* d_i = reinterpret_cast<uint8_t*>(d_s + kk);
    if (k) {
        HALO_CUDA(cudaMemcpyAsync(d_b, bases_affine, k * sizeof(affine_t), cudaMemcpyHostToDevice, ctx->stream));
        HALO_CUDA(cudaMemcpyAsync(d_s, scalars, k * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        if (inf_flags) {
            HALO_CUDA(cudaMemcpyAsync(d_i, inf_flags, k, cudaMemcpyHostToDevice, ctx->stream));
            k_mark_infinity<<<(unsigned)((k + 255) / 256), 256, 0, ctx->stream>>>(d_b, d_i, k);
            ctx->kernel_launches++;
        }
    }
    in[1].bases = d_b;
    in[1].scalars = d_s;
    in[1].n = (uint32_t)k;
    xyzz_t out[2];
    ctx->force_two_lanes = k > 0;
    try {
        msm_batch(ctx, in, 2, out);
    } catch (...) {
        ctx->force_two_lanes = false;
        throw;
    }
    ctx->force_two_lanes = false;
    out_jac_from_xyzz(out[0], out_h_jac);
    out_jac_from_xyzz(out[1], out_small_jac);
    HALO_CATCH(ctx)
}

int halo_msm_multi(halo_ctx* ctx, const halo_msm_desc* descs, uint32_t count, uint64_t* out_jac) {
    if (!ctx || (!descs && count) || (!out_jac && count)) return HALO_EINVAL;
    size_t bytes = 0;
    for (uint32_t k = 0; k < count; k++) {
        const halo_msm_desc& d = descs[k];
        if (d.n && !d.scalars) return HALO_EINVAL;
        if (d.n > 4096) return fail(ctx, HALO_EINVAL, "halo_msm_multi: the MSMs of a batch are meant to be small (n <= 4096)");
        if (!d.bases_affine && d.off + d.n > ctx->n_gens)
            return fail(ctx, HALO_EINVAL, "halo_msm_multi: range exceeds the resident generators");
        bytes += d.n * (sizeof(fr_t) + (d.bases_affine ? sizeof(affine_t) + 16 : 0)) + 64;
    }
    HALO_TRY(ctx)
    // one staging area for the whole batch: per MSM  scalars | bases | infinity flags  (16-byte aligned pieces)
    ctx->stage_misc.reserve(bytes ? bytes : 64);
    char* cur = ctx->stage_misc.as<char>();
    auto take = [&](size_t len) {
        char* p = cur;
        cur += (len + 15) & ~(size_t)15;
        return p;
    };
    std::vector<MsmInput> ins(count);
    for (uint32_t k = 0; k < count; k++) {
        const halo_msm_desc& d = descs[k];
        if (!d.n) continue;
        fr_t* d_s = reinterpret_cast<fr_t*>(take(d.n * sizeof(fr_t)));
        HALO_CUDA(cudaMemcpyAsync(d_s, d.scalars, d.n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        ins[k].scalars = d_s;
        ins[k].n = (uint32_t)d.n;
        if (d.bases_affine) {
            affine_t* d_b = reinterpret_cast<affine_t*>(take(d.n * sizeof(affine_t)));
            HALO_CUDA(cudaMemcpyAsync(d_b, d.bases_affine, d.n * sizeof(affine_t), cudaMemcpyHostToDevice, ctx->stream));
            if (d.inf_flags) {
                uint8_t* d_i = reinterpret_cast<uint8_t*>(take(d.n));
                HALO_CUDA(cudaMemcpyAsync(d_i, d.inf_flags, d.n, cudaMemcpyHostToDevice, ctx->stream));
                k_mark_infinity<<<(unsigned)((d.n + 255) / 256), 256, 0, ctx->stream>>>(d_b, d_i, d.n);
                ctx->kernel_launches++;
            }
            ins[k].bases = d_b;
        } else {
            ins[k].bases = ctx->gens.as<affine_t>() + d.off;
        }
    }
    for (uint32_t k0 = 0; k0 < count; k0 += 4) {  // msm_batch: up to 4 MSMs on two lanes, one synchronisation
        const int c = (int)(count - k0 < 4 ? count - k0 : 4);
        xyzz_t out[4];
        msm_batch(ctx, ins.data() + k0, c, out);
        for (int k = 0; k < c; k++) out_jac_from_xyzz(out[k], out_jac + 12 * (k0 + k));
    }
    HALO_CATCH(ctx)
}

// out != NULL: coefficients to the host.  out == NULL: the polynomial stays on the device (ctx->poly_dev) and
// *degree_out receives its degree (DensePolynomial::degree: index of the highest non-zero coefficient).
static int h_lincomb_impl(halo_ctx* ctx, const uint64_t* h0, uint64_t n_h0, const uint64_t* alphas, const uint64_t* xis,
                          uint64_t m, uint32_t lg_n, uint64_t* out, uint64_t* degree_out) {
    if (!ctx || (!out && !degree_out) || (!h0 && n_h0) || ((!alphas || !xis) && m)) return HALO_EINVAL;
    if (int rc = check_lg(ctx, lg_n, false)) return rc;
    uint64_t n = (uint64_t)1 << lg_n;
    if (n_h0 > n) return fail(ctx, HALO_EINVAL, "halo_h_lincomb: h_0 longer than n");
    HALO_TRY(ctx)
    DevBuf& buf = out ? ctx->stage_scalars : ctx->poly_dev;
    buf.reserve(n * sizeof(fr_t));
    fr_t* d = buf.as<fr_t>();
    HALO_CUDA(cudaMemsetAsync(d, 0, n * sizeof(fr_t), ctx->stream));
    if (n_h0) HALO_CUDA(cudaMemcpyAsync(d, h0, n_h0 * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    const fr_t* al = reinterpret_cast<const fr_t*>(alphas);
    const fr_t* xs = reinterpret_cast<const fr_t*>(xis);
    for (uint64_t i = 0; i < m; i++) vec_h_expand(ctx, xs + i * (lg_n + 1), (int)lg_n, al[i + 1], true, d);  // acc.rs:90-92
    if (out) {
        HALO_CUDA(cudaMemcpyAsync(out, d, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
        HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        // degree: scan down from the top in chunks (the leading coefficient is non-zero except with negligible probability)
        ctx->poly_n = n;
        uint64_t hi = n, deg = 0;
        bool found = false;
        std::vector<fr_t> chunk(64);
        while (hi > 0 && !found) {
            uint64_t lo = hi > 64 ? hi - 64 : 0;
            HALO_CUDA(cudaMemcpyAsync(chunk.data(), d + lo, (hi - lo) * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
            HALO_CUDA(cudaStreamSynchronize(ctx->stream));
            for (uint64_t i = hi; i-- > lo;)
                if (!fp_is_zero(chunk[i - lo])) {
                    deg = i;
                    found = true;
                    break;
                }
            hi = lo;
        }
        *degree_out = deg;
    }
    HALO_CATCH(ctx)
}

int halo_h_lincomb(halo_ctx* ctx, const uint64_t* h0, uint64_t n_h0, const uint64_t* alphas, const uint64_t* xis, uint64_t m,
                   uint32_t lg_n, uint64_t* out) {
    if (!out) return HALO_EINVAL;
    return h_lincomb_impl(ctx, h0, n_h0, alphas, xis, m, lg_n, out, nullptr);
}

int halo_h_lincomb_resident(halo_ctx* ctx, const uint64_t* h0, uint64_t n_h0, const uint64_t* alphas, const uint64_t* xis,
                            uint64_t m, uint32_t lg_n, uint64_t* degree_out) {
    if (!degree_out) return HALO_EINVAL;
    return h_lincomb_impl(ctx, h0, n_h0, alphas, xis, m, lg_n, nullptr, degree_out);
}

}  // extern "C"
