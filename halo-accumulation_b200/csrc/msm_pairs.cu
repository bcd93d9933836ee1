// msm_pairs.cu -- K2b: the first levels of the bucket accumulation as a tree of pairwise AFFINE additions with
// batched inversion (part of the Pippenger MSM behind group.rs:18-26; see msm.cu for the pipeline around it).
//
// The bucket-sorted entry list is laid out so that every bucket's segment starts at a multiple of 2^P slots and is
// padded with "infinity" slots up to the next multiple (msm.cu rounds the histogram before the scan).  One tree pass
// then is a flat, bucket-oblivious map over an array of slots:   out[q] = in[2q] + in[2q+1]   -- pairs never straddle
// buckets, no per-pass offsets or searches, and after P passes bucket b owns the slots [off_b >> P, off_{b+1} >> P),
// which the XYZZ kernel of msm.cu finishes (plus its splitting of oversized buckets, so degenerate scalar
// distributions need nothing special here).
//
// An affine addition costs one inversion; the inversions of a whole pass are shared (Montgomery's trick) through a
// hierarchy that never serialises more than a handful of multiplications per thread:
//   k_pair_fwd    per thread K pairs: den_q (x2 - x1, or 2 y1 for a doubling), running product -> prefix[q], total -> T0[t]
//   k_prod_up     T_{j+1}[u] = product of KU values of T_j, prefixes kept                            (until <= 16384 values)
//   k_inv         one Fermat inversion per surviving value, all in parallel
//   k_prod_down   inverse totals back down: T_j[i] <- 1 / T_j[i]
//   k_pair_bwd    per pair: 1/den from the thread's inverse total and prefix[q], then lambda, x3, y3 -> out[q]
// 5M + 1S per addition (+ 3/K + 3/(K KU) .. for the hierarchy) instead of the 8M + 2S of a mixed XYZZ addition.
// The price is memory traffic (operands are read twice, prefixes written and read once: ~320 B per addition instead of
// a 64-byte gather), which HBM3e carries while the integer pipe stays the bound.
//
// Exactness: infinity operands, P + P (tangent) and P + (-P) are classified identically by the forward and backward
// kernels (same function, same inputs) and never contribute a zero denominator, so duplicate and cancelling bases --
// legal inputs of msm_unchecked -- give the same group element as the XYZZ path.
#include "common.cuh"
#include "msm.cuh"

namespace halo {

constexpr int PT_K = 16;   // pairs per thread in k_pair_fwd / k_pair_bwd
constexpr int PT_KU = 16;  // fan-in of the product hierarchy
constexpr uint32_t PT_INV_MAX = 65536;  // values inverted one per thread at the top of the hierarchy (division steps: 24 us of latency; was 16384 with the 0.105 ms Fermat chain)
constexpr uint32_t PT_SENTINEL = 0xffffffffu;  // entry that stands for the point at infinity (padding)

enum : int { PT_SKIP = 0, PT_ADD = 1, PT_DBL = 2 };

// Infinity inside the intermediate arrays: x = 2^256 - 1 (not a canonical residue), so the forward pass can classify a
// pair from the x coordinates alone.
__device__ __forceinline__ bool pt_x_is_inf(const fq_t& x) { return x.v[7] == 0xffffffffu; }
__device__ __forceinline__ void pt_set_inf(affine_t& p) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        p.x.v[i] = 0xffffffffu;
        p.y.v[i] = 0u;
    }
}

// One operand of pass 0: table / base gather through the entry (index | sign << 31), sign applied.
__device__ __forceinline__ void pt_gather(affine_t& p, bool& inf, uint32_t ent, const affine_t* __restrict__ bases, uint32_t n,
                                          const affine_t* __restrict__ tail_bases) {
    if (ent == PT_SENTINEL) {
        inf = true;
        affine_set_inf(p);
        return;
    }
    const uint32_t idx = ent & 0x7fffffffu;
    p = idx < n ? bases[idx] : tail_bases[idx - n];
    inf = affine_is_inf(p);
    if (ent >> 31) fp_neg(p.y, p.y);
}

// Warp-cooperative version for a whole warp step: lane l needs the operands named by its entry pair e = (e.x, e.y).
// A lane loading its own 64-byte base issues 4 LDG.128 that each touch 32 different lines, and the load/store unit, not
// HBM, becomes the limit (measured: 20 G gathers/s per-thread against 43 G/s when 4 adjacent lanes fetch one base with
// ONE instruction, profiles/r01_random_gather_hbm.jsonl).  So 8 rounds of one LDG.128 per lane fetch the 64 bases of the
// warp (round r: chunk l & 3 of the base of lane (r & 3) * 8 + (l >> 2); rounds 0-3 operand a, 4-7 operand b) and a
// padded shared-memory tile transposes them back to their owners.  All 32 lanes must call it (padding lanes pass the
// sentinel pair).
constexpr int PT_SM_STRIDE = 5;  // uint4 per staged base: 64 B + 16 B pad (conflict-free 64-byte reads by 8 lanes)
struct PairTile {
    uint4 v[32][2][PT_SM_STRIDE];
};
__device__ __forceinline__ void pt_gather_pair_coop(affine_t& a, bool& ainf, affine_t& b, bool& binf, uint2 e,
                                                    const affine_t* __restrict__ bases, uint32_t n,
                                                    const affine_t* __restrict__ tail_bases, PairTile& tile, uint32_t lane) {
    const uint32_t c = lane & 3u;
    constexpr int BATCH = 8;  // loads in flight per lane
#pragma unroll
    for (int r0 = 0; r0 < 8; r0 += BATCH) {
        uint4 v[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const int r = r0 + k;
            const int owner = (r & 3) * 8 + (int)(lane >> 2);
            const uint32_t ent = __shfl_sync(0xffffffffu, r < 4 ? e.x : e.y, owner);
            v[k] = make_uint4(0, 0, 0, 0);
            if (ent != PT_SENTINEL) {
                const uint32_t idx = ent & 0x7fffffffu;
                const affine_t* p = idx < n ? bases + idx : tail_bases + (idx - n);
                v[k] = __ldg(reinterpret_cast<const uint4*>(p) + c);
            }
        }
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const int r = r0 + k;
            tile.v[(r & 3) * 8 + (lane >> 2)][r >> 2][c] = v[k];
        }
    }
    __syncwarp();
    uint4* ap = reinterpret_cast<uint4*>(&a);
    uint4* bp = reinterpret_cast<uint4*>(&b);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ap[k] = tile.v[lane][0][k];
        bp[k] = tile.v[lane][1][k];
    }
    __syncwarp();
    ainf = e.x == PT_SENTINEL || affine_is_inf(a);
    binf = e.y == PT_SENTINEL || affine_is_inf(b);
    if (e.x != PT_SENTINEL && (e.x >> 31)) fp_neg(a.y, a.y);
    if (e.y != PT_SENTINEL && (e.y >> 31)) fp_neg(b.y, b.y);
}

// Classification shared by both directions.  Needs y only when the x coordinates coincide.
//   PT_ADD: den = bx - ax.   PT_DBL: den = 2 ay (ay != 0).   PT_SKIP: no inversion; the result is inf, a or b.
__device__ __forceinline__ int pt_classify(fq_t& den, bool ainf, bool binf, const fq_t& ax, const fq_t& bx, const fq_t& ay,
                                           const fq_t& by) {
    if (ainf || binf) return PT_SKIP;
    fp_sub(den, bx, ax);
    if (!fp_is_zero(den)) return PT_ADD;
    if (fp_eq(ay, by) && !fp_is_zero(ay)) {
        fp_dbl(den, ay);
        return PT_DBL;
    }
    return PT_SKIP;  // a = -b (or a 2-torsion point of an off-curve input): the sum is infinity
}

// pair index of thread-local step s: warps own 32 * PT_K consecutive pairs, lanes interleave (coalesced 128-byte pairs)
__device__ __forceinline__ uint32_t pt_pair_index(uint32_t gwarp, int s, uint32_t lane) { return (gwarp * PT_K + s) * 32u + lane; }

// ---- forward ---------------------------------------------------------------------------------------------------------
template <bool PASS0>
__global__ void __launch_bounds__(128) k_pair_fwd(const uint32_t* __restrict__ entries, const affine_t* __restrict__ bases, uint32_t n,
                                                  const affine_t* __restrict__ tail_bases, const fq_t* __restrict__ in_x,
                                                  const fq_t* __restrict__ in_y, const uint32_t* __restrict__ total_slots, int pass,
                                                  fq_t* __restrict__ prefix, fq_t* __restrict__ totals) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gwarp = tid >> 5, lane = tid & 31u;
    const uint32_t Q = *total_slots >> (pass + 1);
    __shared__ PairTile tiles[PASS0 ? 4 : 1];
    PairTile& tile = tiles[PASS0 ? (threadIdx.x >> 5) : 0];
    fq_t run;
    fp_one(run);
    if (pt_pair_index(gwarp, 0, 0) < Q) {
#pragma unroll 2
        for (int s = 0; s < PT_K; s++) {
            if (pt_pair_index(gwarp, s, 0) >= Q) break;  // warp uniform
            const uint32_t q = pt_pair_index(gwarp, s, lane);
            const bool live = q < Q;
            fq_t ax, bx, ay, by, den;
            bool ainf, binf;
            if (PASS0) {
                uint2 e = make_uint2(PT_SENTINEL, PT_SENTINEL);
                if (live) e = reinterpret_cast<const uint2*>(entries)[q];
                affine_t a, b;
                pt_gather_pair_coop(a, ainf, b, binf, e, bases, n, tail_bases, tile, lane);
                ax = a.x;
                ay = a.y;
                bx = b.x;
                by = b.y;
            } else {
                if (!live) break;
                ax = in_x[2 * (size_t)q];  // the slot arrays are SoA: this pass reads x only, half the bytes of x | y records
                bx = in_x[2 * (size_t)q + 1];
                ainf = pt_x_is_inf(ax);
                binf = pt_x_is_inf(bx);
                if (!ainf && !binf && fp_eq(ax, bx)) {
                    ay = in_y[2 * (size_t)q];
                    by = in_y[2 * (size_t)q + 1];
                } else {
                    fp_zero(ay);
                    fp_zero(by);
                }
            }
            const int mode = live ? pt_classify(den, ainf, binf, ax, bx, ay, by) : PT_SKIP;
            if (mode != PT_SKIP) {
                prefix[q] = run;
                fp_mul(run, run, den);
            }
        }
    }
    totals[tid] = run;
}

// ---- product hierarchy -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prod_up(const fq_t* __restrict__ vals, uint32_t n, fq_t* __restrict__ pre,
                                                 fq_t* __restrict__ totals) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gwarp = tid >> 5, lane = tid & 31u;
    fq_t run;
    fp_one(run);
#pragma unroll 1
    for (int i = 0; i < PT_KU; i++) {
        const uint32_t idx = (gwarp * PT_KU + i) * 32u + lane;
        if (idx >= n) break;
        pre[idx] = run;
        fq_t v = vals[idx];
        fp_mul(run, run, v);
    }
    totals[tid] = run;
}
__global__ void __launch_bounds__(128) k_inv(fq_t* __restrict__ vals, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq_t v = vals[i], r;
    fp_inv(r, v);
    vals[i] = r;
}
// vals[i] <- 1 / vals[i], given the inverse of each thread's product in tot_inv
__global__ void __launch_bounds__(128) k_prod_down(fq_t* __restrict__ vals, uint32_t n, const fq_t* __restrict__ pre,
                                                   const fq_t* __restrict__ tot_inv) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gwarp = tid >> 5, lane = tid & 31u;
    fq_t inv = tot_inv[tid];
#pragma unroll 1
    for (int i = PT_KU - 1; i >= 0; i--) {
        const uint32_t idx = (gwarp * PT_KU + i) * 32u + lane;
        if (idx >= n) continue;
        fq_t v = vals[idx], p = pre[idx], o;
        fp_mul(o, inv, p);
        fp_mul(inv, inv, v);
        vals[idx] = o;
    }
}

// ---- backward ----------------------------------------------------------------------------------------------------------
// Pass 0 stages its operands through shared memory with cp.async: while the warp computes step s (32 pairs: 4M + 1S + 2 each)
// the 64 bases of step s - 1 are in flight, fetched cooperatively (4 adjacent lanes copy the four 16-byte chunks of one
// base with ONE instruction -- the access pattern that reaches HBM's random-access rate, see pt_gather_pair_coop) straight
// into the owner lane's row of the tile.  No registers hold data in flight (the register prefetch and the shuffle
// transpose of round 1 both lost to register pressure), entries are loaded two steps ahead and the step's prefix one step
// ahead (staging alone measured +-1 %: with the operands in shared memory every step still waited for that 32-byte load).
// Two 4 KiB buffers per warp, 32 KiB per CTA; 5 CTAs per SM measured best (6: spills, 4: too few warps).
struct BwdTile {
    uint4 v[2][32][4];  // [operand][owner lane][16-byte chunk]
};
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void pt_stage_pair_async(BwdTile& tile, uint2 e, const affine_t* __restrict__ bases, uint32_t n,
                                                    const affine_t* __restrict__ tail_bases, uint32_t lane) {
    const uint32_t c = lane & 3u;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int owner = (r & 3) * 8 + (int)(lane >> 2);
        const uint32_t ent = __shfl_sync(0xffffffffu, r < 4 ? e.x : e.y, owner);
        if (ent != PT_SENTINEL) {
            const uint32_t idx = ent & 0x7fffffffu;
            const affine_t* p = idx < n ? bases + idx : tail_bases + (idx - n);
            cp_async16(&tile.v[r >> 2][owner][c], reinterpret_cast<const uint4*>(p) + c);
        }
    }
    cp_async_commit();
}
__device__ __forceinline__ void pt_take_staged(affine_t& p, bool& inf, const BwdTile& tile, int operand, uint32_t ent, uint32_t lane) {
    if (ent == PT_SENTINEL) {
        inf = true;
        affine_set_inf(p);
        return;
    }
    uint4* pp = reinterpret_cast<uint4*>(&p);
#pragma unroll
    for (int k = 0; k < 4; k++) pp[k] = tile.v[operand][lane][k];
    inf = affine_is_inf(p);
    if (ent >> 31) fp_neg(p.y, p.y);
}

// The arithmetic of one pair, shared by both instantiations: 1/den from the thread's running inverse and prefix[q].
__device__ __forceinline__ void pt_pair_finish(affine_t& r, const affine_t& a, bool ainf, const affine_t& b, bool binf, fq_t& invtot,
                                               const fq_t& pre) {
    fq_t den;
    const int mode = pt_classify(den, ainf, binf, a.x, b.x, a.y, b.y);
    if (mode == PT_SKIP) {
        if (ainf && binf)
            pt_set_inf(r);
        else if (ainf)
            r = b;
        else if (binf)
            r = a;
        else
            pt_set_inf(r);  // a = -b
    } else {
        fq_t inv, num, lam, t;
        fp_mul(inv, invtot, pre);
        fp_mul(invtot, invtot, den);
        if (mode == PT_ADD) {
            fp_sub(num, b.y, a.y);
        } else {  // tangent: 3 x^2 / (2 y)
            fp_sqr(t, a.x);
            fp_dbl(num, t);
            fp_add(num, num, t);
        }
        fp_mul(lam, num, inv);
        fp_sqr(t, lam);
        fp_sub(t, t, a.x);
        fp_sub(r.x, t, b.x);
        fp_sub(t, a.x, r.x);
        fp_mul(t, lam, t);
        fp_sub(r.y, t, a.y);
    }
}

#ifndef HALO_PAIR_BWD0_BLOCKS
#define HALO_PAIR_BWD0_BLOCKS 5
#endif
__global__ void __launch_bounds__(128, HALO_PAIR_BWD0_BLOCKS) k_pair_bwd0(const uint32_t* __restrict__ entries, const affine_t* __restrict__ bases,
                                                                          uint32_t n, const affine_t* __restrict__ tail_bases,
                                                                          const uint32_t* __restrict__ total_slots,
                                                                          const fq_t* __restrict__ prefix, const fq_t* __restrict__ tot_inv,
                                                                          fq_t* __restrict__ out_x, fq_t* __restrict__ out_y) {
    __shared__ BwdTile tiles[4][2];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gwarp = tid >> 5, lane = tid & 31u;
    const uint32_t Q = *total_slots >> 1;
    if (pt_pair_index(gwarp, 0, 0) >= Q) return;  // warp uniform
    BwdTile* tile = tiles[threadIdx.x >> 5];
    fq_t invtot = tot_inv[tid];
    const uint2* ents = reinterpret_cast<const uint2*>(entries);
    const uint2 none = make_uint2(PT_SENTINEL, PT_SENTINEL);
    auto load_entry = [&](int s) {
        const uint32_t q = pt_pair_index(gwarp, s, lane);
        return q < Q ? ents[q] : none;
    };
    uint2 e0 = load_entry(PT_K - 1), e1 = load_entry(PT_K - 2);
    pt_stage_pair_async(tile[(PT_K - 1) & 1], e0, bases, n, tail_bases, lane);
    // the prefix of the step after this one is loaded a step ahead as well: with the operands staged, this 32-byte load was
    // what every step still waited for
    auto load_prefix = [&](int s) {
        const uint32_t q = pt_pair_index(gwarp, s, lane);
        fq_t v;
        fp_zero(v);
        if (q < Q) v = prefix[q];
        return v;
    };
    fq_t pre0 = load_prefix(PT_K - 1);
#pragma unroll 1
    for (int s = PT_K - 1; s >= 0; s--) {
        const uint2 e2 = s >= 2 ? load_entry(s - 2) : none;
        fq_t pre1;
        if (s >= 1) pre1 = load_prefix(s - 1); else fp_zero(pre1);
        if (s >= 1) {
            pt_stage_pair_async(tile[(s - 1) & 1], e1, bases, n, tail_bases, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const uint32_t q = pt_pair_index(gwarp, s, lane);
        if (q < Q) {
            affine_t a, b, r;
            bool ainf, binf;
            pt_take_staged(a, ainf, tile[s & 1], 0, e0.x, lane);
            pt_take_staged(b, binf, tile[s & 1], 1, e0.y, lane);
            pt_pair_finish(r, a, ainf, b, binf, invtot, pre0);
            out_x[q] = r.x;
            out_y[q] = r.y;
        }
        __syncwarp();  // every lane has read its row before the stage two steps on overwrites this buffer
        e0 = e1;
        e1 = e2;
        pre0 = pre1;
    }
}

// 7 CTAs per SM (72 registers, 24 bytes of spill): measured best of 4 (92 registers) .. 8 (64 registers): MSM 2^24
// 36.16 / 35.71 (6) / 35.52 (7) / 36.41 ms (8)
template <bool PASS0>
__global__ void __launch_bounds__(128, 7) k_pair_bwd(const uint32_t* __restrict__ entries, const affine_t* __restrict__ bases,
                                                     uint32_t n, const affine_t* __restrict__ tail_bases,
                                                     const fq_t* __restrict__ in_x, const fq_t* __restrict__ in_y,
                                                     const uint32_t* __restrict__ total_slots, int pass, const fq_t* __restrict__ prefix,
                                                     const fq_t* __restrict__ tot_inv, fq_t* __restrict__ out_x, fq_t* __restrict__ out_y) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gwarp = tid >> 5, lane = tid & 31u;
    const uint32_t Q = *total_slots >> (pass + 1);
    if (pt_pair_index(gwarp, 0, 0) >= Q) return;
    fq_t invtot = tot_inv[tid];
#pragma unroll 1
    for (int s = PT_K - 1; s >= 0; s--) {
        const uint32_t q = pt_pair_index(gwarp, s, lane);
        if (q >= Q) continue;
        affine_t a, b;
        bool ainf, binf;
        if (PASS0) {
            // per-lane gathers (the round-1 kernel, kept for A/B: halo_set_tuning "pair_bwd_async" = 0)
            const uint2 e = reinterpret_cast<const uint2*>(entries)[q];
            pt_gather(a, ainf, e.x, bases, n, tail_bases);
            pt_gather(b, binf, e.y, bases, n, tail_bases);
        } else {
            a.x = in_x[2 * (size_t)q];
            b.x = in_x[2 * (size_t)q + 1];
            a.y = in_y[2 * (size_t)q];
            b.y = in_y[2 * (size_t)q + 1];
            ainf = pt_x_is_inf(a.x);
            binf = pt_x_is_inf(b.x);
        }
        affine_t r;
        const fq_t pre = prefix[q];
        pt_pair_finish(r, a, ainf, b, binf, invtot, pre);
        out_x[q] = r.x;
        out_y[q] = r.y;
    }
}

// ---- host ------------------------------------------------------------------------------------------------------------------
static inline uint32_t ceil_div(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// Runs `passes` tree passes over the padded entry list (slots_max = host upper bound of the padded slot count, a
// multiple of 2^passes; the exact count is *total_slots on the device).  Returns the array of slots after the last pass.
const affine_t* pair_tree_enqueue(halo_ctx* ctx, MsmWorkspace& ws, cudaStream_t st, const MsmInput& in, const uint32_t* entries,
                                  const uint32_t* total_slots, uint64_t slots_max, int passes) {
    const uint64_t Q0 = slots_max >> 1;
    ws.pt_a.reserve(Q0 * sizeof(affine_t));
    ws.pt_b.reserve((Q0 >> 1) * sizeof(affine_t) + sizeof(affine_t));
    ws.pt_prefix.reserve(Q0 * sizeof(fq_t));
    // hierarchy geometry of pass 0 (the largest): level sizes n0 > n1 > .. until <= PT_INV_MAX
    auto level_sizes = [](uint64_t Q, uint32_t* sizes, uint32_t* grids) {
        int L = 0;
        uint32_t blocks = ceil_div(ceil_div(Q, 32u * PT_K), 4);
        if (blocks == 0) blocks = 1;
        sizes[L] = blocks * 128u;
        grids[L] = blocks;
        while (sizes[L] > PT_INV_MAX) {
            uint32_t b2 = ceil_div(ceil_div(sizes[L], 32u * PT_KU), 4);
            sizes[L + 1] = b2 * 128u;
            grids[L + 1] = b2;
            L++;
        }
        return L + 1;
    };
    uint32_t sizes[8], grids[8];
    int L = level_sizes(Q0, sizes, grids);
    size_t lv_total = 0;
    for (int j = 0; j < L; j++) lv_total += (size_t)sizes[j] * 2;  // values + prefixes per level
    ws.pt_levels.reserve(lv_total * sizeof(fq_t));

    // Slot arrays are SoA: after pass p the Q_p = slots_max >> (p + 1) slots live as x[0 .. Q_p) | y[0 .. Q_p) in the pass's
    // ping-pong buffer (the forward kernel of the next pass reads x only).
    const fq_t* src_x = nullptr;
    const fq_t* src_y = nullptr;
    const uint32_t nb = in.n;
    for (int p = 0; p < passes; p++) {
        const uint64_t Q = slots_max >> (p + 1);
        L = level_sizes(Q, sizes, grids);
        fq_t* lv = ws.pt_levels.as<fq_t>();
        fq_t* vals[8];
        fq_t* pres[8];
        for (int j = 0; j < L; j++) {
            vals[j] = lv;
            pres[j] = lv + sizes[j];
            lv += (size_t)sizes[j] * 2;
        }
        fq_t* dst_x = (p & 1) ? ws.pt_b.as<fq_t>() : ws.pt_a.as<fq_t>();
        fq_t* dst_y = dst_x + Q;
        fq_t* prefix = ws.pt_prefix.as<fq_t>();
        if (p == 0)
            k_pair_fwd<true><<<grids[0], 128, 0, st>>>(entries, in.bases, in.fixed_stride ? 0x7fffffffu : nb, in.tail_bases, nullptr, nullptr,
                                                       total_slots, p, prefix, vals[0]);
        else
            k_pair_fwd<false><<<grids[0], 128, 0, st>>>(nullptr, nullptr, 0, nullptr, src_x, src_y, total_slots, p, prefix, vals[0]);
        for (int j = 0; j + 1 < L; j++) k_prod_up<<<grids[j + 1], 128, 0, st>>>(vals[j], sizes[j], pres[j], vals[j + 1]);
        k_inv<<<ceil_div(sizes[L - 1], 128), 128, 0, st>>>(vals[L - 1], sizes[L - 1]);
        for (int j = L - 2; j >= 0; j--) k_prod_down<<<grids[j + 1], 128, 0, st>>>(vals[j], sizes[j], pres[j], vals[j + 1]);
        if (p == 0 && ctx->tune_pair_bwd_async)
            k_pair_bwd0<<<grids[0], 128, 0, st>>>(entries, in.bases, in.fixed_stride ? 0x7fffffffu : nb, in.tail_bases, total_slots, prefix,
                                                  vals[0], dst_x, dst_y);
        else if (p == 0)
            k_pair_bwd<true><<<grids[0], 128, 0, st>>>(entries, in.bases, in.fixed_stride ? 0x7fffffffu : nb, in.tail_bases, nullptr, nullptr,
                                                       total_slots, p, prefix, vals[0], dst_x, dst_y);
        else
            k_pair_bwd<false><<<grids[0], 128, 0, st>>>(nullptr, nullptr, 0, nullptr, src_x, src_y, total_slots, p, prefix, vals[0], dst_x, dst_y);
        ctx->kernel_launches += 2 + 2 * (uint64_t)(L - 1) + 1;
        src_x = dst_x;
        src_y = dst_y;
    }
    HALO_CUDA(cudaGetLastError());
    return reinterpret_cast<const affine_t*>(src_x);  // x[0 .. Q) | y[0 .. Q) with Q = slots_max >> passes (DIRECT mode of k_accumulate)
}

// In-place inversion of n non-zero field elements through the same product hierarchy (3 multiplications per element
// plus one Fermat inversion per ~256 .. 65536 elements instead of one per element).  `scratch` grows as needed.
void batch_invert(halo_ctx* ctx, cudaStream_t st, fq_t* vals, uint32_t n, DevBuf& scratch) {
    if (n == 0) return;
    uint32_t sizes[8], grids[8];
    int L = 0;
    sizes[0] = n;
    grids[0] = 0;
    while (sizes[L] > PT_INV_MAX) {
        uint32_t b2 = ceil_div(ceil_div(sizes[L], 32u * PT_KU), 4);
        sizes[L + 1] = b2 * 128u;
        grids[L + 1] = b2;
        L++;
    }
    L++;
    size_t total = 0;
    for (int j = 0; j < L; j++) total += (size_t)sizes[j] * (j == 0 ? 1 : 2);  // level 0: prefixes only (values are `vals`)
    scratch.reserve(total * sizeof(fq_t));
    fq_t* lv = scratch.as<fq_t>();
    fq_t* v[8];
    fq_t* pr[8];
    v[0] = vals;
    pr[0] = lv;
    lv += sizes[0];
    for (int j = 1; j < L; j++) {
        v[j] = lv;
        pr[j] = lv + sizes[j];
        lv += (size_t)sizes[j] * 2;
    }
    for (int j = 0; j + 1 < L; j++) k_prod_up<<<grids[j + 1], 128, 0, st>>>(v[j], sizes[j], pr[j], v[j + 1]);
    k_inv<<<ceil_div(sizes[L - 1], 128), 128, 0, st>>>(v[L - 1], sizes[L - 1]);
    for (int j = L - 2; j >= 0; j--) k_prod_down<<<grids[j + 1], 128, 0, st>>>(v[j], sizes[j], pr[j], v[j + 1]);
    ctx->kernel_launches += 2 * (uint64_t)(L - 1) + 1;
    HALO_CUDA(cudaGetLastError());
}

}  // namespace halo
