// fused_loop.cuh -- the whole iteration loop of the windowed re-correlation schedule as ONE cooperative launch.
//
// For the latency-bound shapes (BASELINE configs[0]: one 2^15-sample signal, 512 x 512 dictionary) an iteration of
// the stream-ordered loop is two launches (k_apply, k_corr) of a few microseconds each plus the gaps between them.
// Here every CTA owns its atom pair(s) for the whole pursuit -- the pair spectrum stays in registers -- and one
// iteration is
//   A  every CTA reduces the signal's row maxima to the winner (redundantly: 2 KB from L2),
//   B  loads the residual window around it, subtracts the scaled atom on the fly (two roundings, as k_apply) and
//      transforms the window in its own shared memory (no window spectrum in global memory),
//   C  multiplies with its pair spectrum, inverse-transforms, refreshes block and row maxima of its two rows,
//   D  one grid-wide barrier (arrival counter in global memory; the launch is cooperative, so all CTAs are resident).
// The residual itself is updated by ONE extra CTA per signal that owns no atom pair (blockIdx.x == gridDim.x - 1): it
// selects the winner like everybody else, records the event, computes the new samples, waits until every worker CTA
// has read the old window (a second arrival counter that only the writers wait for) and stores them before it
// arrives at the iteration barrier -- off the critical path of the transforms.
// Reference loop: modules/matchingpursuit.py:298-328.
#pragma once
#include "kernels.cuh"

namespace mpb {

struct FusedArgs {
    const C32* pairspec;      // (npairs, M)
    const float* dict;        // (K, A) unit-normed
    float* residual;          // (B, N)
    float* bm_val;            // (B, nloc, NB)
    int* bm_pos;
    float* row_val;           // (B, nloc): the tables iteration 0 selects from (first pass)
    int* row_pos;
    float* row_val2;          // (B, nloc): second copy -- iteration s selects from copy s & 1 and writes copy (s+1) & 1, so a
    int* row_pos2;            // CTA that runs ahead into its row phase never changes what a slower one is still selecting from
    const C32* tw1;
    const C32* tw2;
    int npairs, nloc, atom_lo, n_atoms, A, N, NB, blk_shift;
    int n_steps;
    int* atom_out;            // (B, n_steps)
    int* pos_out;
    float* val_out;
    unsigned* gbar;           // [0] iteration barrier, [1] "old window read" arrivals; both zero at launch
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Wait until the arrival counter reaches `want`.  The launch is cooperative, so every CTA is resident and the wait is
// microseconds; a counter that has not moved for 10 s means a broken launch, and the kernel traps (the host sees a
// launch failure at its next synchronisation) instead of hanging the device.
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned want) {
    if (ld_acquire_u32(p) >= want) return;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_u32(p) < want) {
        if (global_ns() - t0 > 10000000000ull) __trap();
    }
}
__device__ __forceinline__ void red_relaxed_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int M>
__global__ void __launch_bounds__((BlockFft<M, float>::T < 256 ? 256 : BlockFft<M, float>::T), 1)
k_pursue_fused(const FusedArgs a) {
    using F = BlockFft<M, float>;
    constexpr int TPB = F::T < 256 ? 256 : F::T;
    constexpr int NT = TPB / F::T;   // transforms (atom pairs) per CTA
    constexpr int NW = F::T / 32;    // warps per transform
    constexpr int NWC = TPB / 32;    // warps per CTA
    extern __shared__ __align__(16) unsigned char smraw[];
    C32* stw2 = reinterpret_cast<C32*>(smraw);
    const int sb = threadIdx.x / F::T, tl = threadIdx.x % F::T;
    // two buffers per transform: the exchanges of the passes go through `sm`, the window spectrum (between the two
    // transforms) and the outputs (before the block maxima) through `sm2` -- one CTA per SM, so the space is free and
    // two of the barriers that a shared buffer would need are not
    C32* sm = stw2 + 256 + (size_t)sb * 2 * F::SMEM_CPX;
    C32* sm2 = sm + F::SMEM_CPX;
    float2* sY = reinterpret_cast<float2*>(sm2);         // outputs as (atom 2q, atom 2q+1) pairs, after the transform
    __shared__ float2 s_bv_static[NT * 64];
    float2* sBV = s_bv_static + sb * 64;                 // [which*32 + i] = (value, position as int bits)
    __shared__ float s_v[32];
    __shared__ int s_k[32], s_p[32];
    __shared__ Best s_best;
    for (int i = threadIdx.x; i < 256; i += TPB) stw2[i] = a.tw2[i];

    const int b = blockIdx.y;
    const bool writer = blockIdx.x == gridDim.x - 1;     // no atom pair: events and residual of signal b
    int q = blockIdx.x * NT + sb;
    const bool q_ok = !writer && q < a.npairs;
    if (q >= a.npairs) q = a.npairs - 1;
    const bool second = 2 * q + 1 < a.nloc;
    const int blk = 1 << a.blk_shift;
    const int warp = tl >> 5, lane = tl & 31;
    const int w_all = threadIdx.x >> 5, l_all = threadIdx.x & 31;
    const unsigned nworkers = (gridDim.x - 1) * gridDim.y, nctas = gridDim.x * gridDim.y;
    float* __restrict__ res = a.residual + (size_t)b * a.N;

    // the pair spectrum never changes during a pursuit: registers
    C32 eq[F::E];
    if (!writer) {
        const C32* __restrict__ Eq = a.pairspec + (size_t)q * M;
#pragma unroll
        for (int e = 0; e < F::E; ++e) {
            const float2 y = __ldg(reinterpret_cast<const float2*>(Eq + F::in_index(tl, e)));
            eq[e] = C32{y.x, y.y};
        }
    }
    __syncthreads();

    for (int s = 0; s < a.n_steps; ++s) {
        // ---- A. winner of signal b: max over the row maxima, lowest atom on ties (as block_best) --------------------
        // The tables were written by other CTAs before the barrier: L2 loads (.cg), never a stale L1 line.  Values and
        // positions are fetched together (one L2 round trip), and so are the previous maxima of this CTA's own rows.
        const float* __restrict__ rv = ((s & 1) ? a.row_val2 : a.row_val) + (size_t)b * a.nloc;
        const int* __restrict__ rp = ((s & 1) ? a.row_pos2 : a.row_pos) + (size_t)b * a.nloc;
        float* __restrict__ rv_next = ((s & 1) ? a.row_val : a.row_val2) + (size_t)b * a.nloc;
        int* __restrict__ rp_next = ((s & 1) ? a.row_pos : a.row_pos2) + (size_t)b * a.nloc;
        {
            // CTA-wide: every thread takes nloc/TPB rows (all CTAs read the same few KB; a per-warp redundant scan that
            // needs no CTA barrier was measured slower, 14.9 against 10.9 us per iteration at configs[0])
            float v = -INFINITY;
            int k = INT_MAX, at = 0;
            for (int i = threadIdx.x; i < a.nloc; i += TPB) {
                const float c = __ldcg(rv + i);
                const int cp = __ldcg(rp + i);
                if (c > v) { v = c; k = i; at = cp; }
            }
            const int mine = k;
            warp_argmax(v, k);
            // the lane whose candidate won carries its position along
            const unsigned holders = __ballot_sync(0xffffffffu, mine == k && k != INT_MAX);
            at = __shfl_sync(0xffffffffu, at, holders ? __ffs(holders) - 1 : 0);
            if (l_all == 0) { s_v[w_all] = v; s_k[w_all] = k; s_p[w_all] = at; }
            __syncthreads();
            if (w_all == 0) {
                v = l_all < NWC ? s_v[l_all] : -INFINITY;
                k = l_all < NWC ? s_k[l_all] : INT_MAX;
                at = l_all < NWC ? s_p[l_all] : 0;
                const int mine2 = k;
                warp_argmax(v, k);
                const unsigned h2 = __ballot_sync(0xffffffffu, mine2 == k && k != INT_MAX);
                at = __shfl_sync(0xffffffffu, at, h2 ? __ffs(h2) - 1 : 0);
                if (l_all == 0) {
                    if (k == INT_MAX) {                  // every candidate was NaN or -inf: the first entry
                        k = 0;
                        v = __ldcg(rv);
                        at = __ldcg(rp);
                    }
                    Best best;
                    best.value = v;
                    best.atom = a.atom_lo + k;
                    best.position = at;
                    best.pad = 0;
                    s_best = best;
                }
            }
            __syncthreads();
        }
        Best w = s_best;
        w.atom = min(max(w.atom, 0), a.n_atoms - 1);     // memory safety whatever the tables hold (as k_apply)
        w.position = min(max(w.position, 0), a.N - 1);
        const int p = w.position;
        const float* __restrict__ d = a.dict + (size_t)w.atom * a.A;
        const int keep = min(a.A, a.N - p);
        const bool last_step = s == a.n_steps - 1;

        if (writer) {
            // ---- the writer: event, new residual samples (computed now, stored once nobody reads the old ones) -------
            if (threadIdx.x == 0) {
                a.atom_out[(size_t)b * a.n_steps + s] = w.atom;
                a.pos_out[(size_t)b * a.n_steps + s] = p;
                a.val_out[(size_t)b * a.n_steps + s] = w.value;
            }
            constexpr int PER = 8;                       // samples per thread and round: atoms up to PER * TPB per round
            for (int i0 = 0; i0 < keep; i0 += PER * TPB) {
                float x[PER];
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int i = i0 + j * TPB + threadIdx.x;
                    x[j] = (i < keep) ? __fsub_rn(__ldcg(res + p + i), __fmul_rn(w.value, __ldg(d + i))) : 0.f;
                }
                if (i0 == 0 && !last_step) {             // every worker CTA has read the old window
                    if (threadIdx.x == 0) {
                        wait_counter(a.gbar + 1, (unsigned)(s + 1) * nworkers);
                    }
                    __syncthreads();
                }
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int i = i0 + j * TPB + threadIdx.x;
                    if (i < keep) res[p + i] = x[j];
                }
            }
            if (last_step) break;
        } else {
            if (last_step) break;                        // only the subtraction is left
            const int first = max(0, p - a.A + 1), last = min(a.N - 1, p + a.A - 1);
            const int blk0 = first >> a.blk_shift, t0 = blk0 << a.blk_shift, nvb = (last >> a.blk_shift) - blk0 + 1;

            // ---- B. residual window with the winner already subtracted, forward transform ---------------------------
            C32 r[F::E];
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                const int t = t0 + F::in_index(tl, e);
                float x = (t < a.N) ? __ldcg(res + t) : 0.f;
                const int i = t - p;
                if (i >= 0 && i < keep) x = __fsub_rn(x, __fmul_rn(w.value, __ldg(d + i)));
                r[e] = C32{x, 0.f};
            }
            // previous maxima of this transform's rows (row phase below): in flight behind the transforms
            const int which_own = NW >= 2 ? warp : 0;
            float old_v[2] = {0.f, 0.f};
            int old_p[2] = {0, 0};
            if (q_ok) {
#pragma unroll
                for (int j = 0; j < (NW >= 2 ? 1 : 2); ++j) {
                    const int which = which_own + j;
                    if (which < 2 && (which == 0 || second)) {
                        old_v[j] = __ldcg(rv + 2 * q + which);
                        old_p[j] = __ldcg(rp + 2 * q + which);
                    }
                }
            }
            F::template pass1<-1>(r, tl, sm, a.tw1);
            __syncthreads();
            // every thread of this CTA holds its share of the OLD window: tell the writers (never waited for here)
            // (relaxed: the loads have returned -- their values went through pass 1 -- before this is issued)
            if (threadIdx.x == 0) red_relaxed_add_u32(a.gbar + 1, 1u);
            F::template pass2<-1>(r, tl, sm, stw2);
            __syncthreads();
            F::template pass3<-1>(r, tl, sm);
#pragma unroll
            for (int e = 0; e < F::E; ++e) sm2[F::bin_addr(F::out_index(tl, e))] = r[e];
            __syncthreads();                             // spectrum staged; every pass-3 load of `sm` has returned

            // ---- C. product with the pair spectrum, inverse transform, block and row maxima (as k_corr) --------------
#pragma unroll
            for (int e = 0; e < F::E; ++e) r[e] = cmul(sm2[F::slot_addr(tl, e)], eq[e]);
            F::template pass1<1>(r, tl, sm, a.tw1);
            __syncthreads();
            F::template pass2<1>(r, tl, sm, stw2);
            __syncthreads();
            F::template pass3<1>(r, tl, sm);
            const int stage_end = nvb * blk;
            const int limit = min(stage_end, a.N - t0);  // valid outputs are m in [0, limit)
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                const int m = F::out_index(tl, e);
                if (m < stage_end) sY[m] = make_float2(r[e].x, r[e].y);
            }
            __syncthreads();
            for (int i = warp; i < nvb; i += NW) {
                const int hi = min((i + 1) * blk, limit);
                float va = -INFINITY, vb = -INFINITY;
                int ia = INT_MAX, ib = INT_MAX;
                for (int m = i * blk + lane; m < hi; m += 32) {  // ascending inside the lane: strict > keeps the first maximum
                    const float2 c = sY[m];
                    if (c.x > va) { va = c.x; ia = m; }
                    if (c.y > vb) { vb = c.y; ib = m; }
                }
                warp_argmax(va, ia);
                warp_argmax(vb, ib);
                if (lane == 0 && q_ok) {
                    const int pa = (ia == INT_MAX) ? INT_MAX : t0 + ia;
                    const int pb = (ib == INT_MAX) ? INT_MAX : t0 + ib;
                    const size_t o = ((size_t)b * a.nloc + 2 * q) * a.NB + blk0 + i;
                    a.bm_val[o] = va;
                    a.bm_pos[o] = pa;
                    if (second) {
                        a.bm_val[o + a.NB] = vb;
                        a.bm_pos[o + a.NB] = pb;
                    }
                    sBV[i] = make_float2(va, __int_as_float(pa));
                    sBV[32 + i] = make_float2(vb, __int_as_float(pb));
                }
            }
            __syncthreads();                             // refreshed block maxima are staged in sBV
#pragma unroll
            for (int j = 0; j < (NW >= 2 ? 1 : 2); ++j) {
                const int which = which_own + j;
                if (!q_ok || which >= 2 || (which == 1 && !second)) continue;
                const size_t rowi = (size_t)b * a.nloc + 2 * q + which;
                const size_t o = rowi * a.NB;
                const int old_b = old_p[j] >> a.blk_shift;
                const bool old_ok = old_b < blk0 || old_b >= blk0 + nvb;
                float v = -INFINITY;
                int at = INT_MAX;
                if (lane < nvb) {
                    const float2 c = sBV[which * 32 + lane];
                    v = c.x;
                    at = __float_as_int(c.y);
                }
                if (old_ok) {
                    if (lane == 31) take_better(v, at, old_v[j], old_p[j]);   // nvb < 32: lane 31 is free
                } else {
                    rescan_row(a.bm_val + o, a.bm_pos + o, a.NB, blk0, nvb, lane, v, at, a.blk_shift);
                }
                warp_argmax(v, at);
                if (lane == 0) {
                    rv_next[2 * q + which] = v;
                    rp_next[2 * q + which] = (at == INT_MAX) ? 0 : at;
                }
            }
        }
        // ---- D. grid barrier: tables and residual of this iteration are visible to every CTA -------------------------
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release_add_u32(a.gbar, 1u);             // release: this CTA's table / residual stores come first
            wait_counter(a.gbar, (unsigned)(s + 1) * nctas);
        }
        __syncthreads();
    }
}

}  // namespace mpb
