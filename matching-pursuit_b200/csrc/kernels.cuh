// kernels.cuh -- device kernels of the matching-pursuit engine (sm_100a).
//
// Data layout in HBM (all fp32 / int32, row-major):
//   residual   (B, N)
//   dict       (K, A)             unit-normed atoms (whole dictionary, replicated)
//   pairspec   (P, M) complex     P = ceil(K_owned/2); spectrum of the complex
//                                 sequence d[2q] + i*d[2q+1] with the inverse
//                                 kernel, scaled 1/M: one complex IFFT of
//                                 X_window * pairspec[q] yields the correlation
//                                 of the window with atom 2q (real part) and
//                                 atom 2q+1 (imaginary part)
//   winspec    (W, M) complex     forward FFT of residual windows
//   bm_val/bm_pos (B, K_owned, NB) max / argmax of each `blk` consecutive positions
//   row_val/row_pos (B, K_owned)   max / argmax of each map row
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <climits>
#include "fft_core.cuh"
#include "types.h"

namespace mpb {

constexpr int MODE_BLOCKMAX = 1;
constexpr int MODE_ROWMAX = 2;
constexpr int MODE_DENSE = 4;

#ifndef MPB_TWGEN
#define MPB_TWGEN 1   // pass-1 twiddles generated from one table entry (k_delta: 4.81 -> 4.63 ms per 128-signal iteration;
                      // the same trick on the shared-memory pass-2 table measured slower and is not kept)
#endif
#ifndef MPB_CORR_LEAN
#define MPB_CORR_LEAN 1    // redux.sync block maxima in k_corr for whole blocks of 256 positions
#endif
#ifndef MPB_TWGEN_CORR
#define MPB_TWGEN_CORR 0   // the same in k_corr
#endif

__device__ __forceinline__ C32 ld_stream(const C32* p) {
    float2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));   // not volatile: free to batch
    return {v.x, v.y};
}

// Programmatic dependent launch: the kernels of the iteration loop are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel may be scheduled while its predecessor in the
// stream still runs.  Each of them first lets ITS successor be scheduled (launch_dependents: takes effect once every
// CTA of this grid has issued it or exited, so the successor only fills what is left of the chip) and then waits
// until the predecessor has completed and flushed its memory (wait) before touching anything it produced -- the
// launch latency of the next kernel hides behind the current one.  Both are no-ops under an ordinary launch.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// (value, index) ordering of torch.max over a flattened map: larger value
// wins, equal values go to the lower index.
__device__ __forceinline__ void take_better(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
    }
}
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, off);
        int oi = __shfl_xor_sync(0xffffffffu, i, off);
        take_better(v, i, ov, oi);
    }
}

// Order-preserving map float -> int (for redux.sync max); -0 is folded into +0 first.
__device__ __forceinline__ int float_key(float v) {
    int k = __float_as_int(v + 0.0f);
    return k ^ ((k >> 31) & 0x7fffffff);
}
// Warp maximum of fp32 values in one instruction (redux.sync.max.f32: CREDUX.MAX.F32, sm_100a); NaN lanes are ignored.
__device__ __forceinline__ float redux_max_f32(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
// Warp (max, lowest position of the max): two redux.sync instead of a shuffle tree.
__device__ __forceinline__ void warp_argmax_redux(float& v, int& at) {
    const int key = float_key(v);
    const int kmax = __reduce_max_sync(0xffffffffu, key);
    at = __reduce_min_sync(0xffffffffu, key == kmax ? at : INT_MAX);
    v = __int_as_float(kmax ^ ((kmax >> 31) & 0x7fffffff));
}

// Rescan of one map row's block maxima outside the refreshed blocks [blk0, blk0 + nvb): folds this lane's
// share into (v, at) = (value, position).  Long signals have thousands of blocks per row (4096 at 2^20
// samples) and the winner's own row is rescanned in every iteration, so the loads are issued 16 deep per
// lane and only the values are scanned; the position of the lane's best block is fetched once at the end.
// Equal values resolve to the lowest block, i.e. the lowest position.
// bm_pos == nullptr: position-free tables (SGRAM): the candidate's "position" is the start of its block (any
// position inside the lowest block that holds the maximum orders the same; k_apply resolves the exact one).
__device__ __forceinline__ void rescan_row(const float* __restrict__ bm_val, const int* __restrict__ bm_pos, int NB,
                                           int blk0, int nvb, int lane, float& v, int& at, int blk_shift = 0) {
    float bv = -INFINITY;
    int bi = -1;
    for (int i0 = lane; i0 < NB; i0 += 32 * 16) {
        float c[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = i0 + 32 * u;
            c[u] = (i < NB && (i < blk0 || i >= blk0 + nvb)) ? bm_val[i] : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u)
            if (c[u] > bv) {                              // ascending block index inside the lane
                bv = c[u];
                bi = i0 + 32 * u;
            }
    }
    if (bi >= 0) take_better(v, at, bv, bm_pos ? bm_pos[bi] : (bi << blk_shift));
}

// ---------------------------------------------------------------------------
// y = x / (||x|| + eps) per row (modules/normalization.py:4-6). One warp per row.
// ---------------------------------------------------------------------------
// Dictionary fingerprint: two independent 64-bit lanes, each the sum over all elements of a full-avalanche mix
// (splitmix64 finaliser) of (index, float bits) under a different key -- a sum so that the grid can accumulate it
// in any order, a mix keyed on the whole index so that moving, swapping or compensating values cannot cancel.
// set_dictionary compares both lanes on the device with the previous call's and the table-building kernels below
// return at once when nothing changed (`skip`): the reference re-derives everything from `d` on every call
// (modules/matchingpursuit.py:254), and callers that pass an unchanged dictionary again should not pay for it --
// without a host synchronisation.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void k_fingerprint(const float* __restrict__ d, size_t n, unsigned long long* __restrict__ acc) {
    unsigned long long s0 = 0, s1 = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long e = ((unsigned long long)i << 32) ^ (unsigned long long)(unsigned)__float_as_int(d[i]) ^
                                     ((unsigned long long)(i >> 32) * 0x9e3779b97f4a7c15ull);
        s0 += mix64(e + 0x9e3779b97f4a7c15ull);
        s1 += mix64(~e * 0xd6e8feb86659fd93ull + 0x2545f4914f6cdd1dull);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, s0);
        atomicAdd(acc + 1, s1);
    }
}
// fp[0..1] = fingerprint just accumulated, fp[2..3] = previous one, fp[4] = previous one is valid
__global__ void k_fingerprint_decide(unsigned long long* fp, int* skip, int force) {
    *skip = (!force && fp[4] == 1ull && fp[0] == fp[2] && fp[1] == fp[3]) ? 1 : 0;
    fp[2] = fp[0];
    fp[3] = fp[1];
    fp[0] = 0ull;
    fp[1] = 0ull;
    fp[4] = 1ull;
}

__global__ void k_unit_norm(const float* __restrict__ x, float* __restrict__ y, int rows, int cols, float eps,
                            const int* __restrict__ skip = nullptr) {
    if (skip && *skip) return;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * cols;
    double acc = 0.0;
    for (int i = lane; i < cols; i += 32) {
        double v = xr[i];
        acc += v * v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    const float denom = __fadd_rn((float)sqrt(acc), eps);
    for (int i = lane; i < cols; i += 32) y[(size_t)row * cols + i] = __fdiv_rn(xr[i], denom);
}

// ---------------------------------------------------------------------------
// Spectra of atom pairs.  pair q of this plan = atoms (atom_lo + 2q, atom_lo + 2q + 1).
// ---------------------------------------------------------------------------
// LOCAL: the table is stored in the load order of BlockFft's local-first-exchange form (table_index_local).
template <int M, typename Real, bool LOCAL = false>
__global__ void __launch_bounds__(BlockFft<M, Real>::T)
k_pair_spectra(const float* __restrict__ dict, int A, int atom_lo, int atom_hi,
               const cpx<Real>* __restrict__ tw1, const cpx<Real>* __restrict__ tw2, C32* __restrict__ pairspec,
               const int* __restrict__ skip = nullptr) {
    if (skip && *skip) return;
    using F = BlockFft<M, Real>;
    using C = cpx<Real>;
    extern __shared__ __align__(16) unsigned char smraw[];
    C* sm = reinterpret_cast<C*>(smraw);
    C* stw2 = sm + F::SMEM_CPX;
    const int tl = threadIdx.x, q = blockIdx.x;
    for (int i = tl; i < 256; i += F::T) stw2[i] = tw2[i];
    const int k0 = atom_lo + 2 * q, k1 = k0 + 1;
    const float* d0 = dict + (size_t)k0 * A;
    const float* d1 = dict + (size_t)k1 * A;
    C r[F::E];
#pragma unroll
    for (int e = 0; e < F::E; ++e) {
        const int j = F::in_index(tl, e);
        Real re = (j < A) ? (Real)d0[j] : (Real)0;
        Real im = (j < A && k1 < atom_hi) ? (Real)d1[j] : (Real)0;
        r[e] = {re, im};
    }
    __syncthreads();
    F::template pass1<1>(r, tl, sm, tw1);
    __syncthreads();
    F::template pass2<1>(r, tl, sm, stw2);
    __syncthreads();
    F::template pass3<1>(r, tl, sm);
    const Real scale = (Real)1 / (Real)M;
#pragma unroll
    for (int e = 0; e < F::E; ++e) {
        int m = F::out_index(tl, e);
        if constexpr (LOCAL) m = BlockFft<M, float>::table_index_local(m);
        pairspec[(size_t)q * M + m] = {(float)(r[e].x * scale), (float)(r[e].y * scale)};
    }
}

// ---------------------------------------------------------------------------
// Forward FFT of windows of a row-major real matrix:  winspec[w] = FFT(src[row, t0 : t0+M])
// with zeros outside [0, row_len).  win may be indexed indirectly through
// `nwin_ptr` (device-side count) so the grid can be sized for the worst case.
// ---------------------------------------------------------------------------
template <int M, int STAGED = 0>
__device__ __forceinline__ void window_fft_body(const float* __restrict__ x, int row_len, int t0, int tl,
                                                C32* sm, const C32* __restrict__ tw1, const C32* stw2,
                                                C32* __restrict__ out) {
    using F = BlockFft<M, float>;
    C32 r[F::E];
#pragma unroll
    for (int e = 0; e < F::E; ++e) {
        const int t = t0 + F::in_index(tl, e);
        r[e] = {(t >= 0 && t < row_len) ? x[t] : 0.f, 0.f};
    }
    F::template pass1<-1>(r, tl, sm, tw1);
    __syncthreads();
    F::template pass2<-1>(r, tl, sm, stw2);
    __syncthreads();
    F::template pass3<-1>(r, tl, sm);
#pragma unroll
    for (int e = 0; e < F::E; ++e) {
        const int m = F::out_index(tl, e);
        int at = m;
        if constexpr (STAGED == 1) at = F::bin_addr(m);
        if constexpr (STAGED == 2) at = F::bin_addr_local(m);
        out[at] = r[e];
    }
}

// STAGED: the spectrum is written in BlockFft's pass-1 shared-memory layout (SMEM_CPX entries per window, bin j at
// bin_addr(j); 2: at bin_addr_local(j), the local-first-exchange form) so that a consumer can bulk-copy it into its
// FFT buffer as it is (k_delta).
template <int M, int STAGED = 0>
__global__ void __launch_bounds__(BlockFft<M, float>::T)
k_window_fft(const float* __restrict__ src, long long row_stride, int row_len, const Win* __restrict__ win,
             const C32* __restrict__ tw1, const C32* __restrict__ tw2, C32* __restrict__ winspec,
             const int* __restrict__ skip = nullptr) {
    if (skip && *skip) return;
    using F = BlockFft<M, float>;
    extern __shared__ __align__(16) unsigned char smraw[];
    C32* sm = reinterpret_cast<C32*>(smraw);
    C32* stw2 = sm + F::SMEM_CPX;
    const int tl = threadIdx.x, w = blockIdx.x;
    for (int i = tl; i < 256; i += F::T) stw2[i] = tw2[i];
    const Win wi = win[w];
    __syncthreads();
    window_fft_body<M, STAGED>(src + (long long)wi.row * row_stride, row_len, wi.t0, tl, sm, tw1, stw2,
                               winspec + (size_t)w * (STAGED ? F::SMEM_CPX : M));
}

// ---------------------------------------------------------------------------
// THE HOT KERNEL.  For every (window w, atom pair q):
//     y = IFFT_M( winspec[w] * pairspec[q] )          (fused complex multiply)
//     Re y[m] = <window w shifted by m, atom 2q>,  Im y[m] = <..., atom 2q+1>,  m < nvb*blk
// and then, depending on MODE,
//     BLOCKMAX  max/argmax of every block of `blk` outputs -> bm_val/bm_pos
//     ROWMAX    (only when each (row, pair) is visited by one CTA) re-reduce the
//               two touched map rows over their NB block maxima -> row_val/row_pos
//     DENSE     write the outputs to a dense destination (map, or Gram table)
// grid = (ceil(pairs / NT), window groups); a CTA keeps its pair(s) and walks
// the windows w = blockIdx.y, blockIdx.y + gridDim.y, ...
// ---------------------------------------------------------------------------
struct CorrArgs {
    const C32* winspec;
    const C32* pairspec;
    const Win* win;
    const int* nwin_ptr;  // optional device-side window count
    int nwin;
    int npairs;
    int nloc;        // atoms owned by this plan (K_owned)
    int len;         // row length N (valid outputs satisfy t0 + m < len)
    int NB;          // block-max entries per row
    int blk_shift;   // log2(blk)
    const C32* tw1;
    const C32* tw2;
    float* bm_val;
    int* bm_pos;
    float* row_val;
    int* row_pos;
    int bm_cap;      // ROWMAX kernels: capacity (in positions) of the separate output staging buffer
    float* dense;
    long long dense_row_stride;   // between windows' source rows
    long long dense_atom_stride;  // between atoms
    int dense_col_off;            // column = t0 + m + dense_col_off
    const int* skip;              // optional device flag: nothing to do (unchanged dictionary, Gram build)
    int pos_free;                 // ROWMAX: the tables carry no exact positions (SGRAM, see k_delta NOPOS)
};

// Shared memory: [tw2: 256 complex][per transform: FFT buffer SMEM_CPX complex][per transform, only
// when ROWMAX: a separate block-max staging buffer of bm_cap float2 + 64 (value, position) slots].
// The separate staging buffer lets the step kernel drop two of its CTA barriers per window (the
// FFT buffer is not reused for the outputs, so nothing has to wait for pass 3 to drain or for the
// block-max readers to finish before the next window's pass 1 stores).
#ifndef MPB_CORR_MINB
#define MPB_CORR_MINB 3      // CTAs per SM the register allocation of k_corr aims at (80 registers, 8 B spilled)
#endif
#ifndef MPB_CORR_SEP
#define MPB_CORR_SEP 0       // 1: step kernels stage their outputs in a separate shared buffer (costs the 3rd CTA)
#endif
#ifndef MPB_CORR_MINB_DENSE
#define MPB_CORR_MINB_DENSE 2   // instantiations that also write the dense map keep their outputs live longer: 128 registers (first pass of the map modes 67 -> 52 ms per 128 signals)
#endif
template <int M, int MODE>
__global__ void __launch_bounds__((BlockFft<M, float>::T < 256 ? 256 : BlockFft<M, float>::T),
                                  ((MODE & MODE_DENSE) ? MPB_CORR_MINB_DENSE : MPB_CORR_MINB))
k_corr(const CorrArgs a) {
    pdl_prologue();
    if (a.skip && *a.skip) return;
    // the FFT route inside the map modes usually has no window at all (device-side count): leave before anything is set up
    if (a.nwin_ptr && (int)blockIdx.y >= *a.nwin_ptr) return;
    using F = BlockFft<M, float>;
    constexpr int TPB = F::T < 256 ? 256 : F::T;
    constexpr int NT = TPB / F::T;   // transforms (atom pairs) per CTA
    constexpr int NW = F::T / 32;    // warps per transform
    constexpr bool SEP = MPB_CORR_SEP && (MODE & MODE_ROWMAX) != 0;
    extern __shared__ __align__(16) unsigned char smraw[];
    C32* stw2 = reinterpret_cast<C32*>(smraw);
    const int sb = threadIdx.x / F::T, tl = threadIdx.x % F::T;
    C32* sm = stw2 + 256 + (size_t)sb * F::SMEM_CPX;
    // staging of the outputs as (atom 2q, atom 2q+1) pairs, and of the refreshed block maxima
    float2* sY = SEP ? reinterpret_cast<float2*>(stw2 + 256 + (size_t)NT * F::SMEM_CPX) + (size_t)sb * (a.bm_cap + 64)
                     : reinterpret_cast<float2*>(sm);
    __shared__ float2 s_bv_static[NT * 64];
    float2* sBV = SEP ? sY + a.bm_cap : s_bv_static + sb * 64;   // [which*32 + i] = (value, position as int bits)
    for (int i = threadIdx.x; i < 256; i += TPB) stw2[i] = a.tw2[i];

    const int nwin = a.nwin_ptr ? *a.nwin_ptr : a.nwin;
    const int blk = 1 << a.blk_shift;
    const int warp = tl >> 5, lane = tl & 31;
    __syncthreads();

    // A CTA keeps a pair (group) and walks the windows; launches with fewer CTAs than pair groups along x (the FFT
    // route inside the map modes, whose window list is usually empty) walk the pair groups as well.
    for (int qg = blockIdx.x; qg * NT < a.npairs; qg += gridDim.x) {
    int q = qg * NT + sb;
    const bool q_ok = q < a.npairs;
    if (!q_ok) q = a.npairs - 1;
    const C32* __restrict__ Eq = a.pairspec + (size_t)q * M;
    for (int w = blockIdx.y; w < nwin; w += gridDim.y) {
        const Win wi = a.win[w];
        const C32* __restrict__ X = a.winspec + (size_t)w * M;
        C32 r[F::E];
#pragma unroll
        for (int e = 0; e < F::E; ++e) {
            const int j = F::in_index(tl, e);
            const float2 ev = __ldg(reinterpret_cast<const float2*>(Eq + j));
            r[e] = cmul(ld_stream(X + j), C32{ev.x, ev.y});
        }
        if constexpr (MPB_TWGEN_CORR && F::R1 >= 4) F::template pass1_gen<1>(r, tl, sm, a.tw1);
        else F::template pass1<1>(r, tl, sm, a.tw1);
        __syncthreads();
        F::template pass2<1>(r, tl, sm, stw2);
        __syncthreads();
        F::template pass3<1>(r, tl, sm);

        const int limit = min(wi.nvb * blk, a.len - wi.t0);   // valid outputs are m in [0, limit)
        if constexpr ((MODE & MODE_DENSE) != 0) {
            if (q_ok) {
                float* base = a.dense + (long long)wi.row * a.dense_row_stride + (long long)(2 * q) * a.dense_atom_stride +
                              (wi.t0 + a.dense_col_off);
                const bool second = 2 * q + 1 < a.nloc;
#pragma unroll
                for (int e = 0; e < F::E; ++e) {
                    const int m = F::out_index(tl, e);
                    if (m < limit) {
                        base[m] = r[e].x;
                        if (second) base[a.dense_atom_stride + m] = r[e].y;
                    }
                }
            }
        }
        if constexpr ((MODE & MODE_BLOCKMAX) != 0) {
            if constexpr (!SEP) __syncthreads();  // pass-3 loads finished: the FFT buffer is reused for the outputs
            const int stage_end = wi.nvb * blk;   // <= bm_cap when SEP
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                const int m = F::out_index(tl, e);
                if (m < stage_end) sY[m] = make_float2(r[e].x, r[e].y);
            }
            __syncthreads();
            const bool second = 2 * q + 1 < a.nloc;
            if constexpr (MPB_CORR_LEAN && M == 512) {
                // blocks of 16 positions, one warp per transform: TWO blocks per step (lanes 0-15 / 16-31, segmented
                // redux.sync on an order-preserving key for the value, equality + redux.sync.min for its first
                // position); lane l collects block l and the table is written coalesced at the end.  The generic
                // path below spends a 5-step shuffle tree per block and one lane's scattered stores -- at this block
                // size that was four times the transform itself (the first pass of a configs[3] band).
                const unsigned seg = lane < 16 ? 0x0000ffffu : 0xffff0000u;
                const int valid = min(limit, wi.nvb * 16);
                float keep_va = -INFINITY, keep_vb = -INFINITY;
                int keep_ia = INT_MAX, keep_ib = INT_MAX;
                for (int i2 = 0; i2 < wi.nvb; i2 += 2) {
                    const int m = i2 * 16 + lane;
                    const bool in = m < valid;
                    const float2 c = in ? sY[m] : make_float2(-INFINITY, -INFINITY);
                    const int ka = __reduce_max_sync(seg, float_key(c.x));
                    const int kb = __reduce_max_sync(seg, float_key(c.y));
                    float va = __int_as_float(ka ^ ((ka >> 31) & 0x7fffffff));
                    float vb = __int_as_float(kb ^ ((kb >> 31) & 0x7fffffff));
                    if (!(va == va)) va = -INFINITY;              // a NaN never wins
                    if (!(vb == vb)) vb = -INFINITY;
                    const int ia = __reduce_min_sync(seg, (in && c.x + 0.0f == va) ? m : INT_MAX);
                    const int ib = __reduce_min_sync(seg, (in && c.y + 0.0f == vb) ? m : INT_MAX);
                    // lanes i2 and i2 + 1 take the results of the lower / upper half-warp
                    const int src = (lane == i2 + 1) ? 16 : 0;
                    const float sva = __shfl_sync(0xffffffffu, va, src), svb = __shfl_sync(0xffffffffu, vb, src);
                    const int sia = __shfl_sync(0xffffffffu, ia, src), sib = __shfl_sync(0xffffffffu, ib, src);
                    if (lane == i2 || lane == i2 + 1) {
                        keep_va = sva; keep_vb = svb;
                        keep_ia = sia; keep_ib = sib;
                    }
                }
                if (lane < wi.nvb && q_ok) {
                    const int pa = (keep_ia == INT_MAX) ? INT_MAX : wi.t0 + keep_ia;
                    const int pb = (keep_ib == INT_MAX) ? INT_MAX : wi.t0 + keep_ib;
                    const size_t o = ((size_t)wi.row * a.nloc + 2 * q) * a.NB + wi.blk0 + lane;
                    a.bm_val[o] = keep_va;
                    a.bm_pos[o] = pa;
                    if (second) {
                        a.bm_val[o + a.NB] = keep_vb;
                        a.bm_pos[o + a.NB] = pb;
                    }
                    if constexpr ((MODE & MODE_ROWMAX) != 0) {
                        sBV[lane] = make_float2(keep_va, __int_as_float(pa));
                        sBV[32 + lane] = make_float2(keep_vb, __int_as_float(pb));
                    }
                }
            } else
            for (int i = warp; i < wi.nvb; i += NW) {
                const int hi = min((i + 1) * blk, limit);
                float va = -INFINITY, vb = -INFINITY;
                int ia = INT_MAX, ib = INT_MAX;
                if (MPB_CORR_LEAN && blk >= 32 && hi == (i + 1) * blk) {
                    // whole block of 32..256 positions: NJ = blk/32 (atom 2q, atom 2q+1) pairs per lane, value maxima
                    // by fmax + redux.sync on an order-preserving key, first position by an equality scan +
                    // redux.sync.min
                    auto lean = [&](auto njc) {
                        constexpr int NJ = decltype(njc)::value;
                        float2 c[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) c[j] = sY[i * blk + lane + 32 * j];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) { va = fmaxf(va, c[j].x); vb = fmaxf(vb, c[j].y); }
                        const int ka = __reduce_max_sync(0xffffffffu, float_key(va));
                        const int kb = __reduce_max_sync(0xffffffffu, float_key(vb));
                        va = __int_as_float(ka ^ ((ka >> 31) & 0x7fffffff));
                        vb = __int_as_float(kb ^ ((kb >> 31) & 0x7fffffff));
                        if (!(va == va)) va = -INFINITY;          // a NaN never wins
                        if (!(vb == vb)) vb = -INFINITY;
#pragma unroll
                        for (int j = NJ - 1; j >= 0; --j) {       // descending: the lowest matching position survives
                            ia = (c[j].x + 0.0f == va) ? i * blk + lane + 32 * j : ia;
                            ib = (c[j].y + 0.0f == vb) ? i * blk + lane + 32 * j : ib;
                        }
                        ia = __reduce_min_sync(0xffffffffu, ia);
                        ib = __reduce_min_sync(0xffffffffu, ib);
                    };
                    if (blk == 256) lean(std::integral_constant<int, 8>{});
                    else if (blk == 128) lean(std::integral_constant<int, 4>{});
                    else if (blk == 64) lean(std::integral_constant<int, 2>{});
                    else lean(std::integral_constant<int, 1>{});
                } else {
                    for (int m = i * blk + lane; m < hi; m += 32) {
                        const float2 c = sY[m];
                        if (c.x > va) { va = c.x; ia = m; }
                        if (c.y > vb) { vb = c.y; ib = m; }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {      // the two reductions interleave
                        const float oa = __shfl_xor_sync(0xffffffffu, va, off);
                        const int pa = __shfl_xor_sync(0xffffffffu, ia, off);
                        const float ob = __shfl_xor_sync(0xffffffffu, vb, off);
                        const int pb = __shfl_xor_sync(0xffffffffu, ib, off);
                        take_better(va, ia, oa, pa);
                        take_better(vb, ib, ob, pb);
                    }
                }
                if (lane == 0 && q_ok) {
                    const int pa = (ia == INT_MAX) ? INT_MAX : wi.t0 + ia;
                    const int pb = (ib == INT_MAX) ? INT_MAX : wi.t0 + ib;
                    const size_t o = ((size_t)wi.row * a.nloc + 2 * q) * a.NB + wi.blk0 + i;
                    a.bm_val[o] = va;
                    a.bm_pos[o] = pa;
                    if (second) {
                        a.bm_val[o + a.NB] = vb;
                        a.bm_pos[o + a.NB] = pb;
                    }
                    if constexpr ((MODE & MODE_ROWMAX) != 0) {
                        sBV[i] = make_float2(va, __int_as_float(pa));
                        sBV[32 + i] = make_float2(vb, __int_as_float(pb));
                    }
                }
            }
            if constexpr ((MODE & MODE_ROWMAX) != 0) {
                __syncthreads();  // refreshed block maxima are staged in sBV
                // warp `which` (both atoms on warp 0 when a transform has a single warp) re-derives the row
                // maximum of atom 2q + which.  If the old row maximum sat in a block this window did not
                // touch, it is still valid and only has to be compared with the refreshed blocks;
                // otherwise the whole row of block maxima is rescanned.
                for (int which = warp; which < 2; which += NW) {
                    if (!q_ok || (which == 1 && !second)) continue;
                    const size_t rowi = (size_t)wi.row * a.nloc + 2 * q + which;
                    const size_t o = rowi * a.NB;
                    const float old_v = a.row_val[rowi];
                    const int old_p = a.row_pos[rowi];
                    const int old_b = old_p >> a.blk_shift;
                    const bool old_ok = old_b < wi.blk0 || old_b >= wi.blk0 + wi.nvb;
                    float v = -INFINITY;
                    int at = INT_MAX;      // position (positions order like (block, offset), so ties resolve alike)
                    if (lane < wi.nvb) {
                        const float2 c = sBV[which * 32 + lane];
                        v = c.x;
                        at = __float_as_int(c.y);
                    }
                    if (old_ok) {
                        if (lane == 31) take_better(v, at, old_v, old_p);   // nvb <= 22 < 32: lane 31 is free
                    } else {
                        rescan_row(a.bm_val + o, a.pos_free ? nullptr : a.bm_pos + o, a.NB, wi.blk0, wi.nvb, lane, v, at,
                                   a.blk_shift);
                    }
                    warp_argmax(v, at);
                    if (lane == 0) {
                        a.row_val[rowi] = v;
                        a.row_pos[rowi] = (at == INT_MAX) ? 0 : at;
                    }
                }
            }
        }
        if constexpr (!SEP) __syncthreads();  // buffer free for the next window
    }
    }
}

// ---------------------------------------------------------------------------
// Row maxima over the block-max table (after a full pass). One warp per map row.
// ---------------------------------------------------------------------------
__global__ void k_rowmax(const float* __restrict__ bm_val, const int* __restrict__ bm_pos, int rows, int NB,
                         float* __restrict__ row_val, int* __restrict__ row_pos) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const size_t o = (size_t)row * NB;
    float v = -INFINITY;
    int at = INT_MAX;
    for (int i = lane; i < NB; i += 32) {
        const float c = bm_val[o + i];
        if (c > v) {
            v = c;
            at = i;
        }
    }
    warp_argmax(v, at);
    if (lane == 0) {
        row_val[row] = v;
        row_pos[row] = (at == INT_MAX) ? 0 : bm_pos[o + at];
    }
}

// ---------------------------------------------------------------------------
// Per-signal argmax over the owned rows -> Best record (global atom index).
// One CTA of 256 threads per signal.
// ---------------------------------------------------------------------------
__device__ __forceinline__ Best block_best(const float* __restrict__ row_val, const int* __restrict__ row_pos,
                                           int b, int nloc, int atom_lo, Best* s_best) {
    float v = -INFINITY;
    int k = INT_MAX;
    for (int i = threadIdx.x; i < nloc; i += blockDim.x) {
        const float c = row_val[(size_t)b * nloc + i];
        if (c > v) {
            v = c;
            k = i;
        }
    }
    warp_argmax(v, k);
    __shared__ float s_v[32];
    __shared__ int s_k[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_v[warp] = v;
        s_k[warp] = k;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        v = lane < nw ? s_v[lane] : -INFINITY;
        k = lane < nw ? s_k[lane] : INT_MAX;
        warp_argmax(v, k);
        if (lane == 0) {
            Best best;
            if (k == INT_MAX) {  // every candidate was NaN or -inf: fall back to the first entry
                k = 0;
                v = row_val[(size_t)b * nloc];
            }
            best.value = v;
            best.atom = atom_lo + k;
            best.position = row_pos[(size_t)b * nloc + k];
            best.pad = 0;
            *s_best = best;
        }
    }
    __syncthreads();
    return *s_best;
}

// Position-free tables (SGRAM, k_delta NOPOS): w.position is only known to lie inside the lowest block of the row
// that holds the row maximum; the exact winner is the first position of that block whose map value equals it.
// Called by a whole CTA; every thread returns the resolved record.
__device__ __forceinline__ Best resolve_position(const float* __restrict__ map_row, int N, int blk_shift, Best w) {
    __shared__ int s_first[32];
    const int t0 = (max(w.position, 0) >> blk_shift) << blk_shift;
    const int t1 = min(t0 + (1 << blk_shift), N);
    int cand = INT_MAX;
    for (int t = t0 + (int)threadIdx.x; t < t1; t += blockDim.x)
        if (map_row[t] + 0.0f == w.value) cand = min(cand, t);
    cand = __reduce_min_sync(0xffffffffu, cand);
    if ((threadIdx.x & 31) == 0) s_first[threadIdx.x >> 5] = cand;
    __syncthreads();
    cand = INT_MAX;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) cand = min(cand, s_first[i]);
    __syncthreads();
    if (cand != INT_MAX) w.position = cand;     // nothing equal (NaN / -inf rows): keep the block start
    return w;
}

__global__ void __launch_bounds__(256)
k_local_best(const float* __restrict__ row_val, const int* __restrict__ row_pos, int nloc, int atom_lo,
             Best* __restrict__ best, const float* __restrict__ map = nullptr, int NS = 0, int N = 0, int blk_shift = 0) {
    __shared__ Best s_best;
    Best r = block_best(row_val, row_pos, blockIdx.x, nloc, atom_lo, &s_best);
    if (map) r = resolve_position(map + ((size_t)blockIdx.x * nloc + (r.atom - atom_lo)) * NS, N, blk_shift, r);
    if (threadIdx.x == 0) best[blockIdx.x] = r;
}

// Reduce candidates of several ranks: cand[r*batch + b] -> winner[b].
// Larger value wins; equal values: lower atom, then lower position.
__global__ void k_reduce_best(const Best* __restrict__ cand, int n_ranks, int batch, Best* __restrict__ winner) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    Best w = cand[b];
    for (int r = 1; r < n_ranks; ++r) {
        const Best c = cand[(size_t)r * batch + b];
        if (c.value > w.value || (c.value == w.value && (c.atom < w.atom || (c.atom == w.atom && c.position < w.position))))
            w = c;
    }
    winner[b] = w;
}

// ---------------------------------------------------------------------------
// Apply one winner per signal: record it, subtract value*atom from the residual
// (two roundings, like the reference: fl(r - fl(v*d)), modules/matchingpursuit.py:305, 328;
// truncated at the right edge, :33-56), describe the +-A window whose block
// maxima are now stale, and transform that window for the next correlation.
// One CTA of 256 threads per signal.  SELECT: take the winner from this plan's
// own rows instead of `winner`.
// ---------------------------------------------------------------------------
struct ApplyArgs {
    const float* row_val;
    const int* row_pos;
    const Best* winner;   // used when !SELECT
    const float* dict;    // whole unit dictionary (K, A)
    float* residual;      // (B, N)
    int nloc, atom_lo, A, N, blk_shift, NB;
    int n_atoms;          // atoms of the whole dictionary
    int step, n_steps;
    int* atom_out;        // (B, n_steps) or null
    int* pos_out;
    float* val_out;
    Win* win;             // (B)
    const C32* tw1;
    const C32* tw2;
    C32* winspec;         // (B, M)
    int do_fft;           // 0 on the last step
    // GRAM mode: winners whose atom fits inside the signal are handed to k_gram_update through
    // `upd`; the others (atom truncated at the right edge: the Gram identity does not hold,
    // SURVEY.md A.3) take the FFT route through a compacted window list.
    int gram;
    GramUpdate* upd;      // (B)
    int* trunc_count;     // [2]: counter of this iteration at [step & 1]; the other one is reset here
    int parity;
    // atom sharding over peer memory (SELECT only, world > 1): the local winner is written into every rank's
    // mailbox and the global winner is reduced from the `world` records that arrive in ours.
    MailSlot* const* peer_mail;   // [world] mailboxes, (2, mail_batch, world) slots each
    int world, rank, mail_batch;
    unsigned seq;                 // sequence number of this exchange (never 0)
    int* xerr;                    // set when a record did not arrive in time
    const float* map;             // position-free tables (SGRAM): resident map (B, nloc, NS) the winner's exact position
    int NS;                       // is resolved from; null otherwise
    const float* raw_map;         // LCN selection: row_val/row_pos are the NORMALISED tables; the event's value is the
                                  // raw map entry at the winner (modules/matchingpursuit.py:296)
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// All-to-all of one 12-byte candidate per rank and the reduction with the reference tie-break (max value,
// then lowest atom, then lowest position), inside the kernel that applies the winner: the transfer is `world`
// 8-byte-atomic stores per word over NVLink and a poll of the local mailbox.  Called by a whole CTA.
__device__ __forceinline__ Best exchange_best(const ApplyArgs& a, int b, Best mine, Best* s_slot) {
    const unsigned long long tag = (unsigned long long)a.seq << 32;
    const size_t slot0 = ((size_t)(a.seq & 1u) * a.mail_batch + b) * a.world;
    if ((int)threadIdx.x < a.world) {
        MailSlot* dst = a.peer_mail[threadIdx.x] + slot0 + a.rank;
        st_sys_u64(&dst->w[0], tag | (unsigned)__float_as_int(mine.value));
        st_sys_u64(&dst->w[1], tag | (unsigned)mine.atom);
        st_sys_u64(&dst->w[2], tag | (unsigned)mine.position);
        // collect rank threadIdx.x's record from the local mailbox
        const MailSlot* src = a.peer_mail[a.rank] + slot0 + threadIdx.x;
        unsigned long long w0, w1, w2;
        const unsigned long long t0 = global_ns();
        bool ok = true;
        for (;;) {
            w0 = ld_sys_u64(&src->w[0]);
            w1 = ld_sys_u64(&src->w[1]);
            w2 = ld_sys_u64(&src->w[2]);
            if ((unsigned)(w0 >> 32) == a.seq && (unsigned)(w1 >> 32) == a.seq && (unsigned)(w2 >> 32) == a.seq) break;
            if (*reinterpret_cast<volatile int*>(a.xerr) != 0 || global_ns() - t0 > 20000000000ull) {   // 20 s
                ok = false;
                break;
            }
        }
        Best r;
        if (ok) {
            r.value = __int_as_float((int)(unsigned)w0);
            r.atom = (int)(unsigned)w1;
            r.position = (int)(unsigned)w2;
        } else {
            *a.xerr = 1;
            r = mine;
        }
        r.pad = 0;
        s_slot[threadIdx.x] = r;
    }
    __syncthreads();
    Best w = s_slot[0];
    for (int r = 1; r < a.world; ++r) {
        const Best c = s_slot[r];
        if (c.value > w.value || (c.value == w.value && (c.atom < w.atom || (c.atom == w.atom && c.position < w.position))))
            w = c;
    }
    return w;
}

template <int M, bool SELECT>
__global__ void __launch_bounds__(256)
k_apply(const ApplyArgs a) {
    using F = BlockFft<M, float>;
    extern __shared__ __align__(16) unsigned char smraw[];
    C32* sm = reinterpret_cast<C32*>(smraw);
    C32* stw2 = sm + F::SMEM_CPX;
    __shared__ Best s_best;
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) stw2[i] = a.tw2[i];     // constant table: before the wait
    pdl_prologue();
    Best w;
    if constexpr (SELECT) {
        w = block_best(a.row_val, a.row_pos, b, a.nloc, a.atom_lo, &s_best);
        if (a.map) w = resolve_position(a.map + ((size_t)b * a.nloc + (w.atom - a.atom_lo)) * a.NS, a.N, a.blk_shift, w);
        if (a.raw_map) {
            const int kk = min(max(w.atom - a.atom_lo, 0), a.nloc - 1), pp = min(max(w.position, 0), a.N - 1);
            w.value = a.raw_map[((size_t)b * a.nloc + kk) * a.NS + pp];
        }
        if (a.world > 1) {
            __shared__ Best s_slot[64];
            w = exchange_best(a, b, w, s_slot);
        }
    } else {
        w = a.winner[b];
    }
    // memory safety whatever the map holds (NaN/Inf inputs can leave a row without a valid maximum, and a
    // sharded caller may hand in anything): the winner is forced into the dictionary and into the signal
    w.atom = min(max(w.atom, 0), a.n_atoms - 1);
    w.position = min(max(w.position, 0), a.N - 1);
    const int p = w.position;
    if (threadIdx.x == 0 && a.atom_out) {
        a.atom_out[(size_t)b * a.n_steps + a.step] = w.atom;
        a.pos_out[(size_t)b * a.n_steps + a.step] = p;
        a.val_out[(size_t)b * a.n_steps + a.step] = w.value;
    }
    float* __restrict__ r = a.residual + (size_t)b * a.N;
    const float* __restrict__ d = a.dict + (size_t)w.atom * a.A;
    const int keep = min(a.A, a.N - p);
    for (int i = threadIdx.x; i < keep; i += blockDim.x) r[p + i] = __fsub_rn(r[p + i], __fmul_rn(w.value, d[i]));
    const int first = max(0, p - a.A + 1), last = min(a.N - 1, p + a.A - 1);
    Win wi;
    wi.row = b;
    wi.blk0 = first >> a.blk_shift;
    wi.t0 = wi.blk0 << a.blk_shift;
    wi.nvb = (last >> a.blk_shift) - wi.blk0 + 1;
    int slot = b;
    bool fft = a.do_fft != 0;
    if (a.gram) {
        __shared__ int s_slot;
        const bool truncated = p + a.A > a.N;
        if (threadIdx.x == 0) {
            GramUpdate u;
            u.value = w.value;
            u.atom = w.atom;
            u.position = p;
            u.valid = truncated ? 0 : 1;
            a.upd[b] = u;
            s_slot = truncated ? atomicAdd(a.trunc_count + a.parity, 1) : -1;
            if (b == 0) a.trunc_count[a.parity ^ 1] = 0;   // nobody touches the other counter in this launch
        }
        __syncthreads();
        slot = s_slot;
        fft = fft && truncated;
    }
    if (threadIdx.x == 0 && slot >= 0) a.win[slot] = wi;
    __syncthreads();  // residual writes of this CTA are visible to its own loads below
    if (fft) {
        const int tl = threadIdx.x;
        C32 rr[F::E];
        if (tl < F::T) {
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                const int t = wi.t0 + F::in_index(tl, e);
                rr[e] = {(t < a.N) ? r[t] : 0.f, 0.f};
            }
            F::template pass1<-1>(rr, tl, sm, a.tw1);
        }
        __syncthreads();
        if (tl < F::T) F::template pass2<-1>(rr, tl, sm, stw2);
        __syncthreads();
        if (tl < F::T) {
            F::template pass3<-1>(rr, tl, sm);
            C32* out = a.winspec + (size_t)slot * M;
#pragma unroll
            for (int e = 0; e < F::E; ++e) out[F::out_index(tl, e)] = rr[e];
        }
    }
}

// ---------------------------------------------------------------------------
// GRAM mode update.  One warp per (signal b, owned atom j):
//     map[b, j, t] -= v * G[k*, j, t - p + A - 1]      for t in the +-A window of the winner
// (fused multiply-add, one rounding), then the (max, argmax) of every touched block and of the
// whole row.  Pure streaming: 4 bytes of G read + 8 bytes of map read-modify-written per position.
// gram is (K, nloc, GS) with GS >= 2A-1; G[k, j, l] = sum_i d_k[i + l - (A-1)] * d_j[i].
// ---------------------------------------------------------------------------
struct GramArgs {
    float* map;            // (B, nloc, N)
    const float* gram;     // (K, nloc, GS)
    const GramUpdate* upd; // (B)
    float* bm_val;
    int* bm_pos;
    float* row_val;
    int* row_pos;
    int rows;              // B * nloc
    int nloc, N, NB, blk_shift, A, GS;
};

// The window is walked in chunks of 128 positions (one float4 of the map per lane; the chunk
// start is a multiple of BLK or of 128, so the map accesses are 16-byte aligned, while the Gram
// row is read with scalar loads because its offset t - p + A - 1 has arbitrary alignment).
// BLK >= 128: a block is BLK/128 chunks, reduced over the whole warp; BLK < 128: a chunk holds
// 128/BLK blocks, reduced over segments of BLK/4 lanes.  Lane i ends up holding refreshed block i.
#ifndef MPB_GRAM_MINB
#define MPB_GRAM_MINB 5      // CTAs per SM the register allocation of k_gram_update aims at (4: 0.163, 5: 0.157, 8: 0.198 ms at configs[1])
#endif
template <int BLK>
__global__ void __launch_bounds__(256, MPB_GRAM_MINB)
k_gram_update(const GramArgs a) {
    static_assert(BLK >= 16 && BLK <= 256, "block size");
    constexpr int LPB = BLK >= 128 ? 32 : BLK / 4;         // lanes that share a block inside a chunk
    constexpr int CPB = BLK >= 128 ? BLK / 128 : 1;        // chunks per block
    constexpr int BPC = BLK >= 128 ? 1 : 128 / BLK;        // blocks per chunk
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_prologue();
    if (row >= a.rows) return;
    const int b = row / a.nloc, j = row - b * a.nloc;
    const GramUpdate u = a.upd[b];
    if (!u.valid) return;
    const float old_v = a.row_val[row];      // needed at the very end: requested now
    const int old_p = a.row_pos[row];
    const int p = u.position;
    const float nv = -u.value;
    const int first = max(0, p - a.A + 1), last = min(a.N - 1, p + a.A - 1);
    const int blk0 = first / BLK, nvb = last / BLK - blk0 + 1;   // nvb <= 32
    const float* __restrict__ g = a.gram + ((size_t)u.atom * a.nloc + j) * a.GS + (a.A - 1 - p);   // g[t]
    float* __restrict__ m = a.map + (size_t)row * a.N;
    const size_t bm0 = (size_t)row * a.NB;
    const bool vec_ok = (a.N % 4 == 0);
    float my_v = -INFINITY;
    int my_p = INT_MAX;
    float carry_v = -INFINITY;   // BLK > 128: maximum of the block's earlier chunks
    int carry_p = INT_MAX;
    const int c0 = (blk0 * BLK) / 128 * 128;                  // first chunk start (multiple of 128, <= blk0*BLK)
    const int c_end = (blk0 + nvb) * BLK;
    // The loop is software pipelined one chunk deep: the map and Gram values of chunk c+1 are requested
    // before chunk c is reduced, so two chunks' worth of bytes per warp are in flight (the kernel is a pure
    // stream and was latency bound: long_scoreboard 15.7 per issue at 71 % of the HBM peak).
    auto fetch = [&](int tc, float* x, float* gv) {
        const int t = tc + 4 * lane;
        if (vec_ok && t + 3 < a.N) {
            const float4 q = *reinterpret_cast<const float4*>(m + t);
            x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (t + e < a.N) ? m[t + e] : -INFINITY;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int tt = t + e;
            gv[e] = (tt >= first && tt <= last) ? __ldg(g + tt) : 0.f;
        }
    };
    float xn[4], gn[4];
    fetch(c0, xn, gn);
    for (int tc = c0; tc < c_end; tc += 128) {
        const int t = tc + 4 * lane;
        float x[4], gv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { x[e] = xn[e]; gv[e] = gn[e]; }
        if (tc + 128 < c_end) fetch(tc + 128, xn, gn);
        bool touched = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int tt = t + e;
            if (tt >= first && tt <= last) {
                x[e] = fmaf(nv, gv[e], x[e]);
                touched = true;
            }
        }
        if (touched) {
            if (vec_ok && t + 3 < a.N) {
                *reinterpret_cast<float4*>(m + t) = make_float4(x[0], x[1], x[2], x[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (t + e < a.N) m[t + e] = x[e];
            }
        }
        float best = -INFINITY;
        int at = INT_MAX;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (x[e] > best) {
                best = x[e];
                at = t + e;
            }
#pragma unroll
        for (int off = LPB / 2; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, at, off);
            take_better(best, at, ov, oi);
        }
        if constexpr (CPB > 1) {
            take_better(best, at, carry_v, carry_p);          // earlier chunk: lower positions win ties
            const bool block_done = ((tc + 128) % BLK) == 0;
            carry_v = block_done ? -INFINITY : best;
            carry_p = block_done ? INT_MAX : at;
            if (!block_done) continue;
        }
        // every lane of a segment now holds its block's (max, argmax); hand block (blk0 + lane) to `lane`
        const int first_blk = (BLK >= 128) ? tc / BLK : tc / BLK;            // first block of this chunk
        const int want = blk0 + lane;                                          // block this lane keeps
        const int rel = want - first_blk;
        const bool mine = rel >= 0 && rel < BPC && lane < nvb;
        const int src = mine ? rel * LPB : 0;
        const float sv = __shfl_sync(0xffffffffu, best, src);
        const int sp = __shfl_sync(0xffffffffu, at, src);
        if (mine) {
            my_v = sv;
            my_p = sp;
        }
        // segment leaders publish their block
        if ((lane % LPB) == 0) {
            const int bi = first_blk + lane / LPB;
            if (bi >= blk0 && bi < blk0 + nvb) {
                a.bm_val[bm0 + bi] = best;
                a.bm_pos[bm0 + bi] = at;
            }
        }
    }
    // row maximum: the refreshed blocks sit in registers (lane i < nvb holds block blk0 + i).  If the old row
    // maximum lies outside the window it is still valid and only has to be compared with them; otherwise
    // the row's other block maxima are rescanned (16 loads in flight per lane, rescan_row).
    const int old_b = old_p / BLK;
    float v = lane < nvb ? my_v : -INFINITY;
    int at = lane < nvb ? my_p : INT_MAX;
    if (old_b >= blk0 && old_b < blk0 + nvb) rescan_row(a.bm_val + bm0, a.bm_pos + bm0, a.NB, blk0, nvb, lane, v, at);
    else if (lane == 0) take_better(v, at, old_v, old_p);
    warp_argmax(v, at);
    if (lane == 0) {
        a.row_val[row] = v;
        a.row_pos[row] = (at == INT_MAX) ? 0 : at;
    }
}

// ---------------------------------------------------------------------------
// Asynchronous bulk copies (TMA, 1-D form) and the mbarrier that tracks them.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------
// SPECTRAL-GRAM mode update ("the Gram row is synthesised, not stored").  For signal b with winner
// (k*, p, v) and atom pair q the cross-correlations
//     G[k*, 2q, l] + i G[k*, 2q+1, l] = IFFT_M2( atomspec[k*] * pairspec2[q] )[l + A - 1],   |l| < A
// come out of ONE inverse transform of M2 >= 2A points -- half the length the windowed
// re-correlation needs, and every output is used -- and are applied to the resident map rows
//     map[b, 2q(+1), t] -= v * G[k*, 2q(+1), t - p]          (fused multiply-add)
// followed by the block maxima of the touched blocks and the row maxima, as in k_corr.
//
// The map window (whole blocks: <= bm_cap positions of both rows) is STAGED in shared memory by
// bulk asynchronous copies: the loads are issued before the transform and land while it runs, the
// update is a conflict-free shared-memory read-modify-write straight from the FFT registers
// (thread tl owns outputs tl + T*u + (M2/16)*m3, i.e. consecutive lanes touch consecutive
// positions), and the rows go back with bulk stores that overlap the block-max phase and the next
// signal's transform.  No register ever waits on HBM.
//
// Persistent grid (occupancy x SM count CTAs); work items are (pair group, signal) in pair-major order.
// ---------------------------------------------------------------------------
struct DeltaArgs {
    const C32* atomspec;    // (K, spec_stride): forward spectrum of [0^(A-1), d_k]; MPB_DELTA_SPREF: in BlockFft's
                            // staged layout (spec_stride = SMEM_CPX), else natural order (spec_stride = M2)
    const C32* pairspec2;   // (npairs, M2): inverse-kernel spectrum of d[2q] + i d[2q+1], scaled 1/M2
    const GramUpdate* upd;  // (B)
    int batch, npairs, nloc;
    int ngroups;            // ceil(npairs / transforms per CTA)
    float* map;             // (B, nloc, NS)
    int N, NS, NB, blk_shift, A;
    int cap;                // staged positions per row (bm_cap: whole blocks covering any +-A window)
    const C32* tw1;         // BlockFft<M2> tables
    const C32* tw2;
    float* bm_val;
    int* bm_pos;
    float* row_val;
    int* row_pos;
};

#ifndef MPB_DELTA_MINB
#define MPB_DELTA_MINB 3
#endif
#ifndef MPB_DELTA_NOPOS
#define MPB_DELTA_NOPOS 1      // position-free block/row tables in SGRAM mode (see k_delta's NOPOS)
#endif
#ifndef MPB_DELTA_NOPOS_MIN_ITEMS
#define MPB_DELTA_NOPOS_MIN_ITEMS 1024     // (atom-sharded single signal: 1024 items per rank of configs[4]: 47.6 -> 45.8 us per iteration; 8192 items: 207 -> 204)
#endif
#ifndef MPB_DELTA_DEFER
#define MPB_DELTA_DEFER 1      // three CTA barriers per item instead of four: the row maxima of an item are re-derived
                               // one item later, in the segment where the other warps reduce block maxima (of which the
                               // row warps then take fewer), and the map window is requested after the first barrier
#endif
#ifndef MPB_DELTA_SPREF
#define MPB_DELTA_SPREF 1      // the next item's winner spectrum is bulk-copied (TMA) into the idle FFT buffer while the
                               // block/row maxima of the current item are reduced; the product phase then reads it from
                               // shared memory instead of L2
#endif
#ifndef MPB_DELTA_FREDUX
#define MPB_DELTA_FREDUX 1     // block maxima with redux.sync.max.f32 (CREDUX.MAX.F32, sm_100a) instead of an integer key
#endif
#ifndef MPB_DELTA_TRIMST
#define MPB_DELTA_TRIMST 1     // only the updated range of the staged rows is stored back (16-byte granules)
#endif
#ifndef MPB_DELTA_DB
#define MPB_DELTA_DB 1         // 4096-point transforms only: TWO CTAs per SM with 128 registers -- the pair spectrum stays in
                               // registers for the run of items that share it (no 32 KB L2 read per item) -- and a
                               // DOUBLE-BUFFERED map window: the next item's window is requested right after this item's
                               // stores are issued, into the other buffer, so neither the store drain nor the HBM latency
                               // of the window sits between two items of a CTA
#endif
#ifndef MPB_DELTA_E
#define MPB_DELTA_E 0          // complex values per thread of k_delta's transform (0: BlockFft's default, 16 up to 4096 points)
#endif
#ifndef MPB_DELTA_TPB
#define MPB_DELTA_TPB 256      // minimum CTA size (several transforms share a CTA when one needs fewer threads)
#endif
constexpr int delta_elems(int m2, int want) {      // `want` values per thread if that leaves >= 32 threads and divides evenly
    return (want > 0 && m2 / (want > 0 ? want : 1) >= 32 && want % (m2 / 256) == 0) ? want : 0;
}
#ifndef MPB_DELTA_LOCAL
#define MPB_DELTA_LOCAL 0      // 1: 4096-point transforms in BlockFft's local-first-exchange form (the pass-1 -> pass-2
                               // exchange stays inside a half-warp, __syncwarp; one CTA barrier per transform instead of
                               // two).  Correct (parity suite green) but measured 8.78 against 8.71 ms per 256 signals:
                               // the warps wait at the remaining barrier instead.  Kept as a measured variant.
#endif
#ifndef MPB_DELTA_DB_MIN_BATCH
#define MPB_DELTA_DB_MIN_BATCH 32   // resident signals from which the host launches the DB form (the run of items per
                                    // pair spectrum is the batch; one signal of 8192 items: 204 us without, 238 us with)
#endif
template <int M2, bool DBT = false>
struct DeltaCfg {
    using F = BlockFft<M2, float, delta_elems(M2, MPB_DELTA_E)>;
    static constexpr int TPB = F::T < MPB_DELTA_TPB ? MPB_DELTA_TPB : F::T;
    // the spectra tables are stored in that form's order, so producers and consumer must agree (plan build vs k_delta)
    static constexpr bool LOCAL = MPB_DELTA_LOCAL && MPB_DELTA_SPREF && F::HAS_LOCAL && F::T == 256;
    static constexpr bool DB = DBT && MPB_DELTA_DB && MPB_DELTA_SPREF && !LOCAL && M2 == 4096 && TPB == F::T;
    static constexpr int MINB = DB ? 2 : MPB_DELTA_MINB;
    static constexpr int NSTAGE = DB ? 2 : 1;          // staging buffers (of two rows) per transform
};
// NOPOS: the block and row tables carry no exact positions -- a candidate's "position" is the start of its block
// (blocks order like positions, so ties resolve alike) and k_apply finds the first position of the maximum inside
// that block of the resident map.  Saves the position half of every block reduction (whole blocks of 128/256).
template <int M2, bool NOPOS, bool DBT = false>
__global__ void __launch_bounds__((DeltaCfg<M2, DBT>::TPB), (DeltaCfg<M2, DBT>::MINB))
k_delta(const DeltaArgs a) {
    using F = typename DeltaCfg<M2, DBT>::F;
    constexpr int TPB = DeltaCfg<M2, DBT>::TPB;
    constexpr int NT = TPB / F::T;
    constexpr int NW = F::T / 32;
    constexpr int RW0 = NW >= 4 ? NW - 2 : 0;            // first of the warps that re-derive the row maxima
    constexpr int NROW = NW >= 2 ? 1 : 2;                // rows per such warp
    extern __shared__ __align__(16) unsigned char smraw[];
    C32* stw2 = reinterpret_cast<C32*>(smraw);
    const int sb = threadIdx.x / F::T, tl = threadIdx.x % F::T;
    C32* sm = stw2 + 256 + (size_t)sb * F::SMEM_CPX;
    constexpr bool DB = DeltaCfg<M2, DBT>::DB;
    float* const stage = reinterpret_cast<float*>(stw2 + 256 + (size_t)NT * F::SMEM_CPX) +
                         (size_t)sb * 2 * DeltaCfg<M2, DBT>::NSTAGE * a.cap;
    float* st0 = stage;                                  // DB: re-aimed at the item's buffer at every item top
    float* st1 = st0 + a.cap;
    constexpr bool DEFER = MPB_DELTA_DEFER && NW == 8;   // needs dedicated row warps (M2 >= 4096)
    constexpr bool LOCAL = DeltaCfg<M2>::LOCAL;
    __shared__ float2 s_bv_static[NT * 64 * (DEFER ? 2 : 1)];
    __shared__ __align__(8) unsigned long long s_bar[3 * NT];
    float2* sBV = s_bv_static + sb * 64 * (DEFER ? 2 : 1);   // [which*32 + block] = (value, position as int bits); DEFER: x2 (item parity)
    unsigned long long* bar = s_bar + sb;                // map window landed
    unsigned long long* barS = s_bar + NT + sb;          // winner spectrum landed (MPB_DELTA_SPREF)
    unsigned long long* barF = s_bar + 2 * NT + sb;      // DEFER: every warp has left the block reduction (staging rows free)
    constexpr unsigned SBYTES = (unsigned)(F::SMEM_CPX * sizeof(C32));
    for (int i = threadIdx.x; i < 256; i += TPB) stw2[i] = a.tw2[i];
    if (tl == 0) {
        mbar_init(bar, 1);
        mbar_init(barS, 1);
        mbar_init(barF, NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    const int blk = 1 << a.blk_shift;
    const int warp = tl >> 5, lane = tl & 31;
    const int which0 = warp - RW0;                       // row (0/1) this warp re-derives; outside [0, 2): none
    unsigned phase = 0;
    pdl_prologue();                                      // the table above is constant; everything below is not
    __syncthreads();

    // Work items (pair group g, signal b), g-major so that a CTA keeps its pair spectra hot; the
    // grid is persistent and every CTA takes an equal contiguous share (host: ngroups * batch < 2^31).
    const long long total = (long long)a.ngroups * a.batch;
    int n_items = (int)(total * (blockIdx.x + 1) / gridDim.x - total * blockIdx.x / gridDim.x);
    int g, b;
    {
        const long long it0 = total * blockIdx.x / gridDim.x;
        g = (int)(it0 / a.batch);
        b = (int)(it0 - (long long)g * a.batch);
    }
    C32 r[F::E];
    unsigned phaseS = 0;
    // tl == 0: request signal bb's winner spectrum into this transform's (idle) FFT buffer
    auto request_spectrum = [&](int bb) {
        const GramUpdate un = a.upd[bb];
        if (un.valid) {
            mbar_expect_tx(barS, SBYTES);
            bulk_load(sm, a.atomspec + (size_t)un.atom * F::SMEM_CPX, SBYTES, barS);
        }
    };
    if (MPB_DELTA_SPREF && tl == 0 && n_items > 0) request_spectrum(b);
    // DEFER: the item whose row maxima are still to be re-derived (by the row warps, one item later)
    int pv_row = -1;                                     // map row of its atom 2q; -1: none
    int pv_win = 0;                                      // blk0 << 6 | nvb << 1 | second
    int par = 0;                                         // which half of sBV the current item writes
    unsigned phaseF = 0;
    bool phaseF_armed = false;                           // a block reduction has been run (barF has a phase to wait for)
    // Row maxima of the (row, which) pair from the refreshed block maxima in `bv` and the row's previous maximum.
    auto row_phase = [&](int row0, int which, int rblk0, int rnvb, const float2* bv, float oldv, int oldp) {
        const size_t rowi = (size_t)row0 + which;
        const size_t o = rowi * a.NB;
        float v = -INFINITY;
        int at = INT_MAX;
        if (lane < rnvb) {
            const float2 c = bv[which * 32 + lane];
            v = c.x;
            at = __float_as_int(c.y);
            a.bm_val[o + rblk0 + lane] = v;              // coalesced publication of the refreshed blocks
            if constexpr (!NOPOS) a.bm_pos[o + rblk0 + lane] = at;
        }
        const int old_b = oldp >> a.blk_shift;
        const bool old_ok = old_b < rblk0 || old_b >= rblk0 + rnvb;
        if (old_ok) {
            if (lane == 31) take_better(v, at, oldv, oldp);      // nvb <= 30: lane 31 is free
        } else {
            rescan_row(a.bm_val + o, NOPOS ? nullptr : a.bm_pos + o, a.NB, rblk0, rnvb, lane, v, at, a.blk_shift);
        }
        warp_argmax(v, at);
        if (lane == 0) {
            a.row_val[rowi] = v;
            a.row_pos[rowi] = (at == INT_MAX) ? 0 : at;
        }
    };
    // DB: tl == 0 requests the map window of item (pair qq, signal bb, winner position pos) into staging buffer `which_buf`
    auto request_window_into = [&](int qq, int bb, int pos, int which_buf) {
        const int wfirst = max(0, pos - a.A + 1), wlast = pos + a.A - 1;
        const int wblk0 = wfirst >> a.blk_shift, wnvb = (wlast >> a.blk_shift) - wblk0 + 1;
        const int wstart = wblk0 << a.blk_shift;
        const unsigned bytes = (unsigned)min(wnvb << a.blk_shift, a.NS - wstart) * 4u;
        const float* w0 = a.map + ((size_t)bb * a.nloc + 2 * qq) * a.NS + wstart;
        float* dst = stage + (size_t)which_buf * 2 * a.cap;
        const bool wsecond = 2 * qq + 1 < a.nloc;
        mbar_expect_tx(bar, wsecond ? 2u * bytes : bytes);
        bulk_load(dst, w0, bytes, bar);
        if (wsecond) bulk_load(dst + a.cap, w0 + a.NS, bytes, bar);
    };
    int wpar = 0;                                        // DB: staging buffer of the current item
    bool win_pre = false;                                // DB, tl == 0: the current item's window was requested by the previous one
    int eq_q = -1;                                       // DB: the pair whose spectrum eqr[] holds
    C32 eqr[DB ? F::E : 1];
    for (; n_items > 0; --n_items, b = (b + 1 == a.batch ? 0 : b + 1), g += (b == 0)) {
        const GramUpdate u = a.upd[b];
        const int b_next = (b + 1 == a.batch ? 0 : b + 1);
        if constexpr (DB) {
            st0 = stage + (size_t)wpar * 2 * a.cap;
            st1 = st0 + a.cap;
        }
        if (!u.valid) {                                  // CTA-uniform: this signal takes the FFT route
            if (MPB_DELTA_SPREF && tl == 0 && n_items > 1) request_spectrum(b_next);   // nobody touches the buffer now
            continue;
        }
        int q = g * NT + sb;
        const bool q_ok = q < a.npairs;
        if (!q_ok) q = a.npairs - 1;
        const bool second = 2 * q + 1 < a.nloc;
        const int p = u.position;
        const float nv = -u.value;
        const int first = max(0, p - a.A + 1), last = p + a.A - 1;     // valid => p + A <= N
        const int blk0 = first >> a.blk_shift, nvb = (last >> a.blk_shift) - blk0 + 1;
        const int start = blk0 << a.blk_shift;
        const int cnt = min(nvb << a.blk_shift, a.NS - start);          // staged floats per row (multiple of 4)
        float* __restrict__ m0 = a.map + ((size_t)b * a.nloc + 2 * q) * a.NS + start;
        float* __restrict__ m1 = m0 + a.NS;
        auto request_window = [&]() {                    // tl == 0
            bulk_wait_read0();                           // the previous item's stores have left the staging rows
            const unsigned bytes = (unsigned)cnt * 4u;
            mbar_expect_tx(bar, second ? 2u * bytes : bytes);
            bulk_load(st0, m0, bytes, bar);
            if (second) bulk_load(st1, m1, bytes, bar);
        };
        if constexpr (DEFER) {
            // the staging rows are free once every warp has left the previous item's block reduction (each warp
            // arrives on barF there): usually long ago, so the window is requested right away and lands behind all
            // three passes
            if (tl == 0) {
                if (phaseF_armed) {
                    while (!mbar_try_wait(barF, phaseF)) {}
                    phaseF ^= 1u;
                }
                if (q_ok && !(DB && win_pre)) request_window();
                win_pre = false;
            }
        } else if (tl == 0 && q_ok) request_window();
        // the old row maxima are fetched now so that the row phase never waits on memory (DEFER: those of the
        // PREVIOUS item, whose row phase runs during this one)
        float old_v[NROW];
        int old_p[NROW];
#pragma unroll
        for (int i = 0; i < NROW; ++i) {
            const int which = which0 + i;
            bool mine;
            size_t rowi;
            if constexpr (DEFER) {
                mine = pv_row >= 0 && which >= 0 && which < 2 && (which == 0 || (pv_win & 1));
                rowi = (size_t)(mine ? pv_row + which : 0);
            } else {
                mine = q_ok && which >= 0 && which < 2 && (which == 0 || second);
                rowi = (size_t)b * a.nloc + 2 * q + (mine ? which : 0);
            }
            old_v[i] = mine ? __ldg(a.row_val + rowi) : 0.f;
            old_p[i] = mine ? __ldg(a.row_pos + rowi) : 0;
        }
        {
            // Both spectra are L2 resident (the pair spectrum is re-read by this CTA for every signal,
            // the winner spectrum by every CTA).  Measured dead ends, for the record: requesting the next
            // item's winner spectrum into the dead FFT registers before the block-max phase (spills:
            // 5.35 -> 5.99 ms per iteration of 128 signals), and parking the pair spectrum in tensor
            // memory with tcgen05.st/ld.32x32b (correct, but 14.7 ms).
            const C32* __restrict__ Eq = a.pairspec2 + (size_t)q * M2;
            if constexpr (MPB_DELTA_SPREF) {
                // the pair spectrum comes from L2 while the winner spectrum -- requested during the previous item's
                // reduction phase -- is read from this thread's own slots of the FFT buffer (the slots pass 1
                // overwrites below: no barrier needed)
                if constexpr (DB) {
                    if (eq_q != q) {                     // a CTA walks the whole batch per pair group: once per ~batch items
#pragma unroll
                        for (int e = 0; e < F::E; ++e) {
                            const float2 y = __ldg(reinterpret_cast<const float2*>(Eq + F::in_index(tl, e)));
                            eqr[e] = C32{y.x, y.y};
                        }
                        eq_q = q;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < F::E; ++e) {
                        const float2 y = __ldg(reinterpret_cast<const float2*>(Eq + F::in_index(tl, e)));
                        r[e] = C32{y.x, y.y};
                    }
                }
                while (!mbar_try_wait(barS, phaseS)) {}
                phaseS ^= 1u;
#pragma unroll
                for (int e = 0; e < F::E; ++e) {
                    if constexpr (DB) r[e] = cmul(sm[F::slot_addr(tl, e)], eqr[e]);
                    else if constexpr (LOCAL) r[e] = cmul(sm[F::slot_addr_local(tl, e)], r[e]);
                    else r[e] = cmul(sm[F::slot_addr(tl, e)], r[e]);
                }
            } else {
                const C32* __restrict__ S = a.atomspec + (size_t)u.atom * M2;
#pragma unroll
                for (int e = 0; e < F::E; ++e) {
                    const int j = F::in_index(tl, e);
                    const float2 y = __ldg(reinterpret_cast<const float2*>(Eq + j));
                    const float2 x = __ldg(reinterpret_cast<const float2*>(S + j));
                    r[e] = cmul(C32{x.x, x.y}, C32{y.x, y.y});
                }
            }
        }
        if constexpr (LOCAL) {
            F::template pass1_local<1>(r, tl, sm, a.tw1);
            __syncwarp();                                // the first exchange stays inside each half-warp
            F::template pass2_local<1>(r, tl, sm, stw2);
            __syncthreads();
            F::template pass3_local<1>(r, tl, sm);
        } else {
            if constexpr (MPB_TWGEN && F::R1 >= 4) F::template pass1_gen<1>(r, tl, sm, a.tw1);
            else F::template pass1<1>(r, tl, sm, a.tw1);
            __syncthreads();
            F::template pass2<1>(r, tl, sm, stw2);
            __syncthreads();
            F::template pass3<1>(r, tl, sm);
        }

        if (q_ok) {
            while (!mbar_try_wait(bar, phase)) {}
            phase ^= 1u;
            // output m is lag m - (A-1), i.e. position t = p - (A-1) + m, staged at index i = m + off.
            // Valid outputs are m < 2A-1 with t >= 0:  i in [lo_i, hi_i).  The staging rows hold
            // cap >= M2 + blk floats, so when off >= 0 every index is in bounds and only the store
            // has to be predicated.
            const int off = p - (a.A - 1) - start;
            const int lo_i = max(off, 0), hi_i = off + 2 * a.A - 1;
            const unsigned span = (unsigned)(hi_i - lo_i);
            const int i0 = F::out_index(tl, 0) + off;
            if (off >= 0 && 2 * a.A >= M2) {             // power-of-two atoms away from the left edge: only output M2-1 is void
#pragma unroll
                for (int e = 0; e < F::E; ++e) {
                    const int i = i0 + (F::out_index(0, e) - F::out_index(0, 0));
#if MPB_F32X2
                    // both rows in one packed fused multiply-add (the operands of the pair are two 4-byte loads)
                    const float2 x = __ffma2_rn(make_float2(nv, nv), make_float2(r[e].x, r[e].y), make_float2(st0[i], st1[i]));
                    const float x0 = x.x, x1 = x.y;
#else
                    const float x0 = fmaf(nv, r[e].x, st0[i]);
                    const float x1 = fmaf(nv, r[e].y, st1[i]);
#endif
                    if (F::out_index(F::T - 1, e) != M2 - 1 || tl != F::T - 1) {
                        st0[i] = x0;
                        st1[i] = x1;                     // row 1 of a last odd pair is staged garbage, never stored back
                    }
                }
            } else if (off >= 0) {
#pragma unroll
                for (int e = 0; e < F::E; ++e) {
                    const int i = i0 + (F::out_index(0, e) - F::out_index(0, 0));
                    const float x0 = fmaf(nv, r[e].x, st0[i]);
                    const float x1 = fmaf(nv, r[e].y, st1[i]);
                    if ((unsigned)(i - lo_i) < span) {
                        st0[i] = x0;
                        st1[i] = x1;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < F::E; ++e) {
                    const int i = i0 + (F::out_index(0, e) - F::out_index(0, 0));
                    if ((unsigned)(i - lo_i) < span) {
                        st0[i] = fmaf(nv, r[e].x, st0[i]);
                        st1[i] = fmaf(nv, r[e].y, st1[i]);
                    }
                }
            }
        }
        fence_proxy_async();                             // generic-proxy writes -> visible to the bulk stores; the FFT
                                                         // buffer's generic writes are ordered before its bulk refill
        __syncthreads();                                 // rows updated; FFT buffer free
        if (tl == 0 && q_ok) {
            if constexpr (MPB_DELTA_TRIMST) {
                // only the updated range [lo_i, hi_i) goes back, widened to 16-byte granules (the rest of the staged
                // blocks is unchanged and only needed by the block reduction)
                const int off = p - (a.A - 1) - start;
                const int s_lo = max(off, 0) & ~3;
                const int s_hi = min((off + 2 * a.A - 1 + 3) & ~3, cnt);
                const unsigned bytes = (unsigned)(s_hi - s_lo) * 4u;
                bulk_store(m0 + s_lo, st0 + s_lo, bytes);
                if (second) bulk_store(m1 + s_lo, st1 + s_lo, bytes);
            } else {
                bulk_store(m0, st0, (unsigned)cnt * 4u);
                if (second) bulk_store(m1, st1, (unsigned)cnt * 4u);
            }
            bulk_commit();
        }
        if constexpr (DB) {
            // the NEXT item's window goes into the other staging buffer now: its last reader was the block reduction
            // of the previous item (every warp has passed three barriers since) and the bulk store of the previous
            // item, which `wait_group.read 1` -- all but the group just committed -- has seen leave
            if (tl == 0 && n_items > 1) {
                const GramUpdate un = a.upd[b_next];
                const int qn = (g + (b_next == 0 ? 1 : 0)) * NT + sb;
                if (un.valid && qn < a.npairs) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    request_window_into(qn, b_next, un.position, wpar ^ 1);
                    win_pre = true;
                }
            }
        }
        if (MPB_DELTA_SPREF && tl == 0 && n_items > 1) request_spectrum(b_next);
        if constexpr (DEFER) {
            // the previous item's row maxima: its block maxima were staged before this item's barriers
            if (pv_row >= 0 && which0 >= 0 && which0 < 2 && (which0 == 0 || (pv_win & 1)))
                row_phase(pv_row, which0, pv_win >> 6, (pv_win >> 1) & 31, sBV + (par ^ 1) * 64, old_v[0], old_p[0]);
        }
        float2* sBVw = sBV + (DEFER ? par * 64 : 0);
        if (q_ok) {
            // (row, block) tasks are dealt round-robin to the warps; the refreshed (max, position) pairs
            // only go to shared memory here -- the row warps publish them to bm_val / bm_pos.
            const int ntask = second ? 2 * nvb : nvb;
            if (blk >= 128 && start + (nvb << a.blk_shift) <= a.N) {   // whole blocks of 128 / 256: one / two float4 per lane
                // A task is one block index of BOTH rows (they share every address but the row offset): two (blk 128)
                // or four (blk 256) 16-byte loads, an fmax tree per row and one redux.sync.max per row on an
                // order-preserving key.  DEFER: the two row warps have just re-derived a row maximum each and take the
                // last (up to) two block indices, the other warps share the rest.
                const float* __restrict__ rowl = st0 + 4 * lane;
                const int sh = a.blk_shift;
                const int nrt = DEFER ? min(nvb, 2) : 0;
                const int t_first = !DEFER ? warp : (warp < RW0 ? warp : nvb - nrt + (warp - RW0));
                const int t_end = !DEFER ? nvb : (warp < RW0 ? nvb - nrt : min(nvb, t_first + 1));
                const int t_step = !DEFER ? NW : (warp < RW0 ? RW0 : 1);
                for (int i = t_first; i < t_end; i += t_step) {
                    const float* __restrict__ p0 = rowl + (i << sh);
                    const float4 c00 = *reinterpret_cast<const float4*>(p0);
                    const float4 c10 = *reinterpret_cast<const float4*>(p0 + a.cap);
                    float4 c01 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), c11 = c01;
                    if (blk == 256) {
                        c01 = *reinterpret_cast<const float4*>(p0 + 128);
                        c11 = *reinterpret_cast<const float4*>(p0 + a.cap + 128);
                    }
                    float v0 = fmaxf(fmaxf(fmaxf(c00.x, c00.y), fmaxf(c00.z, c00.w)),
                                     fmaxf(fmaxf(c01.x, c01.y), fmaxf(c01.z, c01.w)));
                    float v1 = fmaxf(fmaxf(fmaxf(c10.x, c10.y), fmaxf(c10.z, c10.w)),
                                     fmaxf(fmaxf(c11.x, c11.y), fmaxf(c11.z, c11.w)));
                    if constexpr (MPB_DELTA_FREDUX && NOPOS) {
                        // NaN operands are ignored by max.f32 / redux.max.f32; only a block of NaNs yields NaN
                        v0 = redux_max_f32(v0) + 0.0f;            // -0 folds into +0, as the key form does
                        v1 = redux_max_f32(v1) + 0.0f;
                    } else {
                        const int k0 = __reduce_max_sync(0xffffffffu, float_key(v0));
                        const int k1 = __reduce_max_sync(0xffffffffu, float_key(v1));
                        v0 = __int_as_float(k0 ^ ((k0 >> 31) & 0x7fffffff));
                        v1 = __int_as_float(k1 ^ ((k1 >> 31) & 0x7fffffff));
                    }
                    if (!(v0 == v0)) v0 = -INFINITY;              // a NaN in the map never wins (as in the comparison-based paths)
                    if (!(v1 == v1)) v1 = -INFINITY;
                    const int pos0 = start + (i << sh);           // NOPOS: the block's start stands for the position
                    if constexpr (NOPOS) {
                        if (lane == 0) {
                            sBVw[i] = make_float2(v0, __int_as_float(pos0));
                            sBVw[32 + i] = make_float2(v1, __int_as_float(pos0));
                        }
                    } else {
                        // first position of the maximum: descending order, so the lowest match survives
                        auto first_at = [&](const float4& lo4, const float4& hi4, float v) {
                            int at = INT_MAX;
                            at = (hi4.w + 0.0f == v) ? 131 : at;
                            at = (hi4.z + 0.0f == v) ? 130 : at;
                            at = (hi4.y + 0.0f == v) ? 129 : at;
                            at = (hi4.x + 0.0f == v) ? 128 : at;
                            at = (lo4.w + 0.0f == v) ? 3 : at;
                            at = (lo4.z + 0.0f == v) ? 2 : at;
                            at = (lo4.y + 0.0f == v) ? 1 : at;
                            at = (lo4.x + 0.0f == v) ? 0 : at;
                            return __reduce_min_sync(0xffffffffu, at == INT_MAX ? at : at + 4 * lane);
                        };
                        const int at0 = first_at(c00, c01, v0), at1 = first_at(c10, c11, v1);
                        if (lane == 0) {
                            sBVw[i] = make_float2(v0, __int_as_float(at0 == INT_MAX ? INT_MAX : pos0 + at0));
                            sBVw[32 + i] = make_float2(v1, __int_as_float(at1 == INT_MAX ? INT_MAX : pos0 + at1));
                        }
                    }
                }
            } else if (blk <= 64 && start + (nvb << a.blk_shift) <= a.N) {
                // short atoms (blocks of 16 / 32 / 64 positions, whole blocks inside the signal): a task is 64
                // consecutive staged positions of one row = 4 / 2 / 1 blocks, two conflict-free loads per lane,
                // reduced with redux.sync over the lanes that share a block.
                const int span = nvb << a.blk_shift;                   // staged positions that belong to blocks
                const int per_row = (span + 63) >> 6;                  // tasks per row
                const int rows_here = second ? 2 : 1;
                const unsigned seg = blk == 16 ? (lane < 16 ? 0x0000ffffu : 0xffff0000u) : 0xffffffffu;
                for (int task = warp; task < rows_here * per_row; task += NW) {
                    const int which = task >= per_row ? 1 : 0;
                    const int j0 = ((task - which * per_row) << 6) + lane, j1 = j0 + 32;
                    const float* __restrict__ row = st0 + which * a.cap;
                    const float c0 = j0 < span ? row[j0] : -INFINITY;
                    const float c1 = j1 < span ? row[j1] : -INFINITY;
                    float v0 = c0, v1 = c1;
                    if (blk == 64) v0 = v1 = fmaxf(c0, c1);            // one block spans both loads
                    int k0 = __reduce_max_sync(seg, float_key(v0));
                    int k1 = blk == 64 ? k0 : __reduce_max_sync(seg, float_key(v1));
                    v0 = __int_as_float(k0 ^ ((k0 >> 31) & 0x7fffffff));
                    v1 = __int_as_float(k1 ^ ((k1 >> 31) & 0x7fffffff));
                    if (!(v0 == v0)) v0 = -INFINITY;                   // a NaN never wins
                    if (!(v1 == v1)) v1 = -INFINITY;
                    int a0 = (c0 + 0.0f == v0) ? j0 : INT_MAX;
                    int a1 = (c1 + 0.0f == v1) ? j1 : INT_MAX;
                    if (blk == 64) a0 = a1 = min(a0, a1);
                    a0 = __reduce_min_sync(seg, a0);
                    if (blk != 64) a1 = __reduce_min_sync(seg, a1);
                    const bool leader = blk == 16 ? (lane & 15) == 0 : lane == 0;
                    if (leader) {
                        const int b0 = j0 >> a.blk_shift, b1 = j1 >> a.blk_shift;   // block indices relative to blk0
                        if (b0 < nvb) sBVw[which * 32 + b0] = make_float2(v0, __int_as_float(a0 == INT_MAX ? INT_MAX : start + a0));
                        if (blk != 64 && b1 < nvb)
                            sBVw[which * 32 + b1] = make_float2(v1, __int_as_float(a1 == INT_MAX ? INT_MAX : start + a1));
                    }
                }
            } else {
                int which = 0, i = warp;
                for (int task = warp; task < ntask; task += NW, i += NW) {
                    if (i >= nvb) { i -= nvb; which = 1; }
                    const float* __restrict__ row = st0 + which * a.cap + (i << a.blk_shift);
                    const int jbase = i << a.blk_shift;      // position of row[0] relative to `start`
                    const int lim = a.N - start - jbase;     // positions j >= lim are beyond the signal
                    float v = -INFINITY;
                    int at = INT_MAX;
                    for (int j = lane; j < blk; j += 32) {
                        const float c = row[j];
                        if (j < lim && c > v) { v = c; at = j; }
                    }
                    warp_argmax_redux(v, at);
                    if (lane == 0)
                        sBVw[which * 32 + i] = make_float2(v, __int_as_float(at == INT_MAX ? INT_MAX : start + jbase + at));
                }
            }
        }
        if constexpr (DEFER) {
            // no barrier here: this item's row maxima are re-derived during the next item (or after the loop)
            __syncwarp();
            if (lane == 0) mbar_arrive(barF);            // this warp reads the staging rows no more
            phaseF_armed = true;
            pv_row = q_ok ? (int)((size_t)b * a.nloc + 2 * q) : -1;
            pv_win = (blk0 << 6) | (nvb << 1) | (second ? 1 : 0);
            par ^= 1;
            if constexpr (DB) wpar ^= 1;
        } else {
            __syncthreads();   // block maxima staged; nobody reads the staging rows any more
#pragma unroll
            for (int wi = 0; wi < NROW; ++wi) {
                const int which = which0 + wi;
                if (!q_ok || which < 0 || which > 1 || (which == 1 && !second)) continue;
                row_phase((int)((size_t)b * a.nloc + 2 * q), which, blk0, nvb, sBV, old_v[wi], old_p[wi]);
            }
            // the next item's bulk loads are issued by tl == 0 after the barrier above; warps running
            // ahead only touch the FFT buffer until the next barrier.
        }
    }
    if constexpr (DEFER) {
        __syncthreads();       // the last item's block maxima are staged
        if (pv_row >= 0 && which0 >= 0 && which0 < 2 && (which0 == 0 || (pv_win & 1))) {
            const size_t rowi = (size_t)pv_row + which0;
            row_phase(pv_row, which0, pv_win >> 6, (pv_win >> 1) & 31, sBV + (par ^ 1) * 64, a.row_val[rowi], a.row_pos[rowi]);
        }
    }
    if (tl == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
}

// ---------------------------------------------------------------------------
// Selection on a dense map (B, K, N) the caller holds.  grid = (G, B): CTA g of
// signal b scans map rows g, g+G, ... with 128-bit loads and writes one
// candidate to part[g*B + b] (rank-major, the layout k_reduce_best takes).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void take_better3(float& v, int& k, int& t, float ov, int ok, int ot) {
    if (ov > v || (ov == v && (ok < k || (ok == k && ot < t)))) {
        v = ov;
        k = ok;
        t = ot;
    }
}

// LCN: the selection runs on fm - avg_pool2d(fm, 9x9, stride 1, zero padding 4) over the
// (atom, time) plane (modules/matchingpursuit.py:286-292); the candidate carries the
// normalised value, k_finish_best swaps in the raw one (:296).
template <bool LCN>
__device__ __forceinline__ float selection_value(const float* __restrict__ base, int K, int N, int k, int t) {
    const float c = base[(size_t)k * N + t];
    if constexpr (!LCN) {
        return c;
    } else {
        float s = 0.f;
#pragma unroll 1
        for (int dk = -4; dk <= 4; ++dk) {
            const int kk = k + dk;
            if (kk < 0 || kk >= K) continue;
            const float* __restrict__ row = base + (size_t)kk * N;
#pragma unroll
            for (int dt = -4; dt <= 4; ++dt) {
                const int tt = t + dt;
                if (tt >= 0 && tt < N) s = __fadd_rn(s, __ldg(row + tt));
            }
        }
        return __fsub_rn(c, __fdiv_rn(s, 81.f));
    }
}

template <bool LCN>
__global__ void __launch_bounds__(256)
k_select_dense(const float* __restrict__ fm, int K, int N, int atom_offset, Best* __restrict__ part) {
    const int b = blockIdx.y, B = gridDim.y;
    const float* __restrict__ base = fm + (size_t)b * K * N;
    float v = -INFINITY;
    int bk = INT_MAX, bt = INT_MAX;
    const bool vec = !LCN && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(fm) & 15) == 0);
    for (int k = blockIdx.x; k < K; k += gridDim.x) {
        const float* __restrict__ row = base + (size_t)k * N;
        if (vec) {
            const float4* __restrict__ row4 = reinterpret_cast<const float4*>(row);
            for (int i = threadIdx.x; i < N / 4; i += blockDim.x) {
                const float4 c = __ldg(row4 + i);
                // ascending t inside the thread: strict > keeps the first maximum
                if (c.x > v) { v = c.x; bk = k; bt = 4 * i; }
                if (c.y > v) { v = c.y; bk = k; bt = 4 * i + 1; }
                if (c.z > v) { v = c.z; bk = k; bt = 4 * i + 2; }
                if (c.w > v) { v = c.w; bk = k; bt = 4 * i + 3; }
            }
        } else {
            for (int t = threadIdx.x; t < N; t += blockDim.x) {
                const float c = selection_value<LCN>(base, K, N, k, t);
                if (c > v) { v = c; bk = k; bt = t; }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
        const int ot = __shfl_xor_sync(0xffffffffu, bt, off);
        take_better3(v, bk, bt, ov, ok, ot);
    }
    __shared__ float s_v[8];
    __shared__ int s_k[8], s_t[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_v[warp] = v; s_k[warp] = bk; s_t[warp] = bt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) take_better3(v, bk, bt, s_v[w], s_k[w], s_t[w]);
        Best r;
        if (bk == INT_MAX) {   // nothing compared greater than -inf (all NaN / -inf)
            r.value = -INFINITY;
            r.atom = INT_MAX;  // loses every tie-break in k_reduce_best
            r.position = INT_MAX;
        } else {
            r.value = v;
            r.atom = bk;       // local row; k_finish_best adds atom_offset
            r.position = bt;
        }
        r.pad = 0;
        part[(size_t)blockIdx.x * B + b] = r;
    }
}

// After k_reduce_best on k_select_dense candidates: a signal whose map held no
// value above -inf gets row 0, position 0; the value becomes the RAW map value at
// the winner (identical to the candidate's unless LCN); atom_offset is applied.
__global__ void k_finish_best(Best* __restrict__ best, int batch, const float* __restrict__ fm, int K, int N,
                              int atom_offset) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    Best r = best[b];
    if (r.atom == INT_MAX) {
        r.atom = 0;
        r.position = 0;
    }
    r.value = fm[((size_t)b * K + r.atom) * N + r.position];
    r.atom += atom_offset;
    best[b] = r;
}

// residual[b, p:p+A] -= value * dict[atom]  (fl(r - fl(v*d)), truncated at the right edge).
// grid = batch, 256 threads.
__global__ void __launch_bounds__(256)
k_subtract(float* __restrict__ residual, int N, const float* __restrict__ dict, int K, int A,
           const Best* __restrict__ winner) {
    const int b = blockIdx.x;
    const Best w = winner[b];
    if (w.atom < 0 || w.atom >= K || w.position < 0 || w.position >= N) return;
    float* __restrict__ r = residual + (size_t)b * N + w.position;
    const float* __restrict__ d = dict + (size_t)w.atom * A;
    const int keep = min(A, N - w.position);
    for (int i = threadIdx.x; i < keep; i += blockDim.x) r[i] = __fsub_rn(r[i], __fmul_rn(w.value, d[i]));
}

// ---------------------------------------------------------------------------
// Decode helpers (modules/matchingpursuit.py:20-58, :305).
// ---------------------------------------------------------------------------
// Deterministic: a thread owns output samples and walks the events in list
// order, so overlapping atoms are summed in the reference's order.
//   out[row_index[e], pos[e] + i] += scale[e] * src[src_index[e], i],  i < A, truncated at N
// src_index == null -> row e of src; scale == null -> 1 (adds src rows as they are).
// row_offsets (n_rows + 1 entries) says the events are sorted by row and which
// range belongs to each row; without it every CTA filters the whole list.
// grid = (ceil(N / 1024), n_rows), 256 threads, 4 samples per thread.
__global__ void __launch_bounds__(256)
k_scatter(float* __restrict__ out, int N, const float* __restrict__ src, int n_src, int A,
          const int* __restrict__ src_index, const int* __restrict__ row_index, const int* __restrict__ pos,
          const float* __restrict__ scale, const int* __restrict__ row_offsets, int n_events) {
    const int b = blockIdx.y;
    const int tile0 = blockIdx.x * 1024;
    float acc[4];
    int t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        t[u] = tile0 + threadIdx.x + 256 * u;
        acc[u] = t[u] < N ? out[(size_t)b * N + t[u]] : 0.f;
    }
    const int e0 = row_offsets ? row_offsets[b] : 0;
    const int e1 = row_offsets ? row_offsets[b + 1] : n_events;
    for (int e = e0; e < e1; ++e) {
        if (row_index[e] != b) continue;
        const int p = pos[e];
        const int k = src_index ? src_index[e] : e;
        if (k < 0 || k >= n_src || p + A <= tile0 || p >= tile0 + 1024 || p < 0) continue;
        const float v = scale ? scale[e] : 1.f;
        const float* d = src + (size_t)k * A;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = t[u] - p;
            if (i >= 0 && i < A) acc[u] = __fadd_rn(acc[u], __fmul_rn(v, d[i]));
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (t[u] < N) out[(size_t)b * N + t[u]] = acc[u];
}

// Long atoms (longer than one window transform can hold) are correlated as P consecutive parts of L samples:
//     fm[b, k, t] = sum_p sub[b, k*P + p, t + p*L],   terms with t + p*L >= N are zero (the signal ends there)
// where sub is the dense map of the (K*P, L) dictionary of parts.  grid = (ceil(N/256), K, B).
__global__ void __launch_bounds__(256)
k_fold_parts(const float* __restrict__ sub, int K, int P, int L, int N, float* __restrict__ out) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= N) return;
    const int k = blockIdx.y, b = blockIdx.z;
    const float* __restrict__ base = sub + ((size_t)b * K + k) * P * (size_t)N;
    float acc = 0.f;
    for (int p = 0; p < P; ++p) {
        const long long tt = (long long)t + (long long)p * L;
        if (tt >= N) break;
        acc = __fadd_rn(acc, base[(size_t)p * N + tt]);
    }
    out[((size_t)b * K + k) * N + t] = acc;
}

// ---------------------------------------------------------------------------
// Incremental local-contrast-norm selection (modules/matchingpursuit.py:286-296).  The selection runs on
//     norm[k, t] = fm[k, t] - avg_pool2d(fm, 9x9, stride 1, zero padding 4)[k, t]
// over the (atom, time) plane.  A step changes the raw map in the +-A window of its winner only (all rows), so the
// normalised map changes in that window widened by 4 columns: this kernel recomputes the normalised values of the
// whole blocks that cover it -- every row, from the RESIDENT raw map -- and refreshes a second block-max / row-max
// hierarchy (nbm_*, nrow_*) that the selection reads.  `full` (first pass): every block of every row.
// The 81-term sums run in the reference's order (rows outer, columns inner, one fp32 running sum, then / 81).
// grid = (ceil(nloc / 8), batch), 256 threads: warp w owns row k0 + w; the CTA walks the window block by block with
// a (16 rows x (blk + 8) columns) shared tile.  The row maximum is re-derived from the row's block table.
// ---------------------------------------------------------------------------
struct LcnArgs {
    const float* map;        // (B, nloc, NS) raw correlation map
    const GramUpdate* upd;   // (B) winner of the step (its +-A window is what changed); unused when full
    float* nbm_val;          // (B, nloc, NB) block maxima of the normalised map
    int* nbm_pos;
    float* nrow_val;         // (B, nloc)
    int* nrow_pos;
    int nloc, N, NS, NB, blk_shift, A, full;
};

__global__ void __launch_bounds__(256)
k_lcn_refresh(const LcnArgs a) {
    extern __shared__ float s_tile[];                    // [16][blk + 8]
    const int b = blockIdx.y, k0 = blockIdx.x * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = 1 << a.blk_shift, pitch = blk + 8;
    pdl_prologue();
    int blk_lo = 0, blk_hi = a.NB;                       // blocks [blk_lo, blk_hi) are refreshed
    if (!a.full) {
        const int p = a.upd[b].position;
        const int c_lo = max(0, p - a.A + 1 - 4), c_hi = min(a.N - 1, p + a.A - 1 + 4);
        blk_lo = c_lo >> a.blk_shift;
        blk_hi = (c_hi >> a.blk_shift) + 1;
    }
    const float* __restrict__ base = a.map + (size_t)b * a.nloc * a.NS;
    const int k = k0 + warp;
    for (int bi = blk_lo; bi < blk_hi; ++bi) {
        const int t0 = bi << a.blk_shift;
        __syncthreads();                                 // the previous block's tile has been consumed
        for (int i = threadIdx.x; i < 16 * pitch; i += 256) {
            const int rr = i / pitch, cc = i - rr * pitch;
            const int kk = k0 - 4 + rr, tt = t0 - 4 + cc;
            s_tile[i] = (kk >= 0 && kk < a.nloc && tt >= 0 && tt < a.N) ? base[(size_t)kk * a.NS + tt] : 0.f;
        }
        __syncthreads();
        if (k < a.nloc) {
            float v = -INFINITY;
            int at = INT_MAX;
            for (int j = lane; j < blk; j += 32) {       // ascending positions inside the lane
                if (t0 + j >= a.N) break;
                float sum = 0.f;
#pragma unroll 1
                for (int dk = 0; dk < 9; ++dk) {
                    const float* __restrict__ row = s_tile + (warp + dk) * pitch + j;
#pragma unroll
                    for (int dt = 0; dt < 9; ++dt) sum = __fadd_rn(sum, row[dt]);
                }
                const float c = __fsub_rn(s_tile[(warp + 4) * pitch + j + 4], __fdiv_rn(sum, 81.f));
                if (c > v) { v = c; at = t0 + j; }
            }
            warp_argmax(v, at);
            if (lane == 0) {
                const size_t o = ((size_t)b * a.nloc + k) * a.NB + bi;
                a.nbm_val[o] = v;
                a.nbm_pos[o] = at;
            }
        }
    }
    if (k < a.nloc) {                                    // row maximum over the row's (now current) block table
        __syncwarp();
        const size_t rowi = (size_t)b * a.nloc + k;
        float v = -INFINITY;
        int at = INT_MAX;
        rescan_row(a.nbm_val + rowi * a.NB, a.nbm_pos + rowi * a.NB, a.NB, 0, 0, lane, v, at);
        warp_argmax(v, at);
        if (lane == 0) {
            a.nrow_val[rowi] = v;
            a.nrow_pos[rowi] = (at == INT_MAX) ? 0 : at;
        }
    }
}

// ---------------------------------------------------------------------------
// Dictionary-learning atom update (modules/matchingpursuit.py:391-417) as ONE launch.  The events of a coding
// pass are grouped by atom in first-seen order; the groups depend on each other through the running signal, so a
// single CTA walks them in order and, per group:
//   1. adds the group's scaled atoms back to the running signal            (:395-396)
//   2. sums the segments under the group's events into the new atom        (:398-400; zero beyond the signal)
//   3. unit-norms it: x / (||x|| + 1e-8), and stores it as row `atom`      (:402-406)
//   4. subtracts new_atom * ||scaled atom|| at every event                  (:408-415)
// Events of one group may overlap in time, so steps 1 and 4 go event by event with a CTA barrier in between
// (the reference's list order); the sums of step 2 run in event order per sample.
// Dynamic shared memory: A floats (the new atom).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_dictionary_update(float* __restrict__ running, int N, float* __restrict__ dict, int A,
                    const int* __restrict__ group_offsets, const int* __restrict__ group_atom, int n_groups,
                    const int* __restrict__ ev_batch, const int* __restrict__ ev_pos, const float* __restrict__ ev_rows) {
    extern __shared__ float s_atom[];
    __shared__ double s_part[32];
    const int tid = threadIdx.x, nthr = blockDim.x;
    auto block_sum = [&](double v) -> double {          // every thread returns the CTA-wide sum
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        __syncthreads();
        if ((tid & 31) == 0) s_part[tid >> 5] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < (nthr >> 5); ++w) t += s_part[w];
        return t;
    };
    for (int g = 0; g < n_groups; ++g) {
        const int e0 = group_offsets[g], e1 = group_offsets[g + 1];
        for (int e = e0; e < e1; ++e) {                  // 1. add the instances back
            float* __restrict__ r = running + (size_t)ev_batch[e] * N + ev_pos[e];
            const float* __restrict__ row = ev_rows + (size_t)e * A;
            const int keep = min(A, N - ev_pos[e]);
            for (int i = tid; i < keep; i += nthr) r[i] = __fadd_rn(r[i], row[i]);
            __syncthreads();
        }
        double sq = 0.0;                                 // 2. sum of the segments
        for (int i = tid; i < A; i += nthr) {
            float acc = 0.f;
            for (int e = e0; e < e1; ++e) {
                const int p = ev_pos[e];
                if (p + i < N) acc = __fadd_rn(acc, running[(size_t)ev_batch[e] * N + p + i]);
            }
            s_atom[i] = acc;
            sq += (double)acc * (double)acc;
        }
        const float denom = __fadd_rn((float)sqrt(block_sum(sq)), 1e-8f);       // 3. unit norm
        for (int i = tid; i < A; i += nthr) {
            const float v = __fdiv_rn(s_atom[i], denom);
            s_atom[i] = v;
            dict[(size_t)group_atom[g] * A + i] = v;
        }
        __syncthreads();
        for (int e = e0; e < e1; ++e) {                  // 4. subtract the re-scaled new atom
            const float* __restrict__ row = ev_rows + (size_t)e * A;
            double a2 = 0.0;
            for (int i = tid; i < A; i += nthr) a2 += (double)row[i] * (double)row[i];
            const float amp = (float)sqrt(block_sum(a2));
            float* __restrict__ r = running + (size_t)ev_batch[e] * N + ev_pos[e];
            const int keep = min(A, N - ev_pos[e]);
            for (int i = tid; i < keep; i += nthr) r[i] = __fsub_rn(r[i], __fmul_rn(s_atom[i], amp));
            __syncthreads();
        }
    }
}

__global__ void k_gather_atoms(float* __restrict__ scaled, const float* __restrict__ dict, int K, int A,
                               const int* __restrict__ atom, const float* __restrict__ val, int n_events) {
    const int e = blockIdx.x;
    if (e >= n_events) return;
    const int k = atom[e];
    const float v = val[e];
    const float* d = dict + (size_t)k * A;
    float* o = scaled + (size_t)e * A;
    for (int i = threadIdx.x; i < A; i += blockDim.x) o[i] = (k >= 0 && k < K) ? __fmul_rn(d[i], v) : 0.f;
}

}  // namespace mpb
