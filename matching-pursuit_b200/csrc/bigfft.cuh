// bigfft.cuh -- batched power-of-two FFTs of real rows up to 2^18 points, built from the
// block FFT of fft_core.cuh.  Used by the helpers around the pursuit (modules/fft.py::fft_convolve,
// modules/decompose.py band split / merge, mp.py's forward), not by the pursuit loop itself.
//
// L = R * L2,  R in {1,2,4,8,16,32},  L2 in {256..8192} (one CTA-resident transform).
// Index split j = j1*L2 + c, f = m1 + R*mu:
//     X[m1 + R*mu] = sum_c w_L2^{c mu} * ( w_L^{c m1} * sum_j1 x[j1*L2 + c] w_R^{j1 m1} )
//   forward  = k_bf_cols_fwd (radix-R across the R sub-blocks + twiddle, elementwise, coalesced)
//              then k_bf_rows<-1> (R independent L2-point transforms per row)
//   inverse  = k_bf_rows<+1> (with the spectrum PRODUCT of up to 4 operands fused into its loads,
//              and the conjugate twiddle fused into its stores) then k_bf_cols_inv.
// Spectra live in the PERMUTED layout spec[row][m1][mu] (bin f = m1 + R*mu); pointwise products do
// not care, and the inverse consumes exactly that layout, so no transpose is ever made.
#pragma once
#include <cuda_runtime.h>
#include "fft_core.cuh"
#include "types.h"

namespace mpb {

constexpr int BF_MAX_OPS = 4;

struct BfGeom {
    int L, R, L2;
};

// ---------------------------------------------------------------------------
// forward columns: real rows (rows, n) zero-padded to L  ->  y (rows, R, L2) complex
// one thread per (row, c); grid = (ceil(L2/256), rows)
// ---------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256)
k_bf_cols_fwd(const float* __restrict__ x, int n, long long row_stride, int L2, const C32* __restrict__ twL,
              C32* __restrict__ y) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int row = blockIdx.y;
    if (c >= L2) return;
    const float* __restrict__ xr = x + (long long)row * row_stride;
    C32 v[R];
#pragma unroll
    for (int j1 = 0; j1 < R; ++j1) {
        const int t = j1 * L2 + c;
        v[j1] = {t < n ? xr[t] : 0.f, 0.f};
    }
    Dft<R, -1, float>::run(v);
    C32* __restrict__ yr = y + (size_t)row * R * L2;
#pragma unroll
    for (int m1 = 0; m1 < R; ++m1) {
        C32 o = v[m1];
        if (m1 > 0) o = cmul(o, cconj(twL[c * m1]));   // c*m1 < L
        yr[(size_t)m1 * L2 + c] = o;
    }
}

// ---------------------------------------------------------------------------
// rows: L2-point transforms of contiguous complex rows.
//   DIR = -1: in = src[t]            (t = transform index = row*R + m1), out = dst[t][natural order]
//   DIR = +1: in = prod_i conj?(op_i[rowmap_i[row]][m1][:]),  out = dst[t][c] * twL[c*m1]
// grid = ceil(n_transforms / NT) CTAs of max(T,128) threads.
// ---------------------------------------------------------------------------
struct BfRowsArgs {
    const C32* ops[BF_MAX_OPS];
    const int* rowmap[BF_MAX_OPS];   // rows_out entries each, or null = identity
    int conj_mask;                   // bit i: conjugate operand i
    int n_ops;
    int n_transforms;                // rows_out * R
    int R;
    const C32* twL;
    const C32* tw1;
    const C32* tw2;
    C32* dst;
};

template <int L2, int DIR>
__global__ void __launch_bounds__((BlockFft<L2, float>::T < 128 ? 128 : BlockFft<L2, float>::T))
k_bf_rows(const BfRowsArgs a) {
    using F = BlockFft<L2, float>;
    constexpr int TPB = F::T < 128 ? 128 : F::T;
    constexpr int NT = TPB / F::T;
    extern __shared__ __align__(16) unsigned char smraw[];
    C32* stw2 = reinterpret_cast<C32*>(smraw);
    const int sb = threadIdx.x / F::T, tl = threadIdx.x % F::T;
    C32* sm = stw2 + 256 + (size_t)sb * F::SMEM_CPX;
    for (int i = threadIdx.x; i < 256; i += TPB) stw2[i] = a.tw2[i];
    int t = blockIdx.x * NT + sb;
    const bool ok = t < a.n_transforms;
    if (!ok) t = a.n_transforms - 1;
    const int row = t / a.R, m1 = t % a.R;
    C32 r[F::E];
    if constexpr (DIR < 0) {
        const C32* __restrict__ src = a.ops[0] + (size_t)t * L2;
#pragma unroll
        for (int e = 0; e < F::E; ++e) r[e] = src[F::in_index(tl, e)];
    } else {
#pragma unroll
        for (int e = 0; e < F::E; ++e) r[e] = {1.f, 0.f};
        for (int i = 0; i < a.n_ops; ++i) {
            const int srow = a.rowmap[i] ? a.rowmap[i][row] : row;
            const C32* __restrict__ src = a.ops[i] + ((size_t)srow * a.R + m1) * L2;
            const bool cj = (a.conj_mask >> i) & 1;
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                C32 v = src[F::in_index(tl, e)];
                if (cj) v.y = -v.y;
                r[e] = (i == 0) ? v : cmul(r[e], v);
            }
        }
    }
    __syncthreads();
    F::template pass1<DIR>(r, tl, sm, a.tw1);
    __syncthreads();
    F::template pass2<DIR>(r, tl, sm, stw2);
    __syncthreads();
    F::template pass3<DIR>(r, tl, sm);
    if (ok) {
        C32* __restrict__ dst = a.dst + (size_t)t * L2;
#pragma unroll
        for (int e = 0; e < F::E; ++e) {
            const int c = F::out_index(tl, e);
            C32 o = r[e];
            if constexpr (DIR > 0) {
                if (m1 > 0) o = cmul(o, a.twL[c * m1]);
            }
            dst[c] = o;
        }
    }
}

// ---------------------------------------------------------------------------
// inverse columns: s (rows, R, L2) complex -> real x[j1*L2 + c] = Re sum_m1 s[m1][c] w_R^{-j1 m1},
// scaled, written for t < n_keep into out (rows, out_stride).  grid = (ceil(L2/256), rows)
// ---------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256)
k_bf_cols_inv(const C32* __restrict__ s, int L2, float scale, int n_keep, long long out_stride,
              float* __restrict__ out) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int row = blockIdx.y;
    if (c >= L2) return;
    const C32* __restrict__ sr = s + (size_t)row * R * L2;
    C32 v[R];
#pragma unroll
    for (int m1 = 0; m1 < R; ++m1) v[m1] = sr[(size_t)m1 * L2 + c];
    Dft<R, 1, float>::run(v);
    float* __restrict__ o = out + (long long)row * out_stride;
#pragma unroll
    for (int j1 = 0; j1 < R; ++j1) {
        const int t = j1 * L2 + c;
        if (t < n_keep) o[t] = v[j1].x * scale;
    }
}

// out[row][t] = sum_{a >= 0, t + a*W < L} full[row][t + a*W],  t < n   (circular wrap of a longer
// linear result onto period W, then crop: N-ary products of modules/fft.py:23-35)
__global__ void k_bf_wrap(const float* __restrict__ full, int L, int W, int n, float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (t >= n) return;
    float acc = 0.f;
    for (int u = t; u < L; u += W) acc += full[(size_t)row * L + u];
    out[(size_t)row * n + t] = acc;
}

// ---------------------------------------------------------------------------
// band transfer between two permuted spectra (modules/decompose.py:5-33, 36-73):
//   T[f'] = X[f'] for lo <= f' < hi (f' <= n_out/2), Hermitian-completed to n_out bins, with the
//   imaginary parts of bin 0 and of bin n_out/2 dropped (irfft semantics), zero elsewhere.
// src layout (rows, R_in, L2_in), dst layout (rows, R_out, L2_out).  one thread per (row, f').
// ---------------------------------------------------------------------------
__global__ void k_bf_band(const C32* __restrict__ src, int n_in, int R_in, int L2_in, C32* __restrict__ dst,
                          int n_out, int R_out, int L2_out, int lo, int hi) {
    const int fp = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (fp >= n_out) return;
    const int half = n_out / 2;
    const bool mirrored = fp > half;
    const int f = mirrored ? n_out - fp : fp;
    C32 v = {0.f, 0.f};
    if (f >= lo && f < hi && f <= n_in / 2) {
        v = src[((size_t)row * R_in + (f % R_in)) * L2_in + f / R_in];
        if (f == 0 || f == half) v.y = 0.f;
        if (mirrored) v.y = -v.y;
    }
    dst[((size_t)row * R_out + (fp % R_out)) * L2_out + fp / R_out] = v;
}

}  // namespace mpb
