// fftconv.cu -- FFT helpers around the pursuit, behind the C ABI of include/mpb200.h:
//   mpb200_fft_convolve   modules/fft.py:23-35 (N-ary zero-padded FFT convolution; also
//                         modules/transfer.py:548-569) -- spectra product fused into the inverse
//   mpb200_spectral_band  modules/decompose.py:5-33, 36-73 (ortho rfft -> keep a band of bins ->
//                         ortho irfft at another length)
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mpb200.h"
#include "bigfft.cuh"
#include "plan.h"

namespace mpb {

extern std::atomic<unsigned long long> g_launches;

#define BF_CUDA(expr)                                                                           \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(MPB200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)
#define BF_LAUNCH(name)                                                                         \
    do {                                                                                        \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                     \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return fail(MPB200_ECUDA, std::string("launch ") + name + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---- twiddle tables, cached per (device, size), never freed -----------------
struct BfTables {
    C32* twL = nullptr;   // exp(+2 pi i t / L), t < L
    C32* tw1 = nullptr;   // BlockFft<L2> tables
    C32* tw2 = nullptr;
};
static std::mutex g_tab_mutex;
static std::map<std::pair<int, long long>, C32*> g_tab;   // (device, key) -> device pointer

static int upload(int dev, long long key, const std::vector<C32>& host, C32** out) {
    auto it = g_tab.find({dev, key});
    if (it != g_tab.end()) { *out = it->second; return MPB200_OK; }
    C32* p = nullptr;
    BF_CUDA(cudaMalloc((void**)&p, host.size() * sizeof(C32)));
    BF_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(C32), cudaMemcpyHostToDevice));
    g_tab[{dev, key}] = p;
    *out = p;
    return MPB200_OK;
}

static int get_tables(const BfGeom& g, BfTables* t) {
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    int dev = 0;
    BF_CUDA(cudaGetDevice(&dev));
    const long double two_pi = 6.283185307179586476925286766559005768L;
    int rc;
    if (g_tab.find({dev, (long long)g.L}) == g_tab.end()) {
        std::vector<C32> h((size_t)g.L);
        for (int i = 0; i < g.L; ++i) {
            long double a = two_pi * (long double)i / (long double)g.L;
            h[i] = {(float)cosl(a), (float)sinl(a)};
        }
        if ((rc = upload(dev, g.L, h, &t->twL))) return rc;
    } else {
        t->twL = g_tab[{dev, (long long)g.L}];
    }
    const long long k1 = (1LL << 40) + g.L2, k2 = (2LL << 40);
    if (g_tab.find({dev, k1}) == g_tab.end()) {
        const int R1 = g.L2 / 256;
        std::vector<C32> h((size_t)g.L2);
        for (int m1 = 0; m1 < R1; ++m1)
            for (int c = 0; c < 256; ++c) {
                long double a = two_pi * (long double)(((long long)c * m1) % g.L2) / (long double)g.L2;
                h[(size_t)m1 * 256 + c] = {(float)cosl(a), (float)sinl(a)};
            }
        if ((rc = upload(dev, k1, h, &t->tw1))) return rc;
    } else {
        t->tw1 = g_tab[{dev, k1}];
    }
    if (g_tab.find({dev, k2}) == g_tab.end()) {
        std::vector<C32> h(256);
        for (int m2 = 0; m2 < 16; ++m2)
            for (int j3 = 0; j3 < 16; ++j3) {
                long double a = two_pi * (long double)((j3 * m2) % 256) / 256.0L;
                h[m2 * 16 + j3] = {(float)cosl(a), (float)sinl(a)};
            }
        if ((rc = upload(dev, k2, h, &t->tw2))) return rc;
    } else {
        t->tw2 = g_tab[{dev, k2}];
    }
    return MPB200_OK;
}

static bool make_geom(long long L, BfGeom* g) {
    if (L < 256 || L > (1 << 18) || (L & (L - 1))) return false;
    int L2 = L < 4096 ? (int)L : 4096;
    int R = (int)(L / L2);
    if (R > 32) { L2 = 8192; R = (int)(L / L2); }
    g->L = (int)L; g->R = R; g->L2 = L2;
    return true;
}

static long long pow2_at_least(long long v) {
    long long p = 256;
    while (p < v) p <<= 1;
    return p;
}

#define BF_DISPATCH_R(r, ...)                                        \
    switch (r) {                                                     \
        case 1: { constexpr int RR = 1; __VA_ARGS__; } break;        \
        case 2: { constexpr int RR = 2; __VA_ARGS__; } break;        \
        case 4: { constexpr int RR = 4; __VA_ARGS__; } break;        \
        case 8: { constexpr int RR = 8; __VA_ARGS__; } break;        \
        case 16: { constexpr int RR = 16; __VA_ARGS__; } break;      \
        case 32: { constexpr int RR = 32; __VA_ARGS__; } break;      \
        default: return fail(MPB200_EINVAL, "unsupported radix");    \
    }
#define BF_DISPATCH_L2(m, ...)                                       \
    switch (m) {                                                     \
        case 256: { constexpr int MM = 256; __VA_ARGS__; } break;    \
        case 512: { constexpr int MM = 512; __VA_ARGS__; } break;    \
        case 1024: { constexpr int MM = 1024; __VA_ARGS__; } break;  \
        case 2048: { constexpr int MM = 2048; __VA_ARGS__; } break;  \
        case 4096: { constexpr int MM = 4096; __VA_ARGS__; } break;  \
        case 8192: { constexpr int MM = 8192; __VA_ARGS__; } break;  \
        default: return fail(MPB200_EINVAL, "unsupported row FFT size"); \
    }

static int launch_cols_fwd(const BfGeom& g, const BfTables& t, const float* x, int n, long long stride, int rows,
                           C32* y, cudaStream_t st) {
    dim3 grid((g.L2 + 255) / 256, rows);
    BF_DISPATCH_R(g.R, (k_bf_cols_fwd<RR><<<grid, 256, 0, st>>>(x, n, stride, g.L2, t.twL, y)));
    BF_LAUNCH("k_bf_cols_fwd");
    return MPB200_OK;
}

template <int DIR>
static int launch_rows(const BfGeom& g, const BfTables& t, BfRowsArgs a, cudaStream_t st) {
    a.R = g.R;
    a.twL = t.twL;
    a.tw1 = t.tw1;
    a.tw2 = t.tw2;
    BF_DISPATCH_L2(g.L2, {
        using F = BlockFft<MM, float>;
        constexpr int TPB = F::T < 128 ? 128 : F::T;
        constexpr int NT = TPB / F::T;
        const size_t smem = (size_t)(256 + NT * F::SMEM_CPX) * sizeof(C32);
        static bool once[64] = {false};           // the attribute is per device
        int dev_ = 0;
        BF_CUDA(cudaGetDevice(&dev_));
        if (dev_ < 0 || dev_ >= 64 || !once[dev_]) {
            BF_CUDA(cudaFuncSetAttribute(k_bf_rows<MM, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev_ >= 0 && dev_ < 64) once[dev_] = true;
        }
        k_bf_rows<MM, DIR><<<(a.n_transforms + NT - 1) / NT, TPB, smem, st>>>(a);
    });
    BF_LAUNCH("k_bf_rows");
    return MPB200_OK;
}

static int launch_cols_inv(const BfGeom& g, const C32* s, int rows, float scale, int n_keep, long long out_stride,
                           float* out, cudaStream_t st) {
    dim3 grid((g.L2 + 255) / 256, rows);
    BF_DISPATCH_R(g.R, (k_bf_cols_inv<RR><<<grid, 256, 0, st>>>(s, g.L2, scale, n_keep, out_stride, out)));
    BF_LAUNCH("k_bf_cols_inv");
    return MPB200_OK;
}

// forward transform of real rows into a fresh stream-ordered buffer (permuted layout)
static int forward_real(const BfGeom& g, const BfTables& t, const float* x, int n, int rows, C32** spec,
                        cudaStream_t st) {
    BF_CUDA(cudaMallocAsync((void**)spec, (size_t)rows * g.L * sizeof(C32), st));
    int rc = launch_cols_fwd(g, t, x, n, n, rows, *spec, st);
    if (rc) return rc;
    BfRowsArgs a = {};
    a.ops[0] = *spec;
    a.n_ops = 1;
    a.n_transforms = rows * g.R;
    a.dst = *spec;   // in place: a CTA reads its whole row before it writes
    return launch_rows<-1>(g, t, a, st);
}

}  // namespace mpb

using namespace mpb;

extern "C" int mpb200_fft_convolve(const float* const* operands, const int32_t* const* row_maps,
                                   const int32_t* operand_rows, int n_ops, int rows_out, int n, int conj_mask,
                                   float scale, float* out, void* stream) {
    if (!operands || !operand_rows || !out || n_ops < 1 || n_ops > BF_MAX_OPS || rows_out < 1 || n < 1)
        return fail(MPB200_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long lin = (long long)n_ops * n - n_ops + 1;      // length of the full linear product
    const long long W = 2LL * n;                                 // the reference's circular length
    BfGeom g;
    if (!make_geom(pow2_at_least(lin > W ? lin : W), &g))
        return fail(MPB200_EINVAL, "fft_convolve: length too large (n_ops * n must stay below 2^18)");
    BfTables t;
    int rc = get_tables(g, &t);
    if (rc) return rc;
    C32* spec[BF_MAX_OPS] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < n_ops; ++i) {
        if (!operands[i] || operand_rows[i] < 1) return fail(MPB200_EINVAL, "null operand");
        rc = forward_real(g, t, operands[i], n, operand_rows[i], &spec[i], st);
        if (rc) return rc;
    }
    C32* s = nullptr;
    BF_CUDA(cudaMallocAsync((void**)&s, (size_t)rows_out * g.L * sizeof(C32), st));
    BfRowsArgs a = {};
    for (int i = 0; i < n_ops; ++i) {
        a.ops[i] = spec[i];
        a.rowmap[i] = row_maps ? row_maps[i] : nullptr;
    }
    a.conj_mask = conj_mask;
    a.n_ops = n_ops;
    a.n_transforms = rows_out * g.R;
    a.dst = s;
    rc = launch_rows<1>(g, t, a, st);
    if (rc) return rc;
    const float total_scale = scale / (float)g.L;
    if (lin <= W) {
        rc = launch_cols_inv(g, s, rows_out, total_scale, n, n, out, st);
        if (rc) return rc;
    } else {
        float* full = nullptr;
        BF_CUDA(cudaMallocAsync((void**)&full, (size_t)rows_out * g.L * sizeof(float), st));
        rc = launch_cols_inv(g, s, rows_out, total_scale, g.L, g.L, full, st);
        if (rc) return rc;
        k_bf_wrap<<<dim3((n + 255) / 256, rows_out), 256, 0, st>>>(full, g.L, (int)W, n, out);
        BF_LAUNCH("k_bf_wrap");
        BF_CUDA(cudaFreeAsync(full, st));
    }
    BF_CUDA(cudaFreeAsync(s, st));
    for (int i = 0; i < n_ops; ++i) BF_CUDA(cudaFreeAsync(spec[i], st));
    return MPB200_OK;
}

extern "C" int mpb200_spectral_band(const float* x, int rows, int n_in, float* out, int n_out, int bin_lo,
                                    int bin_hi, void* stream) {
    if (!x || !out || rows < 1) return fail(MPB200_EINVAL, "bad argument");
    BfGeom gi, go;
    if (!make_geom(n_in, &gi) || !make_geom(n_out, &go))
        return fail(MPB200_EINVAL, "spectral_band: lengths must be powers of two in [256, 2^18]");
    if (bin_lo < 0 || bin_hi < bin_lo) return fail(MPB200_EINVAL, "bad bin range");
    cudaStream_t st = (cudaStream_t)stream;
    BfTables ti, to;
    int rc = get_tables(gi, &ti);
    if (!rc) rc = get_tables(go, &to);
    if (rc) return rc;
    C32* spec = nullptr;
    rc = forward_real(gi, ti, x, n_in, rows, &spec, st);
    if (rc) return rc;
    C32* band = nullptr;
    BF_CUDA(cudaMallocAsync((void**)&band, (size_t)rows * go.L * sizeof(C32), st));
    k_bf_band<<<dim3((n_out + 255) / 256, rows), 256, 0, st>>>(spec, n_in, gi.R, gi.L2, band, n_out, go.R, go.L2,
                                                              bin_lo, bin_hi);
    BF_LAUNCH("k_bf_band");
    BfRowsArgs a = {};
    a.ops[0] = band;
    a.n_ops = 1;
    a.n_transforms = rows * go.R;
    a.dst = band;
    rc = launch_rows<1>(go, to, a, st);
    if (rc) return rc;
    const float scale = (float)(1.0 / std::sqrt((double)n_in * (double)n_out));
    rc = launch_cols_inv(go, band, rows, scale, n_out, n_out, out, st);
    if (rc) return rc;
    BF_CUDA(cudaFreeAsync(band, st));
    BF_CUDA(cudaFreeAsync(spec, st));
    return MPB200_OK;
}

// ---------------------------------------------------------------------------
// mpb200_band_limit -- the spectral mask of modules/conv.py:24-29 (fft_convolve(approx=slice)):
//     y = irfft( mask_bins( rfft( pad(x, L) ) ) ),   L even, arbitrary (N + atom_size in the reference)
// as two direct DFT kernels over the kept bins only (an arithmetic progression bin0, bin0+step, ...):
// the slice is usually a small band, and L = N + A is not a power of two.  The product with the
// atom spectra of the reference is then the engine's ordinary correlation of y (t + i < L never wraps).
// ---------------------------------------------------------------------------
namespace mpb {

// spec[row, jb] = sum_s x[row, s] * exp(-2 pi i * bin_jb * s / L).   grid = (n_bins, rows), 256 threads.
__global__ void __launch_bounds__(256)
k_band_forward(const float* __restrict__ x, int n, int L, int bin0, int bstep, const C32* __restrict__ twL,
               C32* __restrict__ spec) {
    const int jb = blockIdx.x, row = blockIdx.y, n_bins = gridDim.x;
    const unsigned long long j = (unsigned long long)(bin0 + jb * bstep);
    const float* __restrict__ xr = x + (size_t)row * n;
    unsigned idx = (unsigned)((j * threadIdx.x) % (unsigned long long)L);
    const unsigned inc = (unsigned)((j * 256ull) % (unsigned long long)L);
    double re = 0.0, im = 0.0;
    for (int s = threadIdx.x; s < n; s += 256) {
        const C32 w = twL[idx];
        const float v = xr[s];
        re += (double)(v * w.x);
        im -= (double)(v * w.y);
        idx += inc;
        if (idx >= (unsigned)L) idx -= (unsigned)L;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, off);
        im += __shfl_xor_sync(0xffffffffu, im, off);
    }
    __shared__ double s_re[8], s_im[8];
    if ((threadIdx.x & 31) == 0) { s_re[threadIdx.x >> 5] = re; s_im[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { re += s_re[w]; im += s_im[w]; }
        spec[(size_t)row * n_bins + jb] = {(float)re, (float)im};
    }
}

// y[row, t] = (1/L) sum_jb c_jb * Re( spec[row, jb] * exp(+2 pi i * bin_jb * t / L) ),  c = 1 for DC / Nyquist
// (whose imaginary parts a C2R transform ignores), 2 otherwise.   grid = (ceil(L/256), rows), 256 threads.
__global__ void __launch_bounds__(256)
k_band_inverse(const C32* __restrict__ spec, int n_bins, int L, int bin0, int bstep, const C32* __restrict__ twL,
               float* __restrict__ y) {
    const int row = blockIdx.y;
    const int t = blockIdx.x * 256 + threadIdx.x;
    __shared__ C32 s_spec[256];
    const unsigned long long tt = (unsigned long long)(t < L ? t : 0);
    unsigned idx = (unsigned)(((unsigned long long)bin0 * tt) % (unsigned long long)L);
    const unsigned inc = (unsigned)(((unsigned long long)bstep * tt) % (unsigned long long)L);
    float acc = 0.f, comp = 0.f;   // Kahan: up to L/2 terms of mixed sign
    for (int base = 0; base < n_bins; base += 256) {
        __syncthreads();
        if (base + (int)threadIdx.x < n_bins) s_spec[threadIdx.x] = spec[(size_t)row * n_bins + base + threadIdx.x];
        __syncthreads();
        const int m = min(256, n_bins - base);
        for (int i = 0; i < m; ++i) {
            const int bin = bin0 + (base + i) * bstep;
            const C32 w = twL[idx];
            const C32 X = s_spec[i];
            const bool edge = bin == 0 || 2 * bin == L;
            const float term = edge ? X.x * w.x : 2.f * (X.x * w.x - X.y * w.y);
            const float yk = term - comp;
            const float sum = acc + yk;
            comp = (sum - acc) - yk;
            acc = sum;
            idx += inc;
            if (idx >= (unsigned)L) idx -= (unsigned)L;
        }
    }
    if (t < L) y[(size_t)row * L + t] = acc / (float)L;
}

}  // namespace mpb

extern "C" int mpb200_band_limit(const float* x, int rows, int n, int L, int bin0, int bin_step, int n_bins,
                                 float* y, void* stream) {
    using namespace mpb;
    if (!x || !y || rows < 1 || n < 1 || L < n || n_bins < 0 || bin_step < 1 || bin0 < 0)
        return fail(MPB200_EINVAL, "bad argument");
    if (L % 2) return fail(MPB200_EINVAL, "mpb200_band_limit: the transform length must be even (the reference's "
                                          "irfft returns L-1 samples for odd L)");
    if (n_bins > 0 && bin0 + (long long)(n_bins - 1) * bin_step > L / 2)
        return fail(MPB200_EINVAL, "bins beyond L/2");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_bins == 0) {
        BF_CUDA(cudaMemsetAsync(y, 0, (size_t)rows * L * sizeof(float), st));
        return MPB200_OK;
    }
    int dev = 0;
    BF_CUDA(cudaGetDevice(&dev));
    C32* twL = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_tab_mutex);
        const long long key = (3LL << 40) + L;
        auto it = g_tab.find({dev, key});
        if (it != g_tab.end()) {
            twL = it->second;
        } else {
            const long double two_pi = 6.283185307179586476925286766559005768L;
            std::vector<C32> h((size_t)L);
            for (int i = 0; i < L; ++i) {
                long double a = two_pi * (long double)i / (long double)L;
                h[i] = {(float)cosl(a), (float)sinl(a)};
            }
            int rc = upload(dev, key, h, &twL);
            if (rc) return rc;
        }
    }
    // spectrum scratch: stream-ordered allocation (freed on the same stream after the inverse kernel)
    C32* spec = nullptr;
    BF_CUDA(cudaMallocAsync((void**)&spec, (size_t)rows * n_bins * sizeof(C32), st));
    k_band_forward<<<dim3(n_bins, rows), 256, 0, st>>>(x, n, L, bin0, bin_step, twL, spec);
    BF_LAUNCH("k_band_forward");
    k_band_inverse<<<dim3((L + 255) / 256, rows), 256, 0, st>>>(spec, n_bins, L, bin0, bin_step, twL, y);
    BF_LAUNCH("k_band_inverse");
    BF_CUDA(cudaFreeAsync(spec, st));
    return MPB200_OK;
}
