// gemm_corr.cu -- the TENSOR-CORE route of the atom x residual correlation (BASELINE.json north_star (1)):
//
//     fm[b, k, t] = sum_i d[k, i] * x[b, t + i],   t in [0, N),  x zero beyond N
//                                                  (F.pad + F.conv1d + crop, modules/matchingpursuit.py:275-277)
//
// as a Toeplitz/Hankel GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory) with
// split-precision 3xTF32 so that fp32 argmax parity holds:  x = xh + xl, d = dh + dl with xh/dh the top 19 bits,
// and  x*d ~= xh*dh + xh*dl + xl*dh  (the dropped xl*dl term is below 2^-22 relative).
//
// One CTA computes a tile of 128 positions (MMA M, = TMEM lanes) x 256 atoms (MMA N, = TMEM columns), looping over
// the taps in blocks of 32 (four MMAs of K = 8 per operand pairing):
//   * A operand = the HANKEL tile H[t, i] = x[t0 + t + i0 + i].  Its rows overlap (row pitch one sample), which no
//     tensor-map or descriptor stride can express, so it is EXPANDED in shared memory: the signal segment is staged
//     once per CTA, and every tap block writes H (hi and lo halves) in the canonical K-major no-swizzle core-matrix
//     layout (8 rows x 16 bytes per core matrix).
//   * B operand = the dictionary tile D[k, i0 + i]: a pre-pass (k_gemm_pack_dict) splits the dictionary into hi and
//     lo halves and stores every (256 atoms x 32 taps) tile in the core-matrix layout, so a tile pair is ONE 64 KB bulk
//     asynchronous copy (TMA) per tap block instead of 48 stores per thread.
//   * one elected thread issues the 12 MMAs of the block and commits them to an mbarrier; the staging of the next
//     block (other shared-memory stage) runs under them.
//   * epilogue: tcgen05.ld 32 lanes x 32 columns per warp, coalesced stores (lanes = consecutive positions).
//
// This route exists to be MEASURED against the FFT route (profiles/r2_gemm_vs_fft.md, tools/bench_gemm.py);
// mpb200_plan_* never selects it on its own.
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/mpb200.h"
#include "plan.h"

namespace mpb {

extern std::atomic<unsigned long long> g_launches;

constexpr int GM = 128;            // positions per tile
constexpr int GN = 256;            // atoms per tile
constexpr int GK = 32;             // taps per staged block
constexpr int G_THREADS = 128;
constexpr int G_STAGES = 2;
// one stage: H hi/lo (GM x GK) + D hi/lo (GN x GK), fp32
constexpr int G_H_ELEMS = GM * GK, G_D_ELEMS = GN * GK;
constexpr int G_STAGE_BYTES = (2 * G_H_ELEMS + 2 * G_D_ELEMS) * 4;      // 96 KB

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Canonical K-major, no-swizzle operand layout: core matrices of 8 rows x 4 fp32 (16 bytes per row, 128 bytes per
// core matrix); the GK/4 core matrices of one 8-row group are consecutive (leading byte offset 128), 8-row groups
// follow each other (stride byte offset GK/4 * 128).  Element (row r, tap c):
__device__ __forceinline__ int g_core_off(int r, int c) { return ((r >> 3) * (GK / 4) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }

// 64-bit shared-memory matrix descriptor (sm_100): start address, LBO, SBO (all >> 4), version 1, no swizzle.
__device__ __forceinline__ uint64_t g_desc(uint32_t smem_addr) {
    const uint64_t lbo = 128 >> 4, sbo = ((GK / 4) * 128) >> 4;
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// 32-bit instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128.
constexpr uint32_t G_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);

__device__ __forceinline__ void g_mma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(a), "l"(b), "r"(G_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void g_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(g_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void g_split(float v, float& hi, float& lo) {
    hi = __int_as_float(__float_as_int(v) & 0xffffe000);     // the 19 bits TF32 keeps: exact, so lo = v - hi is exact too
    lo = v - hi;
}

// Dictionary pre-pass: packed[((kt * nkb + kb) * 2 + half) * G_D_ELEMS + core_off(r, c)] = hi / lo half of
// d[kt*GN + r, kb*GK + c] (zero beyond K and A).  grid = (nkb, K tiles), GN threads: thread = atom row.
__global__ void __launch_bounds__(GN)
k_gemm_pack_dict(const float* __restrict__ dict, int K, int A, int nkb, float* __restrict__ packed) {
    const int kb = blockIdx.x, kt = blockIdx.y, r = threadIdx.x;
    const int k = kt * GN + r, i0 = kb * GK;
    float* __restrict__ hi = packed + ((size_t)(kt * nkb + kb) * 2) * G_D_ELEMS;
    float* __restrict__ lo = hi + G_D_ELEMS;
    const float* __restrict__ dk = dict + (size_t)(k < K ? k : 0) * A;
    const bool vec = (A & 3) == 0 && k < K;
#pragma unroll
    for (int c1 = 0; c1 < GK / 4; ++c1) {
        float v[4];
        if (vec && i0 + 4 * c1 + 3 < A) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(dk + i0 + 4 * c1));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + 4 * c1 + e;
                v[e] = (k < K && i < A) ? __ldg(dk + i) : 0.f;
            }
        }
        float4 h, l;
        g_split(v[0], h.x, l.x); g_split(v[1], h.y, l.y); g_split(v[2], h.z, l.z); g_split(v[3], h.w, l.w);
        const int o = g_core_off(r, 4 * c1);
        *reinterpret_cast<float4*>(hi + o) = h;
        *reinterpret_cast<float4*>(lo + o) = l;
    }
}

struct GemmCorrArgs {
    const float* signal;   // (B, N)
    const float* packed;   // dictionary tiles from k_gemm_pack_dict
    float* out;            // (B, K, N)
    int N, K, A;
    int spin_mma;          // > 0: peak probe -- skip staging/epilogue traffic and issue this many extra MMA blocks
};

__global__ void __launch_bounds__(G_THREADS, 1)
k_corr_gemm(const GemmCorrArgs a) {
    extern __shared__ __align__(1024) unsigned char graw[];
    float* stage0 = reinterpret_cast<float*>(graw);
    __shared__ __align__(8) unsigned long long s_done[G_STAGES];     // the MMAs that read a stage have completed
    __shared__ __align__(8) unsigned long long s_full[G_STAGES];     // a stage's dictionary tiles have landed
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.x * GM, k0 = blockIdx.y * GN, b = blockIdx.z;
    const int nkb = (a.A + GK - 1) / GK;
    float* sx = stage0 + G_STAGES * (G_STAGE_BYTES / 4);      // signal segment x[t0 .. t0 + GM + nkb*GK)
    const int seg = GM + nkb * GK;
    const float* __restrict__ x = a.signal + (size_t)b * a.N;
    for (int i = tid; i < seg; i += G_THREADS) sx[i] = (t0 + i < a.N) ? x[t0 + i] : 0.f;
    if (tid == 0) {
        for (int s = 0; s < G_STAGES; ++s) { g_mbar_init(&s_done[s], 1); g_mbar_init(&s_full[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(&s_tmem)), "r"(GN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    unsigned phase[G_STAGES] = {0, 0}, fphase[G_STAGES] = {0, 0};
    const int nkt_y = blockIdx.y;
    const int total_blocks = nkb + (a.spin_mma > 0 ? a.spin_mma : 0);
    for (int kb = 0; kb < total_blocks; ++kb) {
        const int s = kb % G_STAGES;
        float* Hh = stage0 + (size_t)s * (G_STAGE_BYTES / 4);
        float* Hl = Hh + G_H_ELEMS;
        float* Dh = Hl + G_H_ELEMS;
        float* Dl = Dh + G_D_ELEMS;
        if (kb >= G_STAGES) {                       // the MMAs that read this stage two blocks ago have completed
            g_mbar_wait(&s_done[s], phase[s]);
            phase[s] ^= 1u;
        }
        if (kb < nkb && tid == 0) {
            // dictionary tiles (hi + lo, adjacent in the stage): one bulk copy, issued first so that it lands under the
            // Hankel expansion below
            const unsigned bytes = 2u * G_D_ELEMS * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(&s_full[s])), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(g_smem_u32(Dh)), "l"(a.packed + ((size_t)(nkt_y * nkb + kb) * 2) * G_D_ELEMS), "r"(bytes),
                           "r"(g_smem_u32(&s_full[s])) : "memory");
        }
        if (kb < nkb) {
            const int i0 = kb * GK;
            // Hankel tile: thread = row (position); 8 x 16-byte stores per half
            {
                const int r = tid;
#pragma unroll
                for (int c1 = 0; c1 < GK / 4; ++c1) {
                    float4 h, l;
                    g_split(sx[r + i0 + 4 * c1 + 0], h.x, l.x);
                    g_split(sx[r + i0 + 4 * c1 + 1], h.y, l.y);
                    g_split(sx[r + i0 + 4 * c1 + 2], h.z, l.z);
                    g_split(sx[r + i0 + 4 * c1 + 3], h.w, l.w);
                    const int o = g_core_off(r, 4 * c1);
                    *reinterpret_cast<float4*>(Hh + o) = h;
                    *reinterpret_cast<float4*>(Hl + o) = l;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            if (kb < nkb) {
                g_mbar_wait(&s_full[s], fphase[s]);
                fphase[s] ^= 1u;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t hh = g_smem_u32(Hh), hl = g_smem_u32(Hl), dh = g_smem_u32(Dh), dl = g_smem_u32(Dl);
#pragma unroll
            for (int ks = 0; ks < GK / 8; ++ks) {
                const uint32_t adv = ks * 2 * 128;                       // two core matrices (8 taps) further along K
                g_mma(tmem, g_desc(hh + adv), g_desc(dh + adv), (kb | ks) ? 1u : 0u);
                g_mma(tmem, g_desc(hh + adv), g_desc(dl + adv), 1u);
                g_mma(tmem, g_desc(hl + adv), g_desc(dh + adv), 1u);
            }
            g_commit(&s_done[s]);
        }
    }
    // drain: every stage's last commit
    for (int s = 0; s < G_STAGES; ++s) {
        const int uses = (total_blocks - s + G_STAGES - 1) / G_STAGES;    // blocks that went through stage s
        if (uses > 0) g_mbar_wait(&s_done[s], phase[s]);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 = positions t0 + 32w + lane; columns = atoms
    if (a.spin_mma <= 0) {
        const int t = t0 + warp * 32 + lane;
        for (int c0 = 0; c0 < GN; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (t < a.N) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int k = k0 + c0 + j;
                    if (k < a.K) a.out[((size_t)b * a.K + k) * a.N + t] = __uint_as_float(v[j]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(GN));
}

}  // namespace mpb

using namespace mpb;

extern "C" int mpb200_correlate_gemm(const float* signal, int batch, int n_samples, const float* d, int n_atoms,
                                     int atom_size, float* fm_out, int spin_blocks, void* stream) {
    if (!signal || !d || batch < 1 || n_samples < 1 || n_atoms < 1 || atom_size < 1 || (!fm_out && spin_blocks <= 0))
        return fail(MPB200_EINVAL, "bad argument");
    if (batch > 65535) return fail(MPB200_EINVAL, "batch must be <= 65535");
    cudaStream_t st = (cudaStream_t)stream;
    const int nkb_h = (atom_size + GK - 1) / GK, nkt = (n_atoms + GN - 1) / GN;
    float* packed = nullptr;
    const size_t packed_elems = (size_t)nkt * nkb_h * 2 * G_D_ELEMS;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) keep_async_pool(dev);
    }
    cudaError_t e = cudaMallocAsync((void**)&packed, packed_elems * sizeof(float), st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MPB200_ENOMEM, std::string("packed dictionary: ") + cudaGetErrorString(e));
    }
    k_gemm_pack_dict<<<dim3(nkb_h, nkt), GN, 0, st>>>(d, n_atoms, atom_size, nkb_h, packed);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GemmCorrArgs a;
    a.signal = signal;
    a.packed = packed;
    a.out = fm_out;
    a.N = n_samples;
    a.K = n_atoms;
    a.A = atom_size;
    a.spin_mma = spin_blocks;
    const int nkb = (atom_size + GK - 1) / GK;
    const size_t smem = (size_t)G_STAGES * G_STAGE_BYTES + (size_t)(GM + nkb * GK) * sizeof(float);
    if (smem > 227 * 1024) {
        cudaFreeAsync(packed, st);
        return fail(MPB200_EINVAL, "atom too long for the staged signal segment of the GEMM route");
    }
    e = cudaFuncSetAttribute(k_corr_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        dim3 grid((n_samples + GM - 1) / GM, (n_atoms + GN - 1) / GN, batch);
        k_corr_gemm<<<grid, G_THREADS, smem, st>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    cudaFreeAsync(packed, st);
    if (e != cudaSuccess) return fail(MPB200_ECUDA, std::string("k_corr_gemm: ") + cudaGetErrorString(e));
    return MPB200_OK;
}
