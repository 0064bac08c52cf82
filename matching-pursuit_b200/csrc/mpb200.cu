// mpb200.cu -- plan management and the extern "C" entry points of include/mpb200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
//             -Xcompiler -fPIC -shared -o libmpb200.so mpb200.cu fftconv.cu
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mpb200.h"
#include "kernels.cuh"
#include "fused_loop.cuh"
#include "plan.h"

namespace mpb {

thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};   // kernels launched by this library (process-wide)

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

void keep_async_pool(int device) {
    static std::mutex mu;
    static bool done[64] = {false};
    if (device < 0 || device >= 64) return;
    std::lock_guard<std::mutex> lock(mu);
    if (done[device]) return;
    done[device] = true;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
}

#define MPB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(MPB200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)

#define MPB_LAUNCH_CHECK(name)                                                                  \
    do {                                                                                        \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                     \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return fail(MPB200_ECUDA, std::string("launch ") + name + ": " + cudaGetErrorString(_e)); \
    } while (0)

#ifndef MPB_GRAM_BATCH_FACTOR
#define MPB_GRAM_BATCH_FACTOR 64     // AUTO takes the Gram table when max_batch * this >= n_atoms (and it fits)
#endif
#ifndef MPB_PDL
#define MPB_PDL 1     // programmatic dependent launch of the iteration-loop kernels (0: ordinary stream order)
#endif

#define MPB_DISPATCH_M(m, ...)                                            \
    switch (m) {                                                          \
        case 512: { constexpr int MM = 512; __VA_ARGS__; } break;         \
        case 1024: { constexpr int MM = 1024; __VA_ARGS__; } break;       \
        case 2048: { constexpr int MM = 2048; __VA_ARGS__; } break;       \
        case 4096: { constexpr int MM = 4096; __VA_ARGS__; } break;       \
        case 8192: { constexpr int MM = 8192; __VA_ARGS__; } break;       \
        default: return fail(MPB200_EINVAL, "unsupported FFT size");      \
    }

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is per device and per kernel, so the
// largest size granted so far is remembered per (device, kernel address); plans on several devices of one
// process, or plans with different staging sizes, each get what they need.
template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(kernel));
    std::lock_guard<std::mutex> lock(mu);
    auto it = granted.find(key);
    if (it != granted.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) granted[key] = bytes;
    return e;
}

// Launch of an iteration-loop kernel (one struct argument) with programmatic stream serialisation allowed: it may be
// scheduled while its predecessor still runs and waits for it on the device (kernels.cuh, pdl_prologue).
template <typename Arg>
static cudaError_t launch_pdl(void (*kernel)(Arg), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Arg& arg) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = MPB_PDL;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, arg);
}

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------
int choose_fft_size(int A) {
    // a step window refreshes up to span = ceil((2A-2)/blk)+1 blocks of blk = M/32 outputs and
    // needs span*blk + A - 1 <= 3A - 3 + M/16 input samples  ->  M >= (3A-3)*16/15
    const long long need = ((long long)(3 * A - 3) * 16 + 14) / 15;
    for (int m = 512; m <= 8192; m *= 2)
        if (m >= need && m >= A + m / 32) return m;
    return 0;
}

// Entries per atom of the SGRAM winner spectra: BlockFft<M2>::SMEM_CPX when k_delta bulk-copies them into its FFT
// buffer (staged layout, MPB_DELTA_SPREF), else M2.
static size_t spec_stride(int M2) {
    if (!MPB_DELTA_SPREF) return (size_t)M2;
    const int R1 = M2 / 256, S = R1 | 1, XS = 16 * S + (R1 < 16 ? R1 : 0);
    return (size_t)16 * XS;
}

template <typename Real>
static void host_twiddles(int M, std::vector<cpx<Real>>& t1, std::vector<cpx<Real>>& t2) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    const int R1 = M / 256;
    t1.resize(M);
    t2.resize(256);
    for (int m1 = 0; m1 < R1; ++m1)
        for (int c = 0; c < 256; ++c) {
            long double a = two_pi * (long double)(((long long)c * m1) % M) / (long double)M;
            t1[m1 * 256 + c] = {(Real)cosl(a), (Real)sinl(a)};
        }
    for (int m2 = 0; m2 < 16; ++m2)
        for (int j3 = 0; j3 < 16; ++j3) {
            long double a = two_pi * (long double)((j3 * m2) % 256) / 256.0L;
            t2[m2 * 16 + j3] = {(Real)cosl(a), (Real)sinl(a)};
        }
}

template <typename T>
static int dev_alloc(Plan* p, T** ptr, size_t count) {
    *ptr = nullptr;
    if (count == 0) return MPB200_OK;
    cudaError_t e = cudaMalloc((void**)ptr, count * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MPB200_ENOMEM, "cudaMalloc of " + std::to_string(count * sizeof(T)) + " bytes failed: " +
                                       cudaGetErrorString(e));
    }
    p->allocs.push_back((void*)*ptr);
    p->bytes += count * sizeof(T);
    return MPB200_OK;
}

static void free_plan(Plan* p) {
    for (void* m : p->ipc_opened) cudaIpcCloseMemHandle(m);
    for (void* e : p->ev_pool) cudaEventDestroy((cudaEvent_t)e);
    for (void* q : p->allocs) cudaFree(q);
    if (p->h_stage) cudaFreeHost(p->h_stage);
    delete p;
}

// ---------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------
static int launch_window_fft_ex(int M, const C32* tw1, const C32* tw2, const float* src, long long row_stride,
                                int row_len, const Win* win, int nwin, C32* winspec, cudaStream_t st,
                                const int* skip = nullptr, bool staged = false) {
    if (nwin <= 0) return MPB200_OK;
    MPB_DISPATCH_M(M, {
        using F = BlockFft<MM, float>;
        const size_t smem = (size_t)(F::SMEM_CPX + 256) * sizeof(C32);
        if (staged) {
            constexpr int FORM = DeltaCfg<MM>::LOCAL ? 2 : 1;      // the layout k_delta<MM> reads
            MPB_CUDA((allow_smem(k_window_fft<MM, FORM>, smem)));
            k_window_fft<MM, FORM><<<nwin, F::T, smem, st>>>(src, row_stride, row_len, win, tw1, tw2, winspec, skip);
        } else {
            MPB_CUDA((allow_smem(k_window_fft<MM, 0>, smem)));
            k_window_fft<MM, 0><<<nwin, F::T, smem, st>>>(src, row_stride, row_len, win, tw1, tw2, winspec, skip);
        }
    });
    MPB_LAUNCH_CHECK("k_window_fft");
    return MPB200_OK;
}

static int launch_window_fft(Plan* p, const float* src, long long row_stride, int row_len, const Win* win, int nwin,
                             C32* winspec, cudaStream_t st, const int* skip = nullptr) {
    return launch_window_fft_ex(p->M, p->tw1, p->tw2, src, row_stride, row_len, win, nwin, winspec, st, skip);
}

template <int MODE>
static int launch_corr(Plan* p, CorrArgs a, int groups, cudaStream_t st, bool pdl = false, int max_gx = 0) {
    if (a.nwin <= 0) return MPB200_OK;
    a.bm_cap = p->bm_cap;
    MPB_DISPATCH_M(p->M, {
        using F = BlockFft<MM, float>;
        constexpr int TPB = F::T < 256 ? 256 : F::T;
        constexpr int NT = TPB / F::T;
        size_t smem = (size_t)(256 + NT * F::SMEM_CPX) * sizeof(C32);
        if (MPB_CORR_SEP && (MODE & MODE_ROWMAX) != 0) smem += (size_t)NT * (p->bm_cap + 64) * sizeof(float2);
        MPB_CUDA(allow_smem(k_corr<MM, MODE>, smem));
        dim3 grid((a.npairs + NT - 1) / NT, groups);
        if (max_gx > 0 && (int)grid.x > max_gx) grid.x = max_gx;      // the CTAs walk the remaining pair groups
        if (pdl) MPB_CUDA(launch_pdl(k_corr<MM, MODE>, grid, dim3(TPB), smem, st, a));
        else k_corr<MM, MODE><<<grid, TPB, smem, st>>>(a);
    });
    MPB_LAUNCH_CHECK("k_corr");
    return MPB200_OK;
}

static int corr_groups(const Plan* p, int nwin) {
    const int tpb_pairs = (p->M >= 4096) ? 1 : (4096 / p->M > 8 ? 8 : 4096 / p->M);  // NT of k_corr
    const int gx = (p->npairs + tpb_pairs - 1) / tpb_pairs;
    int g = (p->sm_count * 8 + gx - 1) / gx;
    if (g < 1) g = 1;
    if (g > nwin) g = nwin;
    if (g > 65535) g = 65535;
    return g;
}

// Grid depth of the FFT route inside the map modes (winners that overhang the right edge): usually no window
// at all, so the launch is kept just deep enough to fill the chip twice instead of eight times.
static int trunc_groups(const Plan* p, int nwin) {
    const int tpb_pairs = (p->M >= 4096) ? 1 : (4096 / p->M > 8 ? 8 : 4096 / p->M);
    int gx = (p->npairs + tpb_pairs - 1) / tpb_pairs;
    if (gx > p->sm_count) gx = p->sm_count;           // launch_corr caps the x extent there for this route
    int g = (p->sm_count * 2 + gx - 1) / gx;
    if (g < 1) g = 1;
    if (g > nwin) g = nwin;
    return g;
}

static CorrArgs base_corr_args(const Plan* p) {
    CorrArgs a;
    memset(&a, 0, sizeof(a));
    a.winspec = p->winspec;
    a.pairspec = p->pairspec;
    a.npairs = p->npairs;
    a.nloc = p->nloc;
    a.len = p->N;
    a.NB = p->NB;
    a.blk_shift = p->blk_shift;
    a.tw1 = p->tw1;
    a.tw2 = p->tw2;
    a.bm_val = p->bm_val;
    a.bm_pos = p->bm_pos;
    a.row_val = p->row_val;
    a.row_pos = p->row_pos;
    return a;
}

// Full correlation of the current residual: block maxima (and the dense map
// when `dense` is given) for all positions, in slabs of at most wcap windows.
static int full_pass(Plan* p, int batch, float* dense, cudaStream_t st, bool with_maxima = false) {
    const int total = batch * p->nchunks;
    for (int w0 = 0; w0 < total; w0 += p->wcap) {
        const int n = total - w0 < p->wcap ? total - w0 : p->wcap;
        int rc = launch_window_fft(p, p->residual, p->N, p->N, p->win_full + w0, n, p->winspec, st);
        if (rc) return rc;
        CorrArgs a = base_corr_args(p);
        a.win = p->win_full + w0;
        a.nwin = n;
        if (dense) {
            const long long rs = (dense == p->map) ? p->NS : p->N;     // the resident map may have padded rows
            a.dense = dense;
            a.dense_row_stride = (long long)p->nloc * rs;
            a.dense_atom_stride = rs;
            a.dense_col_off = 0;
            if (with_maxima) rc = launch_corr<MODE_DENSE | MODE_BLOCKMAX>(p, a, corr_groups(p, n), st);
            else rc = launch_corr<MODE_DENSE>(p, a, corr_groups(p, n), st);
        } else {
            rc = launch_corr<MODE_BLOCKMAX>(p, a, corr_groups(p, n), st);
        }
        if (rc) return rc;
    }
    if (!dense || with_maxima) {
        const int rows = batch * p->nloc;
        k_rowmax<<<(rows + 7) / 8, 256, 0, st>>>(p->bm_val, p->bm_pos, rows, p->NB, p->row_val, p->row_pos);
        MPB_LAUNCH_CHECK("k_rowmax");
    }
    return MPB200_OK;
}

template <bool SELECT>
static int launch_apply(Plan* p, int batch, const Best* winner, int step, int n_steps, int* atom_out, int* pos_out,
                        float* val_out, int do_fft, cudaStream_t st) {
    ApplyArgs a;
    memset(&a, 0, sizeof(a));
    a.row_val = p->row_val;
    a.row_pos = p->row_pos;
    a.winner = winner;
    a.dict = p->dict;
    a.residual = p->residual;
    a.nloc = p->nloc;
    a.atom_lo = p->lo;
    a.n_atoms = p->K;
    a.A = p->A;
    a.N = p->N;
    a.blk_shift = p->blk_shift;
    a.NB = p->NB;
    a.step = step;
    a.n_steps = n_steps;
    a.atom_out = atom_out;
    a.pos_out = pos_out;
    a.val_out = val_out;
    a.win = p->win_step;
    a.tw1 = p->tw1;
    a.tw2 = p->tw2;
    a.winspec = p->winspec;
    a.do_fft = do_fft;
    a.gram = p->mode == MPB200_MODE_GRAM || p->mode == MPB200_MODE_SGRAM;
    a.world = 1;
    if (SELECT && p->xconnected && p->xworld > 1) {
        a.peer_mail = p->peer_mail;
        a.world = p->xworld;
        a.rank = p->xrank;
        a.mail_batch = p->Bmax;
        a.seq = ++p->xseq;
        if (p->xseq == 0xffffffffu) p->xseq = 0;     // 0 is the "empty" tag of a fresh mailbox
        a.xerr = p->xerr;
    }
    a.upd = p->upd;
    a.trunc_count = p->trunc_count;
    a.parity = (int)(p->iter & 1u);
    a.map = (SELECT && p->pos_free) ? p->map : nullptr;
    a.NS = p->NS;
    if (SELECT && p->lcn) {            // select on the normalised tables, report the raw map value
        a.row_val = p->nrow_val;
        a.row_pos = p->nrow_pos;
        a.raw_map = p->map;
    }
    MPB_DISPATCH_M(p->M, {
        using F = BlockFft<MM, float>;
        const size_t smem = (size_t)(F::SMEM_CPX + 256) * sizeof(C32);
        MPB_CUDA((allow_smem(k_apply<MM, SELECT>, smem)));
        MPB_CUDA(launch_pdl(k_apply<MM, SELECT>, dim3(batch), dim3(256), smem, st, a));
    });
    MPB_LAUNCH_CHECK("k_apply");
    return MPB200_OK;
}

// timing marker: everything enqueued on `st` since the previous marker is attributed to `tag`
static void mark(Plan* p, int tag, cudaStream_t st) {
    if (!p->timing) return;
    if (p->ev_used == p->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
        p->ev_pool.push_back((void*)e);
    }
    cudaEventRecord((cudaEvent_t)p->ev_pool[p->ev_used++], st);
    p->ev_tag.push_back(tag);
}

// LCN selection: refresh the normalised block / row maxima from the resident raw map (whole rows after a full pass,
// the winners' windows after a step).
static int lcn_refresh(Plan* p, int batch, bool full, cudaStream_t st) {
    LcnArgs l;
    l.map = p->map;
    l.upd = p->upd;
    l.nbm_val = p->nbm_val;
    l.nbm_pos = p->nbm_pos;
    l.nrow_val = p->nrow_val;
    l.nrow_pos = p->nrow_pos;
    l.nloc = p->nloc;
    l.N = p->N;
    l.NS = p->NS;
    l.NB = p->NB;
    l.blk_shift = p->blk_shift;
    l.A = p->A;
    l.full = full ? 1 : 0;
    const size_t smem = (size_t)16 * (p->blk + 8) * sizeof(float);
    dim3 grid((p->nloc + 7) / 8, batch);
    MPB_CUDA(launch_pdl(k_lcn_refresh, grid, dim3(256), smem, st, l));
    MPB_LAUNCH_CHECK("k_lcn_refresh");
    return MPB200_OK;
}

static int step_refresh_raw(Plan* p, int batch, cudaStream_t st);

// Refresh the map / block maxima / row maxima after k_apply's subtraction.
static int step_refresh(Plan* p, int batch, cudaStream_t st) {
    const bool full_refresh = (p->mode == MPB200_MODE_GRAM || p->mode == MPB200_MODE_SGRAM) && p->refresh_every > 0 &&
                              (p->iter + 1) % (unsigned)p->refresh_every == 0;
    int rc = step_refresh_raw(p, batch, st);
    if (!rc && p->lcn) rc = lcn_refresh(p, batch, full_refresh, st);
    return rc;
}

static int step_refresh_raw(Plan* p, int batch, cudaStream_t st) {
    int rc = MPB200_OK;
    if (p->mode == MPB200_MODE_FULL) {
        rc = full_pass(p, batch, nullptr, st);
    } else if (p->mode == MPB200_MODE_SGRAM) {
        const bool refresh = p->refresh_every > 0 && (p->iter + 1) % (unsigned)p->refresh_every == 0;
        if (refresh) {
            rc = full_pass(p, batch, p->map, st, true);
        } else {
            DeltaArgs d;
            d.atomspec = p->atomspec;
            d.pairspec2 = p->pairspec2;
            d.upd = p->upd;
            d.batch = batch;
            d.npairs = p->npairs;
            d.nloc = p->nloc;
            d.map = p->map;
            d.N = p->N;
            d.NS = p->NS;
            d.cap = p->bm_cap > p->M2 + p->blk ? p->bm_cap : p->M2 + p->blk;   // see k_delta: in-bounds update indices
            d.NB = p->NB;
            d.blk_shift = p->blk_shift;
            d.A = p->A;
            d.tw1 = p->tw1b;
            d.tw2 = p->tw2;
            d.bm_val = p->bm_val;
            d.bm_pos = p->bm_pos;
            d.row_val = p->row_val;
            d.row_pos = p->row_pos;
            MPB_DISPATCH_M(p->M2, {
                using F = typename DeltaCfg<MM>::F;
                constexpr int TPB = DeltaCfg<MM>::TPB;
                constexpr int NT = TPB / F::T;
                // the double-buffered two-CTA form pays when a CTA meets the same pair spectrum for many items in a row
                const bool db = DeltaCfg<MM, true>::DB && batch >= MPB_DELTA_DB_MIN_BATCH;
                const size_t smem = (size_t)(256 + NT * F::SMEM_CPX) * sizeof(C32) +
                                    (size_t)NT * 2 * (db ? 2 : 1) * d.cap * sizeof(float);
                void (*kernel)(DeltaArgs) = p->pos_free ? k_delta<MM, true, false> : k_delta<MM, false, false>;
                if constexpr (DeltaCfg<MM, true>::DB) {
                    if (db) kernel = p->pos_free ? k_delta<MM, true, true> : k_delta<MM, false, true>;
                }
                MPB_CUDA(allow_smem(kernel, smem));
                int& occ = db ? p->delta_occ_db : p->delta_occ;
                if (occ == 0) {
                    MPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, TPB, smem));
                    if (occ < 1) occ = 1;
                }
                d.ngroups = (p->npairs + NT - 1) / NT;
                const long long items = (long long)d.ngroups * batch;
                if (items >= (1LL << 31)) return fail(MPB200_EINVAL, "SGRAM: more than 2^31 (pair, signal) work items per launch");
                long long ctas = (long long)p->sm_count * occ;
                if (ctas > items) ctas = items;
                MPB_CUDA(launch_pdl(kernel, dim3((unsigned)ctas), dim3(TPB), smem, st, d));
            });
            MPB_LAUNCH_CHECK("k_delta");
            mark(p, 4, st);
            CorrArgs a = base_corr_args(p);
            a.win = p->win_step;
            a.nwin = batch;
            a.nwin_ptr = p->trunc_count + (p->iter & 1u);
            a.dense = p->map;
            a.dense_row_stride = (long long)p->nloc * p->NS;
            a.dense_atom_stride = p->NS;
            a.dense_col_off = 0;
            a.pos_free = p->pos_free ? 1 : 0;
            rc = launch_corr<MODE_DENSE | MODE_BLOCKMAX | MODE_ROWMAX>(p, a, trunc_groups(p, batch), st, true, p->sm_count);
        }
    } else if (p->mode == MPB200_MODE_GRAM) {
        const bool refresh = p->refresh_every > 0 && (p->iter + 1) % (unsigned)p->refresh_every == 0;
        if (refresh) {
            rc = full_pass(p, batch, p->map, st, true);
        } else {
            GramArgs g;
            g.map = p->map;
            g.gram = p->gram;
            g.upd = p->upd;
            g.bm_val = p->bm_val;
            g.bm_pos = p->bm_pos;
            g.row_val = p->row_val;
            g.row_pos = p->row_pos;
            g.rows = batch * p->nloc;
            g.nloc = p->nloc;
            g.N = p->N;
            g.NB = p->NB;
            g.blk_shift = p->blk_shift;
            g.A = p->A;
            g.GS = p->GS;
            switch (p->blk) {
                case 16: MPB_CUDA(launch_pdl(k_gram_update<16>, dim3((g.rows + 7) / 8), dim3(256), 0, st, g)); break;
                case 32: MPB_CUDA(launch_pdl(k_gram_update<32>, dim3((g.rows + 7) / 8), dim3(256), 0, st, g)); break;
                case 64: MPB_CUDA(launch_pdl(k_gram_update<64>, dim3((g.rows + 7) / 8), dim3(256), 0, st, g)); break;
                case 128: MPB_CUDA(launch_pdl(k_gram_update<128>, dim3((g.rows + 7) / 8), dim3(256), 0, st, g)); break;
                case 256: MPB_CUDA(launch_pdl(k_gram_update<256>, dim3((g.rows + 7) / 8), dim3(256), 0, st, g)); break;
                default: return fail(MPB200_EINVAL, "unsupported block size");
            }
            MPB_LAUNCH_CHECK("k_gram_update");
            mark(p, 4, st);
            // winners truncated at the right edge: FFT re-correlation of their windows, written into the map
            CorrArgs a = base_corr_args(p);
            a.win = p->win_step;
            a.nwin = batch;
            a.nwin_ptr = p->trunc_count + (p->iter & 1u);
            a.dense = p->map;
            a.dense_row_stride = (long long)p->nloc * p->N;
            a.dense_atom_stride = p->N;
            a.dense_col_off = 0;
            rc = launch_corr<MODE_DENSE | MODE_BLOCKMAX | MODE_ROWMAX>(p, a, trunc_groups(p, batch), st, true, p->sm_count);
        }
    } else {
        CorrArgs a = base_corr_args(p);
        a.win = p->win_step;
        a.nwin = batch;
        rc = launch_corr<MODE_BLOCKMAX | MODE_ROWMAX>(p, a, corr_groups(p, batch), st, true);
    }
    p->iter++;
    return rc;
}

// Gram table: correlation of every atom (left-padded by A-1 zeros) with every owned atom,
// through the same window FFT + pair-spectrum machinery:  gram[k][j][l] = sum_i d_k[i + l - (A-1)] d_j[i].
static int build_gram(Plan* p, cudaStream_t st) {
    for (int k0 = 0; k0 < p->K; k0 += p->wcap) {
        const int n = p->K - k0 < p->wcap ? p->K - k0 : p->wcap;
        int rc = launch_window_fft(p, p->dict, p->A, p->A, p->win_gram + k0, n, p->winspec, st, p->dict_skip);
        if (rc) return rc;
        CorrArgs a = base_corr_args(p);
        a.skip = p->dict_skip;
        a.win = p->win_gram + k0;
        a.nwin = n;
        a.len = p->A;                      // valid outputs: t0 + m < len  <=>  m < 2A-1   (t0 = -(A-1))
        a.dense = p->gram;
        a.dense_row_stride = (long long)p->nloc * p->GS;
        a.dense_atom_stride = p->GS;
        a.dense_col_off = p->A - 1;        // column = t0 + m + (A-1) = m
        rc = launch_corr<MODE_DENSE>(p, a, corr_groups(p, n), st);
        if (rc) return rc;
    }
    return MPB200_OK;
}

static int build_pair_spectra(Plan* p, cudaStream_t st) {
    MPB_DISPATCH_M(p->M, {
        using F = BlockFft<MM, double>;
        const size_t smem = (size_t)(F::SMEM_CPX + 256) * sizeof(cpx<double>);
        MPB_CUDA((allow_smem(k_pair_spectra<MM, double>, smem)));
        k_pair_spectra<MM, double><<<p->npairs, F::T, smem, st>>>(p->dict, p->A, p->lo, p->hi, p->tw1d, p->tw2d,
                                                                   p->pairspec, p->dict_skip);
    });
    MPB_LAUNCH_CHECK("k_pair_spectra");
    return MPB200_OK;
}

// SGRAM tables: pair spectra at M2 and forward spectra of every atom left-padded by A-1 zeros.
static int build_sgram_tables(Plan* p, cudaStream_t st) {
    MPB_DISPATCH_M(p->M2, {
        using F = BlockFft<MM, double>;
        const size_t smem = (size_t)(F::SMEM_CPX + 256) * sizeof(cpx<double>);
        constexpr bool LOCAL = DeltaCfg<MM>::LOCAL;                // the order k_delta<MM> loads the table in
        MPB_CUDA((allow_smem(k_pair_spectra<MM, double, LOCAL>, smem)));
        k_pair_spectra<MM, double, LOCAL><<<p->npairs, F::T, smem, st>>>(p->dict, p->A, p->lo, p->hi, p->tw1bd, p->tw2d,
                                                                          p->pairspec2, p->dict_skip);
    });
    MPB_LAUNCH_CHECK("k_pair_spectra");
    return launch_window_fft_ex(p->M2, p->tw1b, p->tw2, p->dict, p->A, p->A, p->win_atoms, p->K, p->atomspec, st,
                                p->dict_skip, MPB_DELTA_SPREF != 0);
}

// With CUDA's lazy module loading the FIRST use of a kernel may synchronise the context.  A pursuit whose
// kernels wait for other ranks must never meet such a load while a waiting kernel is resident, so every
// kernel the plan can launch is loaded up front.
template <typename K>
static void touch(K kernel) {
    cudaFuncAttributes at;
    if (cudaFuncGetAttributes(&at, kernel) != cudaSuccess) cudaGetLastError();
}
static int preload_kernels(Plan* p) {
    MPB_DISPATCH_M(p->M, {
        touch(k_apply<MM, true>);
        touch(k_apply<MM, false>);
        touch(k_window_fft<MM, 0>);
        touch(k_corr<MM, MODE_BLOCKMAX>);
        touch(k_corr<MM, MODE_DENSE>);
        touch(k_corr<MM, MODE_DENSE | MODE_BLOCKMAX>);
        touch(k_corr<MM, MODE_BLOCKMAX | MODE_ROWMAX>);
        touch(k_corr<MM, MODE_DENSE | MODE_BLOCKMAX | MODE_ROWMAX>);
    });
    if (p->mode == MPB200_MODE_SGRAM) {
        MPB_DISPATCH_M(p->M2, {
            touch(k_delta<MM, true>); touch(k_delta<MM, false>); touch(k_window_fft<MM, 1>); touch(k_window_fft<MM, 2>);
            if constexpr (DeltaCfg<MM, true>::DB) { touch(k_delta<MM, true, true>); touch(k_delta<MM, false, true>); }
        });
    }
    touch(k_gram_update<16>);
    touch(k_gram_update<32>);
    touch(k_gram_update<64>);
    touch(k_gram_update<128>);
    touch(k_gram_update<256>);
    touch(k_rowmax);
    touch(k_local_best);
    touch(k_reduce_best);
    return MPB200_OK;
}

// Device staging of the host-buffer entry point (mpb200_sparse_code_host): the signals of a whole batch and
// (atom, position, value) for `steps` iterations of every signal.  Sized at plan creation for
// MPB200_DEFAULT_MAX_STEPS iterations; MPB200_OPT_MAX_STEPS re-sizes it ahead of time.  A call that needs more
// re-sizes it on the spot (the one allocation the library makes after plan creation, like a vector that grows);
// the outgrown buffers are released at once.
static int reserve_host_staging(Plan* p, size_t steps) {
    const size_t ev = (size_t)p->Bmax * steps;
    if (p->d_signal && ev <= p->ev_cap) return MPB200_OK;
    if (p->d_atom) {
        MPB_CUDA(cudaDeviceSynchronize());
        for (void* q : {(void*)p->d_atom, (void*)p->d_pos, (void*)p->d_val}) {
            cudaFree(q);
            for (auto it = p->allocs.begin(); it != p->allocs.end(); ++it)
                if (*it == q) { p->allocs.erase(it); break; }
        }
        p->bytes -= p->ev_cap * 12;
        p->d_atom = nullptr; p->d_pos = nullptr; p->d_val = nullptr;
        p->ev_cap = 0;
    }
    int rc = MPB200_OK;
    if (!p->d_signal) rc = dev_alloc(p, &p->d_signal, (size_t)p->Bmax * p->N);
    if (!rc) rc = dev_alloc(p, &p->d_atom, ev);
    if (!rc) rc = dev_alloc(p, &p->d_pos, ev);
    if (!rc) rc = dev_alloc(p, &p->d_val, ev);
    if (rc) return rc;
    p->ev_cap = ev;
    return MPB200_OK;
}

static int check_plan(Plan* p, bool need_dict) {
    if (!p) return fail(MPB200_EINVAL, "null plan");
    if (need_dict && !p->dict_set) return fail(MPB200_ESTATE, "mpb200_plan_set_dictionary has not been called");
    cudaError_t e = cudaSetDevice(p->device);
    if (e != cudaSuccess) return fail(MPB200_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return MPB200_OK;
}

}  // namespace mpb

using namespace mpb;

extern "C" {

int mpb200_version(void) { return MPB200_VERSION; }

const char* mpb200_last_error(void) { return g_last_error.c_str(); }

unsigned long long mpb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mpb200_plan_create(mpb200_plan_t* out, int n_atoms, int atom_size, int n_samples, int max_batch, int mode,
                       int atom_lo, int atom_hi, uint64_t gram_budget_bytes) {
    if (!out) return fail(MPB200_EINVAL, "null plan pointer");
    *out = nullptr;
    if (n_atoms < 1 || atom_size < 1 || n_samples < 1 || max_batch < 1)
        return fail(MPB200_EINVAL, "n_atoms, atom_size, n_samples and max_batch must be positive");
    if (atom_lo == 0 && atom_hi == 0) atom_hi = n_atoms;
    if (atom_lo < 0 || atom_hi > n_atoms || atom_lo >= atom_hi)
        return fail(MPB200_EINVAL, "atom shard [atom_lo, atom_hi) must be a non-empty sub-range of [0, n_atoms)");
    if (mode < MPB200_MODE_AUTO || mode > MPB200_MODE_SGRAM) return fail(MPB200_EINVAL, "unknown mode");
    const int M = choose_fft_size(atom_size);
    if (M == 0)
        return fail(MPB200_EINVAL, "atom_size " + std::to_string(atom_size) +
                                       " needs a window FFT longer than 8192 (supported: atom_size <= 2560)");
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(MPB200_ECUDA, "no CUDA device: this library has no CPU path");
    int device = 0, sm_count = 0;
    size_t free_b = 0, total_b = 0;
    MPB_CUDA(cudaGetDevice(&device));
    MPB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    MPB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    Plan* p = new Plan();           // from here on every failure path goes through free_plan
    p->device = device;
    p->sm_count = sm_count;
    p->K = n_atoms;
    p->A = atom_size;
    p->N = n_samples;
    p->Bmax = max_batch;
    p->lo = atom_lo;
    p->hi = atom_hi;
    p->nloc = atom_hi - atom_lo;
    p->npairs = (p->nloc + 1) / 2;
    p->M = M;
    p->blk = M / 32;
    p->blk_shift = 0;
    while ((1 << p->blk_shift) < p->blk) ++p->blk_shift;
    p->NB = (n_samples + p->blk - 1) / p->blk;
    p->NS = n_samples;
    p->vfull = (M - atom_size + 1) / p->blk;
    if (p->vfull > 32) p->vfull = 32;
    p->nchunks = (p->NB + p->vfull - 1) / p->vfull;
    p->bm_cap = ((2 * atom_size - 2) / p->blk + 2) * p->blk;   // positions a step window can refresh

    // GRAM table is indexed [winner atom (any of K)][owned atom][lag]
    const uint64_t gram_bytes = (uint64_t)n_atoms * p->nloc * (2ull * atom_size) * sizeof(float);
    const int ns_pad = (n_samples + 3) & ~3;   // SGRAM map rows are padded to 16 bytes (bulk-copy granularity)
    const uint64_t map_row_bytes = (uint64_t)p->nloc * ns_pad * sizeof(float);             // per signal
    const uint64_t map_bytes = (uint64_t)max_batch * map_row_bytes;
    int M2 = 512;
    while (M2 < 2 * atom_size) M2 *= 2;
    const int nvb_max = (2 * atom_size - 2) / p->blk + 2;
    // per-signal bytes of the structures every mode keeps + the resident map
    const uint64_t per_signal = map_row_bytes + (uint64_t)p->nloc * p->NB * 8 + (uint64_t)p->nloc * 8 +
                                (uint64_t)n_samples * 4;
    const uint64_t sgram_fixed = (uint64_t)n_atoms * spec_stride(M2) * 8 + (uint64_t)p->npairs * (M2 + M) * 8 +
                                 (uint64_t)2048 * M * 8 + (uint64_t)n_atoms * atom_size * 4;
    long long cap = (long long)(((double)free_b * 0.85 - (double)sgram_fixed) / (double)per_signal);
    if (mode == MPB200_MODE_SGRAM && gram_budget_bytes) {   // explicit cap on the resident map
        const long long c2 = (long long)(gram_budget_bytes / map_row_bytes);
        if (c2 < cap) cap = c2;
    }
    if (cap > max_batch) cap = max_batch;
    const bool sgram_ok = M2 <= 8192 && nvb_max <= 30 && cap >= 1;
    if (mode == MPB200_MODE_AUTO) {
        // GRAM when the table and the resident map fit AND the table build (K window transforms per pair, once per
        // dictionary: skipped when the next call brings the same dictionary) is amortised by the batch.  Measured at
        // configs[3]'s bands (1024 x 128, 16 signals): MPB_GRAM_BATCH_FACTOR, see profiles/r2_bench_c4_multiband_1gpu.json.
        const uint64_t budget = gram_budget_bytes ? gram_budget_bytes : (uint64_t)(0.4 * (double)free_b);
        const bool fits = gram_bytes <= budget && gram_bytes + map_bytes <= (uint64_t)(0.8 * (double)free_b);
        if (fits && (long long)max_batch * MPB_GRAM_BATCH_FACTOR >= n_atoms && nvb_max <= 32) mode = MPB200_MODE_GRAM;
        // SGRAM against windowed re-correlation, measured: 46 vs 63 us per atom-step at 32 x 2048 (pair, signal)
        // work items, 268 vs 341 us at 1 x 8192, 58.5 vs 62.7 us per iteration at 1 x 1024 (one rank of configs[4]
        // on 8 GPUs).  Below about a thousand items the persistent grid is mostly empty and the resident map
        // (and its longer first pass) buys nothing.
        else if (sgram_ok && (cap >= max_batch || cap >= 32) &&
                 (long long)(cap < max_batch ? cap : max_batch) * p->npairs >= 1024)
            mode = MPB200_MODE_SGRAM;
        else mode = MPB200_MODE_RECORRELATE;
    }
    p->mode = mode;
    p->GS = 2 * atom_size;
    p->M2 = M2;
    p->Bcap = max_batch;
    if (mode == MPB200_MODE_GRAM && nvb_max > 32) {
        free_plan(p);
        return fail(MPB200_EINVAL, "GRAM mode: window spans more than 32 blocks");
    }
    if (mode == MPB200_MODE_SGRAM) {
        if (!sgram_ok) {
            free_plan(p);
            return fail(MPB200_EINVAL, "SGRAM mode: atom too long for the Gram-row transform, or not even one "
                                       "signal's correlation map fits in free device memory");
        }
        const long long n_sub = (max_batch + cap - 1) / cap;         // balanced sub-batches
        p->Bcap = (int)((max_batch + n_sub - 1) / n_sub);
        p->NS = ns_pad;
        // position-free block tables pay when the refresh kernel dominates (k_apply then resolves the winner's
        // exact position with one extra dependent load); latency-bound shapes keep exact positions
        p->pos_free = MPB_DELTA_NOPOS && p->blk >= 128 && (long long)p->Bcap * p->npairs >= MPB_DELTA_NOPOS_MIN_ITEMS;
    }
    const int alloc_batch = p->Bcap;

    int rc = MPB200_OK;
    std::vector<cpx<float>> t1, t2;
    std::vector<cpx<double>> t1d, t2d;
    host_twiddles<float>(M, t1, t2);
    host_twiddles<double>(M, t1d, t2d);
    const long long total_full = (long long)alloc_batch * p->nchunks;
    p->wcap = (int)(total_full < 2048 ? total_full : 2048);
    if (p->wcap < alloc_batch) p->wcap = alloc_batch;
    std::vector<Win> wf((size_t)total_full);
    for (int b = 0; b < alloc_batch; ++b)
        for (int c = 0; c < p->nchunks; ++c) {
            Win w;
            w.row = b;
            w.blk0 = c * p->vfull;
            w.t0 = w.blk0 * p->blk;
            w.nvb = (p->NB - w.blk0 < p->vfull) ? p->NB - w.blk0 : p->vfull;
            wf[(size_t)b * p->nchunks + c] = w;
        }
#define MPB_TRY(x) do { rc = (x); if (rc) { free_plan(p); return rc; } } while (0)
    MPB_TRY(dev_alloc(p, &p->dict, (size_t)n_atoms * atom_size));
    MPB_TRY(dev_alloc(p, &p->pairspec, (size_t)p->npairs * M));
    MPB_TRY(dev_alloc(p, &p->tw1, (size_t)M));
    MPB_TRY(dev_alloc(p, &p->tw2, (size_t)256));
    MPB_TRY(dev_alloc(p, &p->tw1d, (size_t)M));
    MPB_TRY(dev_alloc(p, &p->tw2d, (size_t)256));
    MPB_TRY(dev_alloc(p, &p->winspec, (size_t)p->wcap * M));
    MPB_TRY(dev_alloc(p, &p->win_full, (size_t)total_full));
    MPB_TRY(dev_alloc(p, &p->win_step, (size_t)alloc_batch));
    MPB_TRY(dev_alloc(p, &p->bm_val, (size_t)alloc_batch * p->nloc * p->NB));
    MPB_TRY(dev_alloc(p, &p->bm_pos, (size_t)alloc_batch * p->nloc * p->NB));
    MPB_TRY(dev_alloc(p, &p->row_val, (size_t)alloc_batch * p->nloc));
    MPB_TRY(dev_alloc(p, &p->row_pos, (size_t)alloc_batch * p->nloc));
    if (p->mode == MPB200_MODE_RECORRELATE) {      // fused one-launch loop: ping-pong copy of the row tables, barrier counters
        MPB_TRY(dev_alloc(p, &p->row_val2, (size_t)alloc_batch * p->nloc));
        MPB_TRY(dev_alloc(p, &p->row_pos2, (size_t)alloc_batch * p->nloc));
        MPB_TRY(dev_alloc(p, &p->gbar, (size_t)2));
    }
    MPB_TRY(dev_alloc(p, &p->residual, (size_t)alloc_batch * n_samples));
    MPB_TRY(dev_alloc(p, &p->best, (size_t)alloc_batch));
    MPB_TRY(dev_alloc(p, &p->fp, (size_t)5));
    MPB_TRY(dev_alloc(p, &p->dict_skip, (size_t)1));
    std::vector<Win> wg;
    if (mode == MPB200_MODE_GRAM) {
        p->gram_bytes = gram_bytes;
        MPB_TRY(dev_alloc(p, &p->gram, (size_t)n_atoms * p->nloc * p->GS));
        MPB_TRY(dev_alloc(p, &p->map, (size_t)alloc_batch * p->nloc * n_samples));
        MPB_TRY(dev_alloc(p, &p->upd, (size_t)alloc_batch));
        MPB_TRY(dev_alloc(p, &p->trunc_count, (size_t)2));
        MPB_TRY(dev_alloc(p, &p->win_gram, (size_t)n_atoms));
        wg.resize((size_t)n_atoms);
        for (int k = 0; k < n_atoms; ++k) {
            Win w;
            w.row = k;
            w.t0 = -(atom_size - 1);
            w.blk0 = 0;
            w.nvb = (2 * atom_size - 1 + p->blk - 1) / p->blk;
            wg[(size_t)k] = w;
        }
    }
    std::vector<cpx<float>> t1b, t2b;
    std::vector<cpx<double>> t1bd, t2bd;
    if (mode == MPB200_MODE_SGRAM) {
        host_twiddles<float>(M2, t1b, t2b);
        host_twiddles<double>(M2, t1bd, t2bd);
        MPB_TRY(dev_alloc(p, &p->map, (size_t)alloc_batch * p->nloc * p->NS));
        MPB_TRY(dev_alloc(p, &p->upd, (size_t)alloc_batch));
        MPB_TRY(dev_alloc(p, &p->trunc_count, (size_t)2));
        MPB_TRY(dev_alloc(p, &p->pairspec2, (size_t)p->npairs * M2));
        MPB_TRY(dev_alloc(p, &p->atomspec, (size_t)n_atoms * spec_stride(M2)));
        MPB_TRY(dev_alloc(p, &p->tw1b, (size_t)M2));
        MPB_TRY(dev_alloc(p, &p->tw1bd, (size_t)M2));
        MPB_TRY(dev_alloc(p, &p->win_atoms, (size_t)n_atoms));
        wg.resize((size_t)n_atoms);
        for (int k = 0; k < n_atoms; ++k) {
            Win w;
            w.row = k;
            w.t0 = -(atom_size - 1);
            w.blk0 = 0;
            w.nvb = 0;
            wg[(size_t)k] = w;
        }
    }
#undef MPB_TRY
    cudaError_t e = cudaSuccess;
    auto up = [&](void* dst, const void* src, size_t bytes) {
        if (e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    };
    up(p->tw1, t1.data(), t1.size() * sizeof(t1[0]));
    up(p->tw2, t2.data(), t2.size() * sizeof(t2[0]));
    up(p->tw1d, t1d.data(), t1d.size() * sizeof(t1d[0]));
    up(p->tw2d, t2d.data(), t2d.size() * sizeof(t2d[0]));
    up(p->win_full, wf.data(), wf.size() * sizeof(Win));
    if (!wg.empty() && p->win_gram) up(p->win_gram, wg.data(), wg.size() * sizeof(Win));
    if (!wg.empty() && p->win_atoms) up(p->win_atoms, wg.data(), wg.size() * sizeof(Win));
    if (p->tw1b) {
        up(p->tw1b, t1b.data(), t1b.size() * sizeof(t1b[0]));
        up(p->tw1bd, t1bd.data(), t1bd.size() * sizeof(t1bd[0]));
    }
    if (e == cudaSuccess) e = cudaMemset(p->fp, 0, 5 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(p->dict_skip, 0, sizeof(int));
    if (p->gram && e == cudaSuccess) e = cudaMemset(p->gram, 0, (size_t)n_atoms * p->nloc * p->GS * sizeof(float));
    if (e != cudaSuccess) {
        free_plan(p);
        return fail(MPB200_ECUDA, std::string("table upload: ") + cudaGetErrorString(e));
    }
    {
        const int rs = reserve_host_staging(p, MPB200_DEFAULT_MAX_STEPS);
        if (rs) { free_plan(p); return rs; }
    }
    *out = reinterpret_cast<mpb200_plan_t>(p);
    return MPB200_OK;
}

int mpb200_plan_destroy(mpb200_plan_t plan) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p) return MPB200_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    free_plan(p);
    return MPB200_OK;
}

int mpb200_plan_info_get(mpb200_plan_t plan, mpb200_plan_info* info) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p || !info) return fail(MPB200_EINVAL, "null argument");
    memset(info, 0, sizeof(*info));
    info->n_atoms = p->K;
    info->atom_size = p->A;
    info->n_samples = p->N;
    info->max_batch = p->Bmax;
    info->mode = p->mode;
    info->fft_size = p->M;
    info->block = p->blk;
    info->n_blocks = p->NB;
    info->atom_lo = p->lo;
    info->atom_hi = p->hi;
    info->resident_batch = p->Bcap;
    info->fft_size2 = p->mode == MPB200_MODE_SGRAM ? p->M2 : 0;
    info->device_bytes = p->bytes;
    info->gram_bytes = p->gram_bytes;
    return MPB200_OK;
}

int mpb200_plan_timing_enable(mpb200_plan_t plan, int enable) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p) return fail(MPB200_EINVAL, "null plan");
    p->timing = enable != 0;
    p->ev_used = 0;
    p->ev_tag.clear();
    return MPB200_OK;
}

int mpb200_plan_timing_read(mpb200_plan_t plan, double* ms_by_tag, int64_t* count_by_tag) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p || !ms_by_tag || !count_by_tag) return fail(MPB200_EINVAL, "null argument");
    for (int t = 0; t < 5; ++t) { ms_by_tag[t] = 0.0; count_by_tag[t] = 0; }
    if (p->ev_used) MPB_CUDA(cudaEventSynchronize((cudaEvent_t)p->ev_pool[p->ev_used - 1]));
    for (size_t i = 1; i < p->ev_used; ++i) {
        const int tag = p->ev_tag[i];
        if (tag == 0) continue;   // a start marker opens a new call; the gap before it is not ours
        float ms = 0.f;
        MPB_CUDA(cudaEventElapsedTime(&ms, (cudaEvent_t)p->ev_pool[i - 1], (cudaEvent_t)p->ev_pool[i]));
        ms_by_tag[tag] += ms;
        count_by_tag[tag] += 1;
    }
    p->ev_used = 0;
    p->ev_tag.clear();
    return MPB200_OK;
}

static int set_dictionary_impl(mpb200_plan_t plan, const float* d, bool normalize, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, false);
    if (rc) return rc;
    if (!d) return fail(MPB200_EINVAL, "null dictionary");
    cudaStream_t st = (cudaStream_t)stream;
    // unchanged dictionary? decided on the device: the kernels below return at once when it is
    const size_t n_dict = (size_t)p->K * p->A;
    int fgrid = (int)((n_dict + 256 * 8 - 1) / (256 * 8));
    if (fgrid > p->sm_count * 8) fgrid = p->sm_count * 8;
    k_fingerprint<<<fgrid, 256, 0, st>>>(d, n_dict, p->fp);
    MPB_LAUNCH_CHECK("k_fingerprint");
    k_fingerprint_decide<<<1, 1, 0, st>>>(p->fp, p->dict_skip,
                                          (!p->dict_set || p->force_tables || p->last_normalize != (int)normalize) ? 1 : 0);
    MPB_LAUNCH_CHECK("k_fingerprint_decide");
    p->force_tables = false;
    p->last_normalize = (int)normalize;
    if (normalize) {
        k_unit_norm<<<(p->K + 7) / 8, 256, 0, st>>>(d, p->dict, p->K, p->A, 1e-8f, p->dict_skip);
        MPB_LAUNCH_CHECK("k_unit_norm");
    } else {
        MPB_CUDA(cudaMemcpyAsync(p->dict, d, (size_t)p->K * p->A * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    rc = build_pair_spectra(p, st);
    if (rc) return rc;
    if (p->mode == MPB200_MODE_GRAM) {
        rc = build_gram(p, st);
        if (rc) return rc;
    }
    if (p->mode == MPB200_MODE_SGRAM) {
        rc = build_sgram_tables(p, st);
        if (rc) return rc;
    }
    p->dict_set = true;
    p->cur_batch = 0;
    return MPB200_OK;
}

int mpb200_plan_set_option(mpb200_plan_t plan, int option, long long value) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p) return fail(MPB200_EINVAL, "null plan");
    switch (option) {
        case MPB200_OPT_REFRESH_EVERY:
            if (value < 0) return fail(MPB200_EINVAL, "refresh_every must be >= 0");
            p->refresh_every = (int)value;
            return MPB200_OK;
        case MPB200_OPT_POSITION_FREE:
            if (p->mode != MPB200_MODE_SGRAM) return fail(MPB200_EINVAL, "position-free tables exist in SGRAM mode only");
            if (value != 0 && p->blk < 128) return fail(MPB200_EINVAL, "position-free tables need blocks of >= 128 positions");
            p->pos_free = value != 0;
            p->delta_occ = p->delta_occ_db = 0;          // another kernel instantiation: re-query its occupancy
            p->cur_batch = 0;          // tables of a batch in flight were built under the other convention
            return MPB200_OK;
        case MPB200_OPT_LOCAL_CONTRAST_NORM: {
            if (value == 0) { p->lcn = false; p->cur_batch = 0; return MPB200_OK; }
            if (p->mode != MPB200_MODE_GRAM && p->mode != MPB200_MODE_SGRAM)
                return fail(MPB200_EINVAL, "local-contrast-norm selection needs a resident map (GRAM or SGRAM mode)");
            if (p->lo != 0 || p->hi != p->K)
                return fail(MPB200_EINVAL, "local-contrast-norm selection is not available on atom-sharded plans "
                                           "(the 9x9 box spans shard boundaries)");
            int rc = check_plan(p, false);
            if (rc) return rc;
            if (!p->nbm_val) {
                rc = dev_alloc(p, &p->nbm_val, (size_t)p->Bcap * p->nloc * p->NB);
                if (!rc) rc = dev_alloc(p, &p->nbm_pos, (size_t)p->Bcap * p->nloc * p->NB);
                if (!rc) rc = dev_alloc(p, &p->nrow_val, (size_t)p->Bcap * p->nloc);
                if (!rc) rc = dev_alloc(p, &p->nrow_pos, (size_t)p->Bcap * p->nloc);
                if (rc) return rc;
            }
            p->lcn = true;
            p->pos_free = false;       // the normalised tables carry exact positions
            p->delta_occ = p->delta_occ_db = 0;
            p->cur_batch = 0;
            return MPB200_OK;
        }
        case MPB200_OPT_FORCE_TABLES:
            p->force_tables = value != 0;
            return MPB200_OK;
        case MPB200_OPT_FUSED_LOOP:
            p->fused_loop = value != 0;
            return MPB200_OK;
        case MPB200_OPT_MAX_STEPS: {
            if (value < 1) return fail(MPB200_EINVAL, "max_steps must be >= 1");
            int rc = check_plan(p, false);
            if (rc) return rc;
            return reserve_host_staging(p, (size_t)value);
        }
        default:
            return fail(MPB200_EINVAL, "unknown option");
    }
}

int mpb200_plan_set_dictionary(mpb200_plan_t plan, const float* d, void* stream) {
    return set_dictionary_impl(plan, d, true, stream);
}

int mpb200_plan_set_dictionary_raw(mpb200_plan_t plan, const float* d, void* stream) {
    return set_dictionary_impl(plan, d, false, stream);
}

int mpb200_plan_get_unit_dictionary(mpb200_plan_t plan, float* out, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    MPB_CUDA(cudaMemcpyAsync(out, p->dict, (size_t)p->K * p->A * sizeof(float), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    return MPB200_OK;
}

int mpb200_begin(mpb200_plan_t plan, const float* signal, int batch, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    if (batch < 1 || batch > p->Bcap)
        return fail(MPB200_EINVAL, "batch must be in [1, resident_batch] for the step-wise interface (" +
                                       std::to_string(p->Bcap) + " signals fit at once in this mode)");
    if (!signal) return fail(MPB200_EINVAL, "null signal");
    cudaStream_t st = (cudaStream_t)stream;
    MPB_CUDA(cudaMemcpyAsync(p->residual, signal, (size_t)batch * p->N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    p->cur_batch = batch;
    p->iter = 0;
    if (p->mode == MPB200_MODE_GRAM || p->mode == MPB200_MODE_SGRAM) {
        MPB_CUDA(cudaMemsetAsync(p->trunc_count, 0, 2 * sizeof(int), st));
        rc = full_pass(p, batch, p->map, st, true);
        if (!rc && p->lcn) rc = lcn_refresh(p, batch, true, st);
        return rc;
    }
    return full_pass(p, batch, nullptr, st);
}

int mpb200_local_best(mpb200_plan_t plan, mpb200_best* best, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    if (p->cur_batch < 1) return fail(MPB200_ESTATE, "mpb200_begin has not been called");
    if (p->lcn) return fail(MPB200_ESTATE, "the step-wise interface does not offer local-contrast-norm selection");
    k_local_best<<<p->cur_batch, 256, 0, (cudaStream_t)stream>>>(p->row_val, p->row_pos, p->nloc, p->lo,
                                                                  reinterpret_cast<Best*>(best),
                                                                  p->pos_free ? p->map : nullptr, p->NS, p->N, p->blk_shift);
    MPB_LAUNCH_CHECK("k_local_best");
    return MPB200_OK;
}

int mpb200_apply(mpb200_plan_t plan, const mpb200_best* winner, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    if (p->cur_batch < 1) return fail(MPB200_ESTATE, "mpb200_begin has not been called");
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_apply<false>(p, p->cur_batch, reinterpret_cast<const Best*>(winner), 0, 1, nullptr, nullptr, nullptr,
                             p->mode != MPB200_MODE_FULL, st);
    if (rc) return rc;
    return step_refresh(p, p->cur_batch, st);
}

int mpb200_residual(mpb200_plan_t plan, float* residual_out, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    if (p->cur_batch < 1) return fail(MPB200_ESTATE, "mpb200_begin has not been called");
    MPB_CUDA(cudaMemcpyAsync(residual_out, p->residual, (size_t)p->cur_batch * p->N * sizeof(float),
                             cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MPB200_OK;
}

int mpb200_reduce_best(const mpb200_best* cand, int n_ranks, int batch, mpb200_best* winner, void* stream) {
    if (!cand || !winner || n_ranks < 1 || batch < 1) return fail(MPB200_EINVAL, "bad argument");
    k_reduce_best<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const Best*>(cand), n_ranks,
                                                                         batch, reinterpret_cast<Best*>(winner));
    MPB_LAUNCH_CHECK("k_reduce_best");
    return MPB200_OK;
}

// The iteration loop of the windowed re-correlation schedule as one cooperative launch (fused_loop.cuh): taken when
// one CTA per (pair group, signal) is resident at once.  Returns 1 when the shape does not qualify.
static int pursue_fused(Plan* p, int batch, int n_steps, int32_t* atom_out, int32_t* pos_out, float* val_out,
                        cudaStream_t st) {
    if (!p->fused_loop || p->mode != MPB200_MODE_RECORRELATE || p->lcn || n_steps < 2) return 1;
    if (p->lo != 0 || p->hi != p->K || (p->xconnected && p->xworld > 1)) return 1;
    int rc = 1;
    MPB_DISPATCH_M(p->M, {
        using F = BlockFft<MM, float>;
        constexpr int TPB = F::T < 256 ? 256 : F::T;
        constexpr int NT = TPB / F::T;
        const size_t smem = (size_t)(256 + 2 * NT * F::SMEM_CPX) * sizeof(C32);
        if (p->fused_occ == 0) {
            MPB_CUDA(allow_smem(k_pursue_fused<MM>, smem));
            int coop = 0;
            MPB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
            int occ = 0;
            MPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pursue_fused<MM>, TPB, smem));
            p->fused_occ = (coop && occ > 0) ? occ : -1;
        }
        const long long groups = (p->npairs + NT - 1) / NT;
        // + 1: a CTA per signal that owns no atom pair -- it records the events and updates the residual
        if (p->fused_occ > 0 && (groups + 1) * batch <= (long long)p->sm_count * p->fused_occ && batch <= 65535) {
            if (!p->gbar) return 1;                // (allocated at plan creation in this mode)
            MPB_CUDA(cudaMemsetAsync(p->gbar, 0, 2 * sizeof(unsigned), st));
            FusedArgs f;
            f.pairspec = p->pairspec;
            f.dict = p->dict;
            f.residual = p->residual;
            f.bm_val = p->bm_val;
            f.bm_pos = p->bm_pos;
            f.row_val = p->row_val;
            f.row_pos = p->row_pos;
            f.row_val2 = p->row_val2;
            f.row_pos2 = p->row_pos2;
            f.tw1 = p->tw1;
            f.tw2 = p->tw2;
            f.npairs = p->npairs;
            f.nloc = p->nloc;
            f.atom_lo = p->lo;
            f.n_atoms = p->K;
            f.A = p->A;
            f.N = p->N;
            f.NB = p->NB;
            f.blk_shift = p->blk_shift;
            f.n_steps = n_steps;
            f.atom_out = atom_out;
            f.pos_out = pos_out;
            f.val_out = val_out;
            f.gbar = p->gbar;
            void* kargs[] = {(void*)&f};
            MPB_CUDA(cudaLaunchCooperativeKernel((const void*)k_pursue_fused<MM>, dim3((unsigned)groups + 1, (unsigned)batch),
                                                 dim3(TPB), kargs, smem, st));
            g_launches.fetch_add(1);
            p->iter += (unsigned)(n_steps - 1);
            rc = MPB200_OK;
        }
    });
    return rc;
}

// One resident batch (<= Bcap signals) through begin + n_steps x (apply, refresh).
static int pursue_resident(Plan* p, const float* signal, int batch, int n_steps, float* residual_out,
                           int32_t* atom_out, int32_t* pos_out, float* val_out, cudaStream_t st) {
    mark(p, 0, st);
    int rc = mpb200_begin(reinterpret_cast<mpb200_plan_t>(p), signal, batch, (void*)st);
    if (rc) return rc;
    mark(p, 1, st);
    const int frc = pursue_fused(p, batch, n_steps, atom_out, pos_out, val_out, st);
    if (frc == MPB200_OK) mark(p, 3, st);
    else if (frc != 1) return frc;
    for (int s = 0; frc == 1 && s < n_steps; ++s) {
        const bool last = (s == n_steps - 1);
        rc = launch_apply<true>(p, batch, nullptr, s, n_steps, atom_out, pos_out, val_out,
                                (!last && p->mode != MPB200_MODE_FULL) ? 1 : 0, st);
        if (rc) return rc;
        mark(p, 2, st);
        if (!last) {
            rc = step_refresh(p, batch, st);
            if (rc) return rc;
            mark(p, 3, st);
        }
    }
    if (n_steps > 0) p->cur_batch = 0;  // block maxima are stale after the last subtraction
    if (residual_out)
        MPB_CUDA(cudaMemcpyAsync(residual_out, p->residual, (size_t)batch * p->N * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
    return MPB200_OK;
}

int mpb200_sparse_code(mpb200_plan_t plan, const float* signal, int batch, int n_steps, float* residual_out,
                       int32_t* atom_out, int32_t* pos_out, float* val_out, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    if (n_steps < 0) return fail(MPB200_EINVAL, "n_steps must be >= 0");
    if (n_steps > 0 && (!atom_out || !pos_out || !val_out)) return fail(MPB200_EINVAL, "null output");
    if (batch < 1 || batch > p->Bmax) return fail(MPB200_EINVAL, "batch must be in [1, max_batch]");
    if (!signal) return fail(MPB200_EINVAL, "null signal");
    const bool sharded = p->lo != 0 || p->hi != p->K;
    if (sharded && !(p->xconnected && p->xworld > 1))
        return fail(MPB200_ESTATE, "mpb200_sparse_code on an atom-sharded plan needs a connected exchange "
                                   "(mpb200_exchange_create/connect); otherwise use begin/local_best/apply");
    if (p->xconnected && p->xworld > 1 && batch > p->Bcap)
        return fail(MPB200_EINVAL, "atom-sharded pursuit: the batch must fit the resident capacity of every rank");
    cudaStream_t st = (cudaStream_t)stream;
    // signals are independent problems: the map modes walk the batch in balanced resident sub-batches
    const int n_sub = (batch + p->Bcap - 1) / p->Bcap;
    const int per = (batch + n_sub - 1) / n_sub;
    for (int b0 = 0; b0 < batch; b0 += per) {
        const int nb = batch - b0 < per ? batch - b0 : per;
        const size_t eo = (size_t)b0 * (size_t)n_steps;
        rc = pursue_resident(p, signal + (size_t)b0 * p->N, nb, n_steps,
                             residual_out ? residual_out + (size_t)b0 * p->N : nullptr,
                             atom_out ? atom_out + eo : nullptr, pos_out ? pos_out + eo : nullptr,
                             val_out ? val_out + eo : nullptr, st);
        if (rc) return rc;
    }
    return MPB200_OK;
}

// ---------------------------------------------------------------------------
// atom-sharded exchange over peer memory
// ---------------------------------------------------------------------------
int mpb200_exchange_create(mpb200_plan_t plan, int world, int rank, unsigned char* handle_out) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, false);
    if (rc) return rc;
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(MPB200_EINVAL, "bad world/rank (world <= 64)");
    if (p->mail) return fail(MPB200_ESTATE, "the plan already has a mailbox");
    const size_t slots = (size_t)2 * p->Bmax * world;
    // plain cudaMalloc (not a pool): the allocation is exported through CUDA IPC
    rc = dev_alloc(p, &p->mail, slots);
    if (!rc) rc = dev_alloc(p, &p->peer_mail, (size_t)world);
    if (!rc) rc = dev_alloc(p, &p->xerr, (size_t)1);
    if (rc) return rc;
    MPB_CUDA(cudaMemset(p->mail, 0, slots * sizeof(MailSlot)));
    MPB_CUDA(cudaMemset(p->xerr, 0, sizeof(int)));
    rc = preload_kernels(p);
    if (rc) return rc;
    p->xworld = world;
    p->xrank = rank;
    p->xseq = 0;
    if (handle_out) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
        MPB_CUDA(cudaIpcGetMemHandle(&h, p->mail));
        memcpy(handle_out, &h, 64);
    }
    return MPB200_OK;
}

static int finish_connect(Plan* p, const std::vector<MailSlot*>& ptrs) {
    MPB_CUDA(cudaMemcpy(p->peer_mail, ptrs.data(), ptrs.size() * sizeof(MailSlot*), cudaMemcpyHostToDevice));
    p->xconnected = true;
    return MPB200_OK;
}

int mpb200_exchange_connect(mpb200_plan_t plan, const unsigned char* handles) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, false);
    if (rc) return rc;
    if (!p->mail || !handles) return fail(MPB200_ESTATE, "mpb200_exchange_create has not been called");
    std::vector<MailSlot*> ptrs((size_t)p->xworld, nullptr);
    for (int r = 0; r < p->xworld; ++r) {
        if (r == p->xrank) { ptrs[(size_t)r] = p->mail; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(MPB200_ECUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " +
                                          cudaGetErrorString(e));
        }
        p->ipc_opened.push_back(q);
        ptrs[(size_t)r] = reinterpret_cast<MailSlot*>(q);
    }
    return finish_connect(p, ptrs);
}

int mpb200_exchange_mailbox(mpb200_plan_t plan, void** mailbox) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p || !mailbox) return fail(MPB200_EINVAL, "null argument");
    if (!p->mail) return fail(MPB200_ESTATE, "mpb200_exchange_create has not been called");
    *mailbox = p->mail;
    return MPB200_OK;
}

int mpb200_exchange_connect_local(mpb200_plan_t plan, void* const* mailboxes) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, false);
    if (rc) return rc;
    if (!p->mail || !mailboxes) return fail(MPB200_ESTATE, "mpb200_exchange_create has not been called");
    std::vector<MailSlot*> ptrs((size_t)p->xworld, nullptr);
    for (int r = 0; r < p->xworld; ++r) ptrs[(size_t)r] = reinterpret_cast<MailSlot*>(mailboxes[r]);
    if (ptrs[(size_t)p->xrank] != p->mail) return fail(MPB200_EINVAL, "mailboxes[rank] must be this plan's own mailbox");
    return finish_connect(p, ptrs);
}

int mpb200_exchange_status(mpb200_plan_t plan, int* timed_out) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, false);
    if (rc) return rc;
    if (!timed_out) return fail(MPB200_EINVAL, "null argument");
    *timed_out = 0;
    if (p->xerr) MPB_CUDA(cudaMemcpy(timed_out, p->xerr, sizeof(int), cudaMemcpyDeviceToHost));
    return MPB200_OK;
}

int mpb200_exchange_disconnect(mpb200_plan_t plan) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, false);
    if (rc) return rc;
    MPB_CUDA(cudaDeviceSynchronize());
    for (void* m : p->ipc_opened) cudaIpcCloseMemHandle(m);
    p->ipc_opened.clear();
    p->xconnected = false;
    return MPB200_OK;
}

int mpb200_sparse_code_host(mpb200_plan_t plan, const float* signal_host, int batch, int n_steps,
                            float* residual_out_host, int32_t* atom_out_host, int32_t* pos_out_host,
                            float* val_out_host, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    if (batch < 1 || batch > p->Bmax) return fail(MPB200_EINVAL, "batch must be in [1, max_batch]");
    if (n_steps < 0 || !signal_host) return fail(MPB200_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sig_bytes = (size_t)batch * p->N * sizeof(float);
    const size_t ev = (size_t)batch * (size_t)n_steps;
    if (!p->d_signal || (size_t)p->Bmax * (size_t)n_steps > p->ev_cap) {
        rc = reserve_host_staging(p, (size_t)(n_steps > 0 ? n_steps : 1));     // beyond the reserved iteration count
        if (rc) return rc;
    }
    MPB_CUDA(cudaMemcpyAsync(p->d_signal, signal_host, sig_bytes, cudaMemcpyHostToDevice, st));
    // the residual overwrites the staged signal (a sub-batch's input is consumed before its residual is stored)
    rc = mpb200_sparse_code(plan, p->d_signal, batch, n_steps, residual_out_host ? p->d_signal : nullptr, p->d_atom,
                            p->d_pos, p->d_val, stream);
    if (rc) return rc;
    if (n_steps > 0) {
        MPB_CUDA(cudaMemcpyAsync(atom_out_host, p->d_atom, ev * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MPB_CUDA(cudaMemcpyAsync(pos_out_host, p->d_pos, ev * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MPB_CUDA(cudaMemcpyAsync(val_out_host, p->d_val, ev * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (residual_out_host)
        MPB_CUDA(cudaMemcpyAsync(residual_out_host, p->d_signal, sig_bytes, cudaMemcpyDeviceToHost, st));
    MPB_CUDA(cudaStreamSynchronize(st));
    return MPB200_OK;
}

int mpb200_correlate(mpb200_plan_t plan, const float* signal, int batch, float* fm_out, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    int rc = check_plan(p, true);
    if (rc) return rc;
    if (batch < 1 || batch > p->Bmax) return fail(MPB200_EINVAL, "batch must be in [1, max_batch]");
    if (!signal || !fm_out) return fail(MPB200_EINVAL, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    p->cur_batch = 0;
    for (int b0 = 0; b0 < batch; b0 += p->Bcap) {
        const int nb = batch - b0 < p->Bcap ? batch - b0 : p->Bcap;
        MPB_CUDA(cudaMemcpyAsync(p->residual, signal + (size_t)b0 * p->N, (size_t)nb * p->N * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
        rc = full_pass(p, nb, fm_out + (size_t)b0 * p->nloc * p->N, st);
        if (rc) return rc;
    }
    return MPB200_OK;
}

// Scratch for k_select_dense's partial winners: a stream-ordered allocation (cudaMallocAsync / cudaFreeAsync on
// the caller's stream), so the entry point keeps no state between calls and concurrent streams never share it.
static int select_dense_impl(const float* fm, int batch, int n_atoms, int n_samples, int atom_offset,
                             mpb200_best* best, bool lcn, void* stream) {
    if (!fm || !best || batch < 1 || n_atoms < 1 || n_samples < 1) return fail(MPB200_EINVAL, "bad argument");
    int dev = 0;
    MPB_CUDA(cudaGetDevice(&dev));
    int sms = 148;
    MPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int G = (sms * 8 + batch - 1) / batch;   // CTAs per signal: fill the chip ~8 deep
    if (G > n_atoms) G = n_atoms;
    if (G < 1) G = 1;
    const size_t need = (size_t)G * batch;
    cudaStream_t st = (cudaStream_t)stream;
    Best* part = nullptr;
    keep_async_pool(dev);
    {
        cudaError_t e = cudaMallocAsync((void**)&part, need * sizeof(Best), st);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(MPB200_ENOMEM, std::string("select scratch: ") + cudaGetErrorString(e));
        }
    }
    if (lcn) k_select_dense<true><<<dim3(G, batch), 256, 0, st>>>(fm, n_atoms, n_samples, atom_offset, part);
    else k_select_dense<false><<<dim3(G, batch), 256, 0, st>>>(fm, n_atoms, n_samples, atom_offset, part);
    MPB_LAUNCH_CHECK("k_select_dense");
    k_reduce_best<<<(batch + 127) / 128, 128, 0, st>>>(part, G, batch, reinterpret_cast<Best*>(best));
    MPB_LAUNCH_CHECK("k_reduce_best");
    k_finish_best<<<(batch + 127) / 128, 128, 0, st>>>(reinterpret_cast<Best*>(best), batch, fm, n_atoms, n_samples,
                                                       atom_offset);
    MPB_LAUNCH_CHECK("k_finish_best");
    MPB_CUDA(cudaFreeAsync(part, st));
    return MPB200_OK;
}

int mpb200_select_dense(const float* fm, int batch, int n_atoms, int n_samples, int atom_offset, mpb200_best* best,
                        void* stream) {
    return select_dense_impl(fm, batch, n_atoms, n_samples, atom_offset, best, false, stream);
}

int mpb200_select_lcn(const float* fm, int batch, int n_atoms, int n_samples, int atom_offset, mpb200_best* best,
                      void* stream) {
    return select_dense_impl(fm, batch, n_atoms, n_samples, atom_offset, best, true, stream);
}

int mpb200_subtract(float* residual, int batch, int n_samples, const float* d_unit, int n_atoms, int atom_size,
                    const mpb200_best* winner, void* stream) {
    if (!residual || !d_unit || !winner || batch < 1 || n_samples < 1 || n_atoms < 1 || atom_size < 1)
        return fail(MPB200_EINVAL, "bad argument");
    k_subtract<<<batch, 256, 0, (cudaStream_t)stream>>>(residual, n_samples, d_unit, n_atoms, atom_size,
                                                        reinterpret_cast<const Best*>(winner));
    MPB_LAUNCH_CHECK("k_subtract");
    return MPB200_OK;
}

int mpb200_scatter_add(float* out, int batch, int n_samples, const float* d_unit, int n_atoms, int atom_size,
                       const int32_t* atom, const int32_t* batch_index, const int32_t* pos, const float* val,
                       const int32_t* row_offsets, int n_events, void* stream) {
    if (!out || !d_unit || batch < 1 || n_samples < 1 || n_atoms < 1 || atom_size < 1 || n_events < 0)
        return fail(MPB200_EINVAL, "bad argument");
    if (n_events == 0) return MPB200_OK;
    if (!atom || !batch_index || !pos || !val) return fail(MPB200_EINVAL, "null event array");
    dim3 grid((n_samples + 1023) / 1024, batch);
    k_scatter<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n_samples, d_unit, n_atoms, atom_size, atom, batch_index,
                                                      pos, val, row_offsets, n_events);
    MPB_LAUNCH_CHECK("k_scatter");
    return MPB200_OK;
}

int mpb200_scatter_rows(float* out, int n_rows, int n_samples, const float* rows, int atom_size,
                        const int32_t* row_index, const int32_t* pos, const int32_t* row_offsets, int n_events,
                        void* stream) {
    if (!out || n_rows < 1 || n_samples < 1 || atom_size < 1 || n_events < 0)
        return fail(MPB200_EINVAL, "bad argument");
    if (n_events == 0) return MPB200_OK;
    if (!rows || !row_index || !pos) return fail(MPB200_EINVAL, "null event array");
    dim3 grid((n_samples + 1023) / 1024, n_rows);
    k_scatter<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n_samples, rows, n_events, atom_size, nullptr, row_index,
                                                      pos, nullptr, row_offsets, n_events);
    MPB_LAUNCH_CHECK("k_scatter");
    return MPB200_OK;
}

int mpb200_gather_atoms(float* scaled, const float* d_unit, int n_atoms, int atom_size, const int32_t* atom,
                        const float* val, int n_events, void* stream) {
    if (!scaled || !d_unit || !atom || !val || n_atoms < 1 || atom_size < 1 || n_events < 0)
        return fail(MPB200_EINVAL, "bad argument");
    if (n_events == 0) return MPB200_OK;
    k_gather_atoms<<<n_events, 128, 0, (cudaStream_t)stream>>>(scaled, d_unit, n_atoms, atom_size, atom, val, n_events);
    MPB_LAUNCH_CHECK("k_gather_atoms");
    return MPB200_OK;
}

int mpb200_dictionary_update(float* running, int batch, int n_samples, float* d_unit, int n_atoms, int atom_size,
                             const int32_t* group_offsets, const int32_t* group_atom, int n_groups,
                             const int32_t* ev_batch, const int32_t* ev_pos, const float* ev_rows, int n_events,
                             void* stream) {
    if (!running || !d_unit || batch < 1 || n_samples < 1 || n_atoms < 1 || atom_size < 1 || n_groups < 0 || n_events < 0)
        return fail(MPB200_EINVAL, "bad argument");
    if (n_groups == 0 || n_events == 0) return MPB200_OK;
    if (!group_offsets || !group_atom || !ev_batch || !ev_pos || !ev_rows) return fail(MPB200_EINVAL, "null event array");
    const size_t smem = (size_t)atom_size * sizeof(float);
    if (smem > 200 * 1024) return fail(MPB200_EINVAL, "atom too long for the update kernel's shared-memory copy");
    MPB_CUDA(allow_smem(k_dictionary_update, smem));
    k_dictionary_update<<<1, 1024, smem, (cudaStream_t)stream>>>(running, n_samples, d_unit, atom_size, group_offsets,
                                                                 group_atom, n_groups, ev_batch, ev_pos, ev_rows);
    MPB_LAUNCH_CHECK("k_dictionary_update");
    return MPB200_OK;
}

int mpb200_fold_parts(const float* sub_map, int batch, int n_atoms, int n_parts, int part_len, int n_samples,
                      float* fm_out, void* stream) {
    if (!sub_map || !fm_out || batch < 1 || n_atoms < 1 || n_parts < 1 || part_len < 1 || n_samples < 1)
        return fail(MPB200_EINVAL, "bad argument");
    if (n_atoms > 65535 || batch > 65535) return fail(MPB200_EINVAL, "n_atoms and batch must be <= 65535");
    dim3 grid((n_samples + 255) / 256, n_atoms, batch);
    k_fold_parts<<<grid, 256, 0, (cudaStream_t)stream>>>(sub_map, n_atoms, n_parts, part_len, n_samples, fm_out);
    MPB_LAUNCH_CHECK("k_fold_parts");
    return MPB200_OK;
}

int mpb200_unit_norm(const float* x, float* y, int rows, int cols, float eps, void* stream) {
    if (!x || !y || rows < 1 || cols < 1) return fail(MPB200_EINVAL, "bad argument");
    k_unit_norm<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, y, rows, cols, eps);
    MPB_LAUNCH_CHECK("k_unit_norm");
    return MPB200_OK;
}

}  // extern "C"
