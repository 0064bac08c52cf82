// host_emulation.cpp -- CPU walk-through of the block FFT (test tool, not product).
//
// Compiled with g++ by tests/test_fft_core_host.py.  Each "thread" of a block is
// visited in turn, one pass at a time (a pass boundary is where the kernel has
// a __syncthreads), so the register/shared-memory index arithmetic of
// fft_core.cuh is exercised exactly as on the device.
#include <cmath>
#include <vector>
#include "fft_core.cuh"

using namespace mpb;

template <int M, typename Real>
static void run_fft(int dir, const double* in_re, const double* in_im, double* out_re, double* out_im) {
    using F = BlockFft<M, Real>;
    using C = cpx<Real>;
    std::vector<C> tw1(F::TW1), tw2(F::TW2), sm(F::SMEM_CPX);
    const double two_pi = 6.283185307179586476925286766559;
    for (int m1 = 0; m1 < F::R1; ++m1)
        for (int c = 0; c < 256; ++c) {
            double a = two_pi * (double)((long long)c * m1 % M) / M;
            tw1[m1 * 256 + c] = {(Real)std::cos(a), (Real)std::sin(a)};
        }
    for (int m2 = 0; m2 < 16; ++m2)
        for (int j3 = 0; j3 < 16; ++j3) {
            double a = two_pi * (double)(j3 * m2 % 256) / 256;
            tw2[m2 * 16 + j3] = {(Real)std::cos(a), (Real)std::sin(a)};
        }
    std::vector<C> regs((size_t)F::T * F::E);
    for (int tl = 0; tl < F::T; ++tl)
        for (int e = 0; e < F::E; ++e) {
            int j = F::in_index(tl, e);
            regs[(size_t)tl * F::E + e] = {(Real)in_re[j], (Real)in_im[j]};
        }
    for (int tl = 0; tl < F::T; ++tl) {
        if (dir > 0) F::template pass1<1>(&regs[(size_t)tl * F::E], tl, sm.data(), tw1.data());
        else F::template pass1<-1>(&regs[(size_t)tl * F::E], tl, sm.data(), tw1.data());
    }
    for (int tl = 0; tl < F::T; ++tl) {
        if (dir > 0) F::template pass2<1>(&regs[(size_t)tl * F::E], tl, sm.data(), tw2.data());
        else F::template pass2<-1>(&regs[(size_t)tl * F::E], tl, sm.data(), tw2.data());
    }
    for (int tl = 0; tl < F::T; ++tl) {
        if (dir > 0) F::template pass3<1>(&regs[(size_t)tl * F::E], tl, sm.data());
        else F::template pass3<-1>(&regs[(size_t)tl * F::E], tl, sm.data());
    }
    for (int tl = 0; tl < F::T; ++tl)
        for (int e = 0; e < F::E; ++e) {
            int m = F::out_index(tl, e);
            out_re[m] = regs[(size_t)tl * F::E + e].x;
            out_im[m] = regs[(size_t)tl * F::E + e].y;
        }
}

template <typename Real>
static int dispatch(int M, int dir, const double* ir, const double* ii, double* orr, double* oi) {
    switch (M) {
        case 256: run_fft<256, Real>(dir, ir, ii, orr, oi); return 0;
        case 512: run_fft<512, Real>(dir, ir, ii, orr, oi); return 0;
        case 1024: run_fft<1024, Real>(dir, ir, ii, orr, oi); return 0;
        case 2048: run_fft<2048, Real>(dir, ir, ii, orr, oi); return 0;
        case 4096: run_fft<4096, Real>(dir, ir, ii, orr, oi); return 0;
        case 8192: run_fft<8192, Real>(dir, ir, ii, orr, oi); return 0;
    }
    return 1;
}

extern "C" int emu_fft(int M, int dir, int use_double, const double* in_re, const double* in_im,
                       double* out_re, double* out_im) {
    return use_double ? dispatch<double>(M, dir, in_re, in_im, out_re, out_im)
                      : dispatch<float>(M, dir, in_re, in_im, out_re, out_im);
}

// bank-conflict census of the shared-memory layout: worst number of distinct
// addresses that share one 8-byte bank within a half-warp, per pass.
extern "C" int emu_max_conflict(int M, int pass) {
    auto census = [&](auto tag) {
        using F = decltype(tag);
        int worst = 1;
        for (int half = 0; half < F::T / 16; ++half) {
            for (int u = 0; u < (pass == 1 ? F::NB1 : F::NB2); ++u)
                for (int k = 0; k < (pass == 1 ? F::R1 : 16); ++k) {
                    int count[16] = {0};
                    for (int lane = 0; lane < 16; ++lane) {
                        int tl = half * 16 + lane, a;
                        int beta = tl + F::T * u;
                        if (pass == 1) a = F::addr(k, beta >> 4, beta & 15);
                        else if (pass == 2) a = F::addr(beta >> 4, k, beta & 15);
                        else a = F::addr(beta % F::R1, beta / F::R1, k);
                        count[a & 15]++;
                    }
                    for (int b = 0; b < 16; ++b) worst = count[b] > worst ? count[b] : worst;
                }
        }
        return worst;
    };
    switch (M) {
        case 256: return census(BlockFft<256, float>{});
        case 512: return census(BlockFft<512, float>{});
        case 1024: return census(BlockFft<1024, float>{});
        case 2048: return census(BlockFft<2048, float>{});
        case 4096: return census(BlockFft<4096, float>{});
        case 8192: return census(BlockFft<8192, float>{});
    }
    return -1;
}
