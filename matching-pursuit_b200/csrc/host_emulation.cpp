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

// The local-first-exchange form of the 4096-point transform (BlockFft::pass*_local): pass 1 and pass 2 of a
// half-warp are run back to back before the next half-warp is touched -- on the device only a __syncwarp() separates
// them, so any value that crossed a half-warp would be missing here -- then one "CTA barrier", then pass 3.
// `staged` != 0: the input is first laid out as a staged spectrum (bin_addr_local) and read from the buffer slots
// that pass 1 overwrites, the way k_delta consumes the winner spectrum.
extern "C" int emu_fft_local(int dir, int staged, const double* in_re, const double* in_im, double* out_re,
                             double* out_im, double* table_re, double* table_im) {
    constexpr int M = 4096;
    using F = BlockFft<M, float>;
    using C = cpx<float>;
    std::vector<C> tw1(F::TW1), tw2(F::TW2), sm(F::SMEM_CPX, C{0.f, 0.f});
    const double two_pi = 6.283185307179586476925286766559;
    for (int m1 = 0; m1 < F::R1; ++m1)
        for (int c = 0; c < 256; ++c) {
            double a = two_pi * (double)((long long)c * m1 % M) / M;
            tw1[m1 * 256 + c] = {(float)std::cos(a), (float)std::sin(a)};
        }
    for (int m2 = 0; m2 < 16; ++m2)
        for (int j3 = 0; j3 < 16; ++j3) {
            double a = two_pi * (double)(j3 * m2 % 256) / 256;
            tw2[m2 * 16 + j3] = {(float)std::cos(a), (float)std::sin(a)};
        }
    // a table stored in the local form's load order: entry j1*256 + tl is thread tl's j1-th input
    for (int j = 0; j < M; ++j) {
        table_re[F::table_index_local(j)] = in_re[j];
        table_im[F::table_index_local(j)] = in_im[j];
    }
    if (staged)
        for (int j = 0; j < M; ++j) sm[F::bin_addr_local(j)] = {(float)in_re[j], (float)in_im[j]};
    std::vector<C> regs((size_t)F::T * F::E);
    for (int half = 0; half < F::T / 16; ++half) {
        for (int tl = half * 16; tl < half * 16 + 16; ++tl) {
            C* r = &regs[(size_t)tl * F::E];
            for (int e = 0; e < F::E; ++e) {
                if (staged) r[e] = sm[F::slot_addr_local(tl, e)];
                else r[e] = {(float)table_re[e * 256 + tl], (float)table_im[e * 256 + tl]};
                if (F::in_index_local(tl, e) != F::table_index_local(F::table_index_local(F::in_index_local(tl, e)))) return 2;
            }
        }
        for (int tl = half * 16; tl < half * 16 + 16; ++tl) {
            C* r = &regs[(size_t)tl * F::E];
            if (dir > 0) F::pass1_local<1>(r, tl, sm.data(), tw1.data());
            else F::pass1_local<-1>(r, tl, sm.data(), tw1.data());
        }
        for (int tl = half * 16; tl < half * 16 + 16; ++tl) {
            C* r = &regs[(size_t)tl * F::E];
            if (dir > 0) F::pass2_local<1>(r, tl, sm.data(), tw2.data());
            else F::pass2_local<-1>(r, tl, sm.data(), tw2.data());
        }
    }
    for (int tl = 0; tl < F::T; ++tl) {
        C* r = &regs[(size_t)tl * F::E];
        if (dir > 0) F::pass3_local<1>(r, tl, sm.data());
        else F::pass3_local<-1>(r, tl, sm.data());
    }
    for (int tl = 0; tl < F::T; ++tl)
        for (int e = 0; e < F::E; ++e) {
            int m = F::out_index(tl, e);
            out_re[m] = regs[(size_t)tl * F::E + e].x;
            out_im[m] = regs[(size_t)tl * F::E + e].y;
        }
    return 0;
}

// Worst number of 8-byte words of a half-warp that share a bank, over every shared-memory instruction of the local
// form (pass-1 stores, pass-2 loads and stores, pass-3 loads, staged-spectrum loads).
extern "C" int emu_max_conflict_local() {
    using F = BlockFft<4096, float>;
    int worst = 1;
    for (int half = 0; half < 16; ++half)
        for (int k = 0; k < 16; ++k)
            for (int which = 0; which < 3; ++which) {
                int count[16] = {0};
                for (int lane = 0; lane < 16; ++lane) {
                    const int tl = half * 16 + lane;
                    int a;
                    if (which == 0) a = F::P(k, tl & 15, tl >> 4);          // pass-1 store / staged load: fixed m1
                    else if (which == 1) a = F::P(tl & 15, k, tl >> 4);     // pass-2 load (fixed x) and store (fixed m2)
                    else a = F::P(tl & 15, tl >> 4, k);                     // pass-3 load: fixed j3
                    count[a & 15]++;
                }
                for (int b = 0; b < 16; ++b) worst = count[b] > worst ? count[b] : worst;
            }
    return worst;
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
