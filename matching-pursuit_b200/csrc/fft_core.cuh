// fft_core.cuh -- register-radix / shared-memory three-pass FFT building blocks.
//
// Everything here is __host__ __device__ so that the index arithmetic and the
// butterflies can be exercised on the CPU (tests/test_fft_core_host.py compiles
// csrc/host_emulation.cpp with g++ and walks the "threads" of a block
// sequentially, phase by phase) before any GPU minute is spent.
//
// Transform:   out[m] = sum_j in[j] * exp(DIR * 2*pi*i * j*m / M),   M = 256 * R1,
// R1 in {1,2,4,8,16,32}.  Decomposition j = j1*256 + x*16 + j3,
// m = m1 + R1*m2 + 16*R1*m3:
//   pass 1  radix-R1 over j1  (-> m1),  twiddle w_M^{c*m1},   c = x*16 + j3
//   pass 2  radix-16 over x   (-> m2),  twiddle w_256^{j3*m2}
//   pass 3  radix-16 over j3  (-> m3)
// Input and output are both in natural order; the two exchanges between the
// passes go through one shared-memory buffer with a bank-conflict-free layout
// (see BlockFft::addr).
#pragma once
#include <cstdint>
#include <type_traits>

#if defined(__CUDACC__)
#define MPB_HD __host__ __device__ __forceinline__
#else
#define MPB_HD inline
#endif

namespace mpb {

template <typename T>
struct cpx {
    T x, y;
};

template <typename T> MPB_HD cpx<T> operator+(cpx<T> a, cpx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> MPB_HD cpx<T> operator-(cpx<T> a, cpx<T> b) { return {a.x - b.x, a.y - b.y}; }
#ifndef MPB_F32X2
#define MPB_F32X2 1   // complex add/sub as Blackwell packed fp32x2 instructions (FADD2 / FFMA2: one issue slot for both parts; k_delta -1.8%, k_corr -4.8%)
#endif
#if MPB_F32X2 && defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
__device__ __forceinline__ cpx<float> operator+(cpx<float> a, cpx<float> b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return {r.x, r.y};
}
__device__ __forceinline__ cpx<float> operator-(cpx<float> a, cpx<float> b) {
    const float2 r = __ffma2_rn(make_float2(b.x, b.y), make_float2(-1.f, -1.f), make_float2(a.x, a.y));
    return {r.x, r.y};
}
#endif
template <typename T> MPB_HD cpx<T> cmul(cpx<T> a, cpx<T> b) {
    return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
#ifndef MPB_CMUL2
#define MPB_CMUL2 1   // complex multiply as TWO packed instructions: FMUL2 a.F32x2 * b.x ; FFMA2 -a.F32x2.LO_HI.NP * b.y + t
                      // (the swap and the per-half negation of the first operand are operand modifiers of FFMA2, the scalar
                      // factors are broadcast operands or immediates)
#endif
#if MPB_F32X2 && MPB_CMUL2 && defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
__device__ __forceinline__ cpx<float> cmul(cpx<float> a, cpx<float> b) {
    const float2 t = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.x));
    const float2 r = __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), t);
    return {r.x, r.y};
}
#define MPB_CMUL2_ACTIVE 1
#else
#define MPB_CMUL2_ACTIVE 0
#endif
template <typename T> MPB_HD cpx<T> cconj(cpx<T> a) { return {a.x, -a.y}; }

template <int B, int E, typename F>
MPB_HD void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// cos(2*pi*m/32), m = 0..8
template <int M32>
struct Cos32 {
    static constexpr double value =
        M32 == 0 ? 1.0 :
        M32 == 1 ? 0.9807852804032304491261822 :
        M32 == 2 ? 0.9238795325112867561281832 :
        M32 == 3 ? 0.8314696123025452370787884 :
        M32 == 4 ? 0.7071067811865475244008444 :
        M32 == 5 ? 0.5555702330196022247428308 :
        M32 == 6 ? 0.3826834323650897717284600 :
        M32 == 7 ? 0.1950903220161282678482849 : 0.0;
};

// v * exp(DIR * 2*pi*i * NUM / DEN) with NUM, DEN compile-time; DEN | 32.
template <int DIR, int DEN, int NUM, typename T>
MPB_HD cpx<T> tw_const(cpx<T> v) {
    constexpr int q = (((NUM % DEN) + DEN) % DEN) * (32 / DEN);  // angle in 32nds of a turn, 0..31
    if constexpr (q == 0) {
        return v;
    } else if constexpr (q == 16) {
        return {-v.x, -v.y};
    } else if constexpr (q == 8) {           // * (DIR * i)
        if constexpr (DIR > 0) return {-v.y, v.x};
        else return {v.y, -v.x};
    } else if constexpr (q == 24) {          // * (-DIR * i)
        if constexpr (DIR > 0) return {v.y, -v.x};
        else return {-v.y, v.x};
    } else if constexpr (MPB_CMUL2_ACTIVE && std::is_same<T, float>::value) {
        // the whole rotation as one two-instruction complex multiply by the constant exp(DIR * 2 pi i q / 32)
        constexpr int quad = q / 8, r = q % 8;
        constexpr T c = (T)Cos32<r>::value, s = (T)Cos32<8 - r>::value;
        constexpr T wr = quad == 0 ? c : quad == 1 ? -s : quad == 2 ? -c : s;
        constexpr T wi0 = quad == 0 ? s : quad == 1 ? c : quad == 2 ? -s : -c;
        constexpr T wi = DIR > 0 ? wi0 : -wi0;
        return cmul(v, cpx<T>{wr, wi});
    } else {
        // reduce to first quadrant: angle = quad*8 + r, r in 1..7
        constexpr int quad = q / 8, r = q % 8;
        constexpr T c = (T)Cos32<r>::value;
        constexpr T s = (T)Cos32<8 - r>::value;   // sin(2 pi r/32) = cos(2 pi (8-r)/32)
        constexpr T ss = DIR > 0 ? s : -s;
        cpx<T> t;
        if constexpr (r == 4) {
            // 45 degrees: c == s, two multiplies instead of four
            constexpr T h = (T)Cos32<4>::value;
            if constexpr (DIR > 0) t = {(v.x - v.y) * h, (v.x + v.y) * h};
            else t = {(v.x + v.y) * h, (v.y - v.x) * h};
        } else {
            t = {v.x * c - v.y * ss, v.x * ss + v.y * c};
        }
        // multiply by (DIR*i)^quad
        if constexpr (quad == 0) return t;
        else if constexpr (quad == 2) return {-t.x, -t.y};
        else if constexpr ((quad == 1) == (DIR > 0)) return {-t.y, t.x};
        else return {t.y, -t.x};
    }
}

// In-register DFT of R points, natural order in and out:
//   v[k] <- sum_n v[n] * exp(DIR * 2*pi*i * n*k / R)
template <int R, int DIR, typename T>
struct Dft {
    static MPB_HD void run(cpx<T>* v) {
        static_assert(R == 8 || R == 16 || R == 32, "radix");
        constexpr int R1 = 4, R2 = R / 4;
        cpx<T> tmp[R];
        static_for<0, R2>([&](auto n2c) {
            constexpr int n2 = decltype(n2c)::value;
            cpx<T> col[R1];
            static_for<0, R1>([&](auto n1c) {
                constexpr int n1 = decltype(n1c)::value;
                col[n1] = v[n1 * R2 + n2];
            });
            Dft<R1, DIR, T>::run(col);
            static_for<0, R1>([&](auto k1c) {
                constexpr int k1 = decltype(k1c)::value;
                tmp[k1 * R2 + n2] = tw_const<DIR, R, n2 * k1, T>(col[k1]);
            });
        });
        static_for<0, R1>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            cpx<T> row[R2];
            static_for<0, R2>([&](auto n2c) {
                constexpr int n2 = decltype(n2c)::value;
                row[n2] = tmp[k1 * R2 + n2];
            });
            Dft<R2, DIR, T>::run(row);
            static_for<0, R2>([&](auto k2c) {
                constexpr int k2 = decltype(k2c)::value;
                v[k1 + R1 * k2] = row[k2];
            });
        });
    }
};
template <int DIR, typename T>
struct Dft<1, DIR, T> {
    static MPB_HD void run(cpx<T>*) {}
};
template <int DIR, typename T>
struct Dft<2, DIR, T> {
    static MPB_HD void run(cpx<T>* v) {
        cpx<T> a = v[0], b = v[1];
        v[0] = a + b;
        v[1] = a - b;
    }
};
template <int DIR, typename T>
struct Dft<4, DIR, T> {
    static MPB_HD void run(cpx<T>* v) {
        cpx<T> t0 = v[0] + v[2], t1 = v[0] - v[2], t2 = v[1] + v[3];
        cpx<T> t3 = tw_const<DIR, 4, 1, T>(v[1] - v[3]);
        v[0] = t0 + t2;
        v[2] = t0 - t2;
        v[1] = t1 + t3;
        v[3] = t1 - t3;
    }
};

// k = hi + (k - hi) with hi the largest power of two < k (k/2 when k is a power of two): both parts are < k.
template <int K, int P = 1, bool DONE = (2 * P >= K)>
struct SplitPoint {
    static constexpr int value = SplitPoint<K, 2 * P>::value;
};
template <int K, int P>
struct SplitPoint<K, P, true> {
    static constexpr int value = P;
};

// Three-pass block FFT of M = 256*R1 points carried by T = M/E threads, each
// holding E = max(R1,16) complex values in registers.
//
// Shared buffer layout between the passes, logical index (m1, x, j3) with
// m1 < R1, x < 16, j3 < 16:
//      addr = x*XS + j3*S + m1,   S = R1|1 (odd),  XS = 16*S + (R1 < 16 ? R1 : 0)
//  * pass-1 store / pass-2 load+store: one instruction has fixed (m1 or x), the
//    16 lanes of a half-warp walk j3 -> stride S (odd) -> 16 distinct 8-byte banks
//  * pass-3 load: fixed j3, lanes walk mu = m1 + R1*m2 -> consecutive addresses
//    inside one x, and the XS padding keeps consecutive x on distinct banks.
// E_ = 0: E = max(R1, 16); E_ = 32 halves the threads per transform (two butterflies per thread and pass).
template <int M_, typename Real, int E_ = 0>
struct BlockFft {
    using C = cpx<Real>;
    static constexpr int M = M_;
    static constexpr int R1 = M / 256;
    static_assert(R1 == 1 || R1 == 2 || R1 == 4 || R1 == 8 || R1 == 16 || R1 == 32, "M must be 256..8192, power of 2");
    static constexpr int E = E_ ? E_ : (R1 > 16 ? R1 : 16);     // complex values per thread
    static_assert(E % 16 == 0 && E % R1 == 0, "E must be a multiple of 16 and of R1");
    static constexpr int T = M / E;                 // threads per transform
    static constexpr int NB1 = E / R1;              // pass-1 columns per thread
    static constexpr int NB2 = E / 16;              // pass-2 / pass-3 butterflies per thread
    static constexpr int S = R1 | 1;
    static constexpr int XS = 16 * S + (R1 < 16 ? R1 : 0);
    static constexpr int SMEM_CPX = 16 * XS;        // >= M
    static constexpr int TW1 = M;                   // tw1[m1*256 + c] = exp(+2 pi i c m1 / M)
    static constexpr int TW2 = 256;                 // tw2[m2*16 + j3] = exp(+2 pi i j3 m2 / 256)

    static MPB_HD int addr(int m1, int x, int j3) { return x * XS + j3 * S + m1; }

    // Which input element lives in register slot e of thread tl before pass 1.
    static MPB_HD int in_index(int tl, int e) {
        int u = e / R1, j1 = e % R1;
        return j1 * 256 + (tl + T * u);
    }
    // Which output element lives in register slot e of thread tl after pass 3.
    static MPB_HD int out_index(int tl, int e) {
        int u = e / 16, m3 = e % 16;
        return (tl + T * u) + (M / 16) * m3;
    }

    // Where pass 1 stores register slot e of thread tl.  Every thread owns its NB1 runs of R1 consecutive
    // entries, so a spectrum that is STAGED in this layout (input bin in_index(tl, e) kept at slot_addr(tl, e))
    // can be read and then overwritten by pass 1 without a barrier in between.
    static MPB_HD int slot_addr(int tl, int e) {
        const int u = e / R1, j1 = e % R1, c = tl + T * u;
        return addr(j1, c >> 4, c & 15);
    }
    // The staged address of input bin j = j1*256 + c.
    static MPB_HD int bin_addr(int j) {
        const int j1 = j >> 8, c = j & 255;
        return addr(j1, c >> 4, c & 15);
    }

    // r[u*R1 + j1] holds in[j1*256 + c], c = tl + T*u.
    template <int DIR>
    static MPB_HD void pass1(C* r, int tl, C* sm, const C* __restrict__ tw1) {
        static_for<0, NB1>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const int c = tl + T * u;
            Dft<R1, DIR, Real>::run(r + u * R1);
            const int x = c >> 4, j3 = c & 15;
            static_for<0, R1>([&](auto m1c) {
                constexpr int m1 = decltype(m1c)::value;
                C v = r[u * R1 + m1];
                if constexpr (m1 > 0) {
                    C w = tw1[m1 * 256 + c];
                    if constexpr (DIR < 0) w.y = -w.y;
                    v = cmul(v, w);
                }
                sm[addr(m1, x, j3)] = v;
            });
        });
    }

    // z^1 .. z^(P-1) from z: every power is the product of two earlier ones (z^(2^k + j) = z^(2^k) * z^j), so
    // P-2 complex multiplications, at most log2(P) deep -- the rounding error stays at a few ulp.
    template <int P>
    static MPB_HD void powers(C z, C* zp) {
        zp[1] = z;
        static_for<2, P>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr int hi = SplitPoint<k>::value;
            zp[k] = cmul(zp[hi], zp[k - hi]);
        });
    }

    // pass 1 with the twiddles w^(c*m1) generated from the one table entry w^c: trades R1-1 table loads per
    // column (L1/L2 traffic on the busiest unit of k_delta) for R1-2 complex multiplications.
    template <int DIR>
    static MPB_HD void pass1_gen(C* r, int tl, C* sm, const C* __restrict__ tw1) {
        static_assert(R1 >= 4, "generated twiddles need a radix >= 4 first pass");
        static_for<0, NB1>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const int c = tl + T * u;
            Dft<R1, DIR, Real>::run(r + u * R1);
            const int x = c >> 4, j3 = c & 15;
            C z = tw1[256 + c];
            if constexpr (DIR < 0) z.y = -z.y;
            C zp[R1];
            powers<R1>(z, zp);
            sm[addr(0, x, j3)] = r[u * R1];
            static_for<1, R1>([&](auto m1c) {
                constexpr int m1 = decltype(m1c)::value;
                sm[addr(m1, x, j3)] = cmul(r[u * R1 + m1], zp[m1]);
            });
        });
    }

    template <int DIR>
    static MPB_HD void pass2(C* r, int tl, C* sm, const C* __restrict__ tw2) {
        static_for<0, NB2>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const int beta = tl + T * u;
            const int j3 = beta & 15, m1 = beta >> 4;
            static_for<0, 16>([&](auto xc) {
                constexpr int x = decltype(xc)::value;
                r[u * 16 + x] = sm[addr(m1, x, j3)];
            });
            Dft<16, DIR, Real>::run(r + u * 16);
            static_for<0, 16>([&](auto m2c) {
                constexpr int m2 = decltype(m2c)::value;
                C v = r[u * 16 + m2];
                if constexpr (m2 > 0) {
                    C w = tw2[m2 * 16 + j3];
                    if constexpr (DIR < 0) w.y = -w.y;
                    v = cmul(v, w);
                }
                sm[addr(m1, m2, j3)] = v;
            });
        });
    }

    // ------------------------------------------------------------------------------------------------------------
    // "Local first exchange" form (M = 4096, 256 threads x 16 values): the inputs are dealt to the threads so that the
    // exchange between pass 1 and pass 2 stays inside a HALF-WARP -- 16 consecutive lanes trade 16 x 16 values through
    // their own 272-entry region of the buffer and need a __syncwarp(), not a CTA barrier -- which leaves ONE CTA
    // barrier per transform (between pass 2 and pass 3) instead of two.
    //   thread tl:  x = tl & 15, j3 = tl >> 4  holds inputs  j1*256 + c,  c = x*16 + j3   (the caller's tables are
    //               stored in that order: in_index_local; a spectrum staged in the buffer sits at slot_addr_local)
    //   buffer:     P(a, b, j3) = j3*XS + a*S + b      a = m1,  b = x (before pass 2) or m2 (after)
    //   pass 1      stores P(m1, x, j3), m1 = 0..15            (lanes of a half-warp: consecutive b)
    //   pass 2      thread tl: m1 = tl & 15, j3 = tl >> 4; loads P(m1, x, j3), x = 0..15 (same half-warp's region),
    //               stores its outputs IN PLACE: P(m1, m2, j3)  (lanes: stride S, odd -> conflict free)
    //   pass 3      thread tl: m1 = tl & 15, m2 = tl >> 4; loads P(m1, m2, j3), j3 = 0..15; outputs as in pass3:
    //               r[m3] = out[tl + (M/16)*m3]
    // ------------------------------------------------------------------------------------------------------------
    static constexpr bool HAS_LOCAL = (R1 == 16 && E == 16);
    static MPB_HD int P(int a, int b, int j3) { return j3 * XS + a * S + b; }
    static MPB_HD int in_index_local(int tl, int e) { return e * 256 + ((tl & 15) * 16 + (tl >> 4)); }
    static MPB_HD int slot_addr_local(int tl, int e) { return P(e, tl & 15, tl >> 4); }
    // where the local form keeps input bin j = j1*256 + c of a staged spectrum / the position of bin j in a table
    // stored in the local form's load order (thread-major: entry j1*256 + tl is thread tl's j1-th input)
    static MPB_HD int bin_addr_local(int j) { const int c = j & 255; return P(j >> 8, c >> 4, c & 15); }
    static MPB_HD int table_index_local(int j) { const int c = j & 255; return (j & ~255) | ((c & 15) << 4) | (c >> 4); }

    template <int DIR>
    static MPB_HD void pass1_local(C* r, int tl, C* sm, const C* __restrict__ tw1) {
        static_assert(HAS_LOCAL, "the local-exchange form exists for M = 4096 only");
        const int x = tl & 15, j3 = tl >> 4, c = x * 16 + j3;
        Dft<16, DIR, Real>::run(r);
        C z = tw1[256 + c];
        if constexpr (DIR < 0) z.y = -z.y;
        C zp[16];
        powers<16>(z, zp);
        sm[P(0, x, j3)] = r[0];
        static_for<1, 16>([&](auto m1c) {
            constexpr int m1 = decltype(m1c)::value;
            sm[P(m1, x, j3)] = cmul(r[m1], zp[m1]);
        });
    }

    template <int DIR>
    static MPB_HD void pass2_local(C* r, int tl, C* sm, const C* __restrict__ tw2) {
        const int m1 = tl & 15, j3 = tl >> 4;
        static_for<0, 16>([&](auto xc) {
            constexpr int x = decltype(xc)::value;
            r[x] = sm[P(m1, x, j3)];
        });
        Dft<16, DIR, Real>::run(r);
        static_for<0, 16>([&](auto m2c) {
            constexpr int m2 = decltype(m2c)::value;
            C v = r[m2];
            if constexpr (m2 > 0) {
                C w = tw2[m2 * 16 + j3];
                if constexpr (DIR < 0) w.y = -w.y;
                v = cmul(v, w);
            }
            sm[P(m1, m2, j3)] = v;
        });
    }

    template <int DIR>
    static MPB_HD void pass3_local(C* r, int tl, const C* sm) {
        const int m1 = tl & 15, m2 = tl >> 4;
        static_for<0, 16>([&](auto jc) {
            constexpr int j3 = decltype(jc)::value;
            r[j3] = sm[P(m1, m2, j3)];
        });
        Dft<16, DIR, Real>::run(r);
    }

    // afterwards r[u*16 + m3] = out[(tl + T*u) + (M/16)*m3]
    template <int DIR>
    static MPB_HD void pass3(C* r, int tl, const C* sm) {
        static_for<0, NB2>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            const int mu = tl + T * u;
            const int m1 = mu % R1, m2 = mu / R1;
            static_for<0, 16>([&](auto jc) {
                constexpr int j3 = decltype(jc)::value;
                r[u * 16 + j3] = sm[addr(m1, m2, j3)];
            });
            Dft<16, DIR, Real>::run(r + u * 16);
        });
    }
};

}  // namespace mpb
