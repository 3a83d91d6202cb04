// Warp-level FP64 complex FFT of length N = 8 * 8 * R3 (R3 = 20 -> N = 1280), sign +i.
//
// One warp owns one transform: each lane keeps 40 complex points in registers and
// runs three in-register radix passes (8, 8, R3) separated by two transposes through
// a warp-private shared-memory buffer (real and imaginary parts in two rounds, so the
// buffer is only ~N doubles).  No block-level synchronisation is involved: warps of a
// CTA run independent transforms and only ever __syncwarp().
//
// Index maps (validated against numpy in tools/fft_model.py, which mirrors this file):
//   n = n1*(N/8) + n2*R3 + n3,   k = k1 + 8*k2 + 64*k3
//   load   : v[j*8+n1]   = x[n1*(N/8) + t + 32*j]                 (j<5, lane t)
//   pass 1 : radix-8 over n1, times w_N^{(t+32j) k1}
//   xchg 1 : lane t takes pairs p = t+32*j' -> (k1,n3) = (p%8, p/8), all n2
//   pass 2 : radix-8 over n2, times w_{N/8}^{n3 k2}
//   xchg 2 : lane t takes q = t+32*u -> (k1,k2) = (q%8, q/8), all n3
// Shared-memory layouts (8-byte words, 16 bank pairs; a 64-bit warp access is two half-warp
// wavefronts when the 16 lanes of each half hit 16 distinct bank pairs):
//   xchg 1 : word k1*S1 + n2*R3 + n3 with S1 = N/8 + 2 = 2 (mod 16): stores are lane-
//            contiguous, loads see k1 = all eight values and two adjacent n3 per half-warp
//            -> 2 k1 + n3 distinct
//   xchg 2 : word 64*n3 + (q xor 8*(n3 & 1)), q = k1 + 8 k2: compact, and both the stores
//            (k1 x two adjacent n3) and the loads (k1 x two adjacent k2) are conflict-free
//   pass 3 : radix-R3 over n3 (Good-Thomas 5 x R3/5, no inner twiddles)
//   result : v[u*R3+k3]  = X[t + 32*u + 64*k3]
//
// Every function is __host__ __device__ so that tools/host_check.cu can run the very
// same code lane by lane on the CPU.
//
// The transform is templated on the complex type Z: double2 (the product path), float2, or
// Z2 = a complex number whose components are PACKED PAIRS of floats (F2: two independent
// transforms, i.e. two wavelengths of the same row pair, in the f32x2 instructions of sm_100 -
// FADD2 / FMUL2 / FFMA2 - at half the instruction count of two float2 transforms).  The stage-B
// row kernel uses Z2 for row pairs whose every entry is below exp(-25) of the OTF peak
// (psfr_hot.cu).  One exchange word is 8 bytes in every instantiation (a double, a float2, or
// the F2 of one component), so all of them move through the same conflict-free layouts:
// double2 and Z2 in two rounds (real parts, imaginary parts), float2 in one.  Single-precision
// twiddles come from a float2 copy of the tables (a double -> float conversion costs as much
// as four FMAs).
#pragma once
#include <cuda_runtime.h>

#define PSFR_HD __host__ __device__ __forceinline__

namespace psfr {

template <class Z>
struct ZTraits;
template <>
struct ZTraits<double2> {
    using S = double;   // scalar
    using W = double;   // shared-memory exchange word: one component per round
    static constexpr int Rounds = 2;
};
template <>
struct ZTraits<float2> {
    using S = float;
    using W = float2;   // both components in one 8-byte word, one round
    static constexpr int Rounds = 1;
};
// two independent single-precision values in one 64-bit register pair
struct F2 {
    float2 v;
    PSFR_HD F2() {}
    PSFR_HD F2(float a, float b) {
        v.x = a;
        v.y = b;
    }
    PSFR_HD explicit F2(double c) { v.x = v.y = (float)c; }   // constants, same in both halves
};
PSFR_HD F2 operator+(F2 a, F2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    F2 r;
    r.v = __fadd2_rn(a.v, b.v);
    return r;
#else
    return F2(a.v.x + b.v.x, a.v.y + b.v.y);
#endif
}
PSFR_HD F2 operator*(F2 a, F2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    F2 r;
    r.v = __fmul2_rn(a.v, b.v);
    return r;
#else
    return F2(a.v.x * b.v.x, a.v.y * b.v.y);
#endif
}
// a * b + c in one FFMA2
PSFR_HD F2 sfma(F2 a, F2 b, F2 c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    F2 r;
    r.v = __ffma2_rn(a.v, b.v, c.v);
    return r;
#else
    return F2(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y));
#endif
}
// a - b = b * (-1) + a, exact, one FFMA2 (the packed add has no negated operand)
PSFR_HD F2 operator-(F2 a, F2 b) { return sfma(b, F2(-1.f, -1.f), a); }
PSFR_HD double sfma(double a, double b, double c) { return a * b + c; }
PSFR_HD float sfma(float a, float b, float c) { return a * b + c; }

struct Z2 {
    F2 x, y;
};
template <>
struct ZTraits<Z2> {
    using S = F2;
    using W = float2;   // the F2 of one component per round
    static constexpr int Rounds = 2;
};

template <class Z>
PSFR_HD Z mkz(typename ZTraits<Z>::S x, typename ZTraits<Z>::S y) {
    Z r;
    r.x = x;
    r.y = y;
    return r;
}
// twiddle table entry (double2, or float2 from the pre-rounded single-precision table) in the
// transform's precision
template <class Z, class TW>
PSFR_HD Z ztw(const TW& w) {
    using S = typename ZTraits<Z>::S;
    return mkz<Z>((S)w.x, (S)w.y);
}

template <>
PSFR_HD Z2 ztw<Z2, float2>(const float2& w) {
    Z2 r;
    r.x = F2(w.x, w.x);
    r.y = F2(w.y, w.y);
    return r;
}

template <class Z>
PSFR_HD Z cadd(Z a, Z b) { return mkz<Z>(a.x + b.x, a.y + b.y); }
template <class Z>
PSFR_HD Z csub(Z a, Z b) { return mkz<Z>(a.x - b.x, a.y - b.y); }
template <class Z>
PSFR_HD Z cmul(Z a, Z b) { return mkz<Z>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <>
PSFR_HD Z2 cmul<Z2>(Z2 a, Z2 b) { return mkz<Z2>(a.x * b.x - a.y * b.y, sfma(a.y, b.x, a.x * b.y)); }
// a + i b and a - i b (no negation: the packed type has none for free)
template <class Z>
PSFR_HD Z caddi(Z a, Z b) { return mkz<Z>(a.x - b.y, a.y + b.x); }
template <class Z>
PSFR_HD Z csubi(Z a, Z b) { return mkz<Z>(a.x + b.y, a.y - b.x); }

// ---- small DFTs, sign +i:  X[k] = sum_n x[n] exp(+2 pi i n k / R) -------------------
template <class Z>
PSFR_HD void dft4(Z& x0, Z& x1, Z& x2, Z& x3) {
    Z a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
    x0 = cadd(a, c);
    x2 = csub(a, c);
    x1 = caddi(b, d);
    x3 = csubi(b, d);
}

template <class Z>
PSFR_HD void dft8(Z* x) {
    using S = typename ZTraits<Z>::S;
    const S h = (S)0.70710678118654752440, mh = (S)-0.70710678118654752440;
    Z e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    Z o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    // o_k *= w8^k, w8 = exp(+i pi/4)
    o1 = mkz<Z>(h * (o1.x - o1.y), h * (o1.x + o1.y));
    o3 = mkz<Z>(mh * (o3.x + o3.y), h * (o3.x - o3.y));
    x[0] = cadd(e0, o0);
    x[4] = csub(e0, o0);
    x[1] = cadd(e1, o1);
    x[5] = csub(e1, o1);
    x[2] = caddi(e2, o2);   // o2 * w8^2 = i o2
    x[6] = csubi(e2, o2);
    x[3] = cadd(e3, o3);
    x[7] = csub(e3, o3);
}

// 16-point DFT as 4 x 4 (n = 4a + b, k = ka + 4 kb): radix-4 over a, twiddles w16^(b ka), radix-4 over b
template <class Z>
PSFR_HD void dft16(Z* x) {
    using S = typename ZTraits<Z>::S;
    const S h = (S)0.70710678118654752440, mh = (S)-0.70710678118654752440;
    const S c = (S)0.92387953251128675613, s = (S)0.38268343236508977173;   // cos, sin of pi/8
    const S mc = (S)-0.92387953251128675613, ms = (S)-0.38268343236508977173;
    Z y[16];   // y[b*4 + ka]
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        Z t0 = x[b], t1 = x[4 + b], t2 = x[8 + b], t3 = x[12 + b];
        dft4(t0, t1, t2, t3);
        y[b * 4 + 0] = t0;
        y[b * 4 + 1] = t1;
        y[b * 4 + 2] = t2;
        y[b * 4 + 3] = t3;
    }
    // v * (wr + i wi) without negations of packed values: re = wr x - wi y, im = wr y + wi x (mwi = -wi)
#define tw(v, wr, wi, mwi) mkz<Z>(sfma((mwi), (v).y, (wr) * (v).x), sfma((wi), (v).x, (wr) * (v).y))
    y[1 * 4 + 1] = tw(y[1 * 4 + 1], c, s, ms);                       // w16^1
    y[1 * 4 + 2] = mkz<Z>(h * (y[1 * 4 + 2].x - y[1 * 4 + 2].y), h * (y[1 * 4 + 2].x + y[1 * 4 + 2].y));   // w16^2
    y[1 * 4 + 3] = tw(y[1 * 4 + 3], s, c, mc);                       // w16^3
    y[2 * 4 + 1] = mkz<Z>(h * (y[2 * 4 + 1].x - y[2 * 4 + 1].y), h * (y[2 * 4 + 1].x + y[2 * 4 + 1].y));   // w16^2
    {
        const Z v = y[2 * 4 + 2];                                    // w16^4 = i
        y[2 * 4 + 2] = mkz<Z>(v.y * (S)(-1.0), v.x);
    }
    y[2 * 4 + 3] = mkz<Z>(mh * (y[2 * 4 + 3].x + y[2 * 4 + 3].y), h * (y[2 * 4 + 3].x - y[2 * 4 + 3].y));  // w16^6
    y[3 * 4 + 1] = tw(y[3 * 4 + 1], s, c, mc);                       // w16^3
    y[3 * 4 + 2] = mkz<Z>(mh * (y[3 * 4 + 2].x + y[3 * 4 + 2].y), h * (y[3 * 4 + 2].x - y[3 * 4 + 2].y));  // w16^6
    y[3 * 4 + 3] = tw(y[3 * 4 + 3], mc, ms, s);                      // w16^9 = -(c + i s)
#undef tw
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        Z t0 = y[ka], t1 = y[4 + ka], t2 = y[8 + ka], t3 = y[12 + ka];
        dft4(t0, t1, t2, t3);
        x[ka] = t0;
        x[ka + 4] = t1;
        x[ka + 8] = t2;
        x[ka + 12] = t3;
    }
}

template <class Z>
PSFR_HD void dft5(Z& x0, Z& x1, Z& x2, Z& x3, Z& x4) {
    using S = typename ZTraits<Z>::S;
    const S c1 = (S)0.30901699437494742410;   // cos(2pi/5)
    const S c2 = (S)-0.80901699437494742410;  // cos(4pi/5)
    const S s1 = (S)0.95105651629515357212;   // sin(2pi/5)
    const S s2 = (S)0.58778525229247312917;   // sin(4pi/5)
    const S ms1 = (S)-0.95105651629515357212;
    Z t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    Z a1 = mkz<Z>(sfma(c2, t2.x, sfma(c1, t1.x, x0.x)), sfma(c2, t2.y, sfma(c1, t1.y, x0.y)));
    Z a2 = mkz<Z>(sfma(c1, t2.x, sfma(c2, t1.x, x0.x)), sfma(c1, t2.y, sfma(c2, t1.y, x0.y)));
    Z b1 = mkz<Z>(sfma(s1, t3.x, s2 * t4.x), sfma(s1, t3.y, s2 * t4.y));
    Z b2 = mkz<Z>(sfma(ms1, t4.x, s2 * t3.x), sfma(ms1, t4.y, s2 * t3.y));
    x0 = mkz<Z>(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
    x1 = caddi(a1, b1);
    x4 = csubi(a1, b1);
    x2 = caddi(a2, b2);
    x3 = csubi(a2, b2);
}

template <int RA, class Z>
PSFR_HD void dft_small(Z* x) {
    static_assert(RA == 1 || RA == 2 || RA == 4 || RA == 8, "unsupported small radix");
    if constexpr (RA == 2) {
        Z a = x[0];
        x[0] = cadd(a, x[1]);
        x[1] = csub(a, x[1]);
    } else if constexpr (RA == 4) {
        dft4(x[0], x[1], x[2], x[3]);
    } else if constexpr (RA == 8) {
        dft8(x);
    }
}

__host__ __device__ constexpr int modinv(int a, int m) {
    for (int i = 1; i < m; ++i)
        if ((a * i) % m == 1) return i;
    return 1;
}

// Prime-factor (Good-Thomas) DFT of length R3 = 5 * RA, RA in {1,2,4,8}; natural order in
// and out.  n = (5 na + RA nb) mod R3,  k = (5 inv5 ka + RA invRA kb) mod R3.
template <int R3, class Z>
PSFR_HD void dft_r3(Z* x) {
    constexpr int RA = R3 / 5;
    constexpr int I5 = (RA == 1) ? 0 : modinv(5 % RA == 0 ? 1 : 5 % RA, RA);
    constexpr int IA = modinv(RA % 5, 5);
    Z y[R3];
#pragma unroll
    for (int nb = 0; nb < 5; ++nb) {
        Z z[RA];
#pragma unroll
        for (int na = 0; na < RA; ++na) z[na] = x[(5 * na + RA * nb) % R3];
        dft_small<RA>(z);
#pragma unroll
        for (int ka = 0; ka < RA; ++ka) y[ka * 5 + nb] = z[ka];
    }
#pragma unroll
    for (int ka = 0; ka < RA; ++ka) {
        dft5(y[ka * 5 + 0], y[ka * 5 + 1], y[ka * 5 + 2], y[ka * 5 + 3], y[ka * 5 + 4]);
#pragma unroll
        for (int kb = 0; kb < 5; ++kb) x[(5 * I5 * ka + RA * IA * kb) % R3] = y[ka * 5 + kb];
    }
}

// ---- geometry of the warp transform --------------------------------------------------
template <int R3>
struct FftGeom {
    static constexpr int N = 64 * R3;
    static constexpr int TL = 32;            // lanes per transform
    static constexpr int NQ = 64 / TL;       // (k1,k2) pairs per lane in pass 3
    static constexpr int S1 = N / 8 + 2;     // exchange-1 row stride (doubles)
    static constexpr int NAT = N + N / 16;   // skewed natural-order dump
    static constexpr int XBUF = (8 * S1 > NAT ? 8 * S1 : NAT);   // exchange 2 is compact (N words)
    static constexpr int TW1 = 5 * 7 * TL;   // double2 entries: [j][k1-1][t]
    static constexpr int TW2 = 7 * R3;       // double2 entries: [k2-1][n3]
    static_assert(N == 40 * TL, "only the one-warp-per-transform geometry is implemented");
};

PSFR_HD double comp_get(const double2& a, int c) { return c ? a.y : a.x; }
PSFR_HD void comp_set(double2& a, int c, double v) {
    if (c) a.y = v; else a.x = v;
}
// exchange word of round c: component c of a double2, the whole float2
PSFR_HD double word_get(const double2& a, int c) { return c ? a.y : a.x; }
PSFR_HD void word_set(double2& a, int c, double v) {
    if (c) a.y = v; else a.x = v;
}
PSFR_HD float2 word_get(const float2& a, int) { return a; }
PSFR_HD void word_set(float2& a, int, float2 v) { a = v; }
PSFR_HD float2 word_get(const Z2& a, int c) { return c ? a.y.v : a.x.v; }
PSFR_HD void word_set(Z2& a, int c, float2 v) {
    if (c) a.y.v = v; else a.x.v = v;
}

template <int R3, class Z, class TW>
PSFR_HD void fft_pass1(Z* v, const TW* tw1, int t) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        dft8(v + j * 8);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1)
            v[j * 8 + k1] = cmul(v[j * 8 + k1], ztw<Z>(tw1[(j * 7 + (k1 - 1)) * G::TL + t]));
    }
}

template <int R3, class Z>
PSFR_HD void fft_x1_store(const Z* v, typename ZTraits<Z>::W* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) sm[k1 * G::S1 + t + G::TL * j] = word_get(v[j * 8 + k1], c);
}

template <int R3, class Z>
PSFR_HD void fft_x1_load(Z* v, const typename ZTraits<Z>::W* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int p = t + G::TL * j;
        const int base = (p % 8) * G::S1 + (p / 8);
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) word_set(v[j * 8 + n2], c, sm[base + n2 * R3]);
    }
}

template <int R3, class Z, class TW>
PSFR_HD void fft_pass2(Z* v, const TW* tw2, int t) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int n3 = (t + G::TL * j) / 8;
        dft8(v + j * 8);
#pragma unroll
        for (int k2 = 1; k2 < 8; ++k2)
            v[j * 8 + k2] = cmul(v[j * 8 + k2], ztw<Z>(tw2[(k2 - 1) * R3 + n3]));
    }
}

template <int R3, class Z>
PSFR_HD void fft_x2_store(const Z* v, typename ZTraits<Z>::W* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int p = t + G::TL * j;
        const int n3 = p / 8;
        const int base = 64 * n3 + ((p % 8) ^ (8 * (n3 & 1)));
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) sm[base ^ (8 * k2)] = word_get(v[j * 8 + k2], c);   // q = k1 + 8 k2
    }
}

template <int R3, class Z>
PSFR_HD void fft_x2_load(Z* v, const typename ZTraits<Z>::W* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int u = 0; u < G::NQ; ++u)
#pragma unroll
        for (int n3 = 0; n3 < R3; ++n3)
            word_set(v[u * R3 + n3], c, sm[64 * n3 + ((t + G::TL * u) ^ (8 * (n3 & 1)))]);
}

template <int R3, class Z>
PSFR_HD void fft_pass3(Z* v) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int u = 0; u < G::NQ; ++u) dft_r3<R3>(v + u * R3);
}

// skewed natural-order address of output k (conflict-free dump and strided gathers)
PSFR_HD int nat_addr(int k) { return k + (k >> 4); }

template <int R3, class Z>
PSFR_HD void fft_dump(const Z* v, typename ZTraits<Z>::W* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int u = 0; u < G::NQ; ++u)
#pragma unroll
        for (int k3 = 0; k3 < R3; ++k3)
            // nat_addr(t + c) = nat_addr(t) + c + c/16 for c a multiple of 16: constant offsets
            sm[nat_addr(t) + (G::TL * u + 64 * k3) + ((G::TL * u + 64 * k3) >> 4)] = word_get(v[u * R3 + k3], c);
}

// Twiddle tables (host fills them in double precision; see psfr_api.cu / host_check.cu)
//   tw1[(j*7 + k1-1)*32 + t] = exp(+2 pi i (t+32j) k1 / N),  tw2[(k2-1)*R3 + n3] = exp(+2 pi i n3 k2 / (N/8))

#ifdef __CUDACC__
// The whole transform for one warp (device).  v: 40 points in the "load" layout on entry,
// in the "result" layout on exit.  sm: warp-private buffer of FftGeom<R3>::XBUF doubles.
template <int R3, class Z, class TW>
__device__ __forceinline__ void warp_fft(Z* v, double* sm_raw, const TW* tw1, const TW* tw2, int t) {
    using W = typename ZTraits<Z>::W;
    W* sm = reinterpret_cast<W*>(sm_raw);
    fft_pass1<R3>(v, tw1, t);
#pragma unroll
    for (int c = 0; c < ZTraits<Z>::Rounds; ++c) {
        fft_x1_store<R3>(v, sm, t, c);
        __syncwarp();
        fft_x1_load<R3>(v, sm, t, c);
        __syncwarp();
    }
    fft_pass2<R3>(v, tw2, t);
#pragma unroll
    for (int c = 0; c < ZTraits<Z>::Rounds; ++c) {
        fft_x2_store<R3>(v, sm, t, c);
        __syncwarp();
        fft_x2_load<R3>(v, sm, t, c);
        __syncwarp();
    }
    fft_pass3<R3>(v);
}
#endif

}  // namespace psfr
