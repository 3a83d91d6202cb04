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
#pragma once
#include <cuda_runtime.h>

#define PSFR_HD __host__ __device__ __forceinline__

namespace psfr {

PSFR_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
PSFR_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
PSFR_HD double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * i and a * (-i)
PSFR_HD double2 cmuli(double2 a) { return make_double2(-a.y, a.x); }
PSFR_HD double2 cmulni(double2 a) { return make_double2(a.y, -a.x); }

// ---- small DFTs, sign +i:  X[k] = sum_n x[n] exp(+2 pi i n k / R) -------------------
PSFR_HD void dft4(double2& x0, double2& x1, double2& x2, double2& x3) {
    double2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = cmuli(csub(x1, x3));
    x0 = cadd(a, c);
    x2 = csub(a, c);
    x1 = cadd(b, d);
    x3 = csub(b, d);
}

PSFR_HD void dft8(double2* x) {
    const double h = 0.70710678118654752440;
    double2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    double2 o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    // o_k *= w8^k, w8 = exp(+i pi/4)
    o1 = make_double2(h * (o1.x - o1.y), h * (o1.x + o1.y));
    o2 = cmuli(o2);
    o3 = make_double2(-h * (o3.x + o3.y), h * (o3.x - o3.y));
    x[0] = cadd(e0, o0);
    x[4] = csub(e0, o0);
    x[1] = cadd(e1, o1);
    x[5] = csub(e1, o1);
    x[2] = cadd(e2, o2);
    x[6] = csub(e2, o2);
    x[3] = cadd(e3, o3);
    x[7] = csub(e3, o3);
}

PSFR_HD void dft5(double2& x0, double2& x1, double2& x2, double2& x3, double2& x4) {
    const double c1 = 0.30901699437494742410;   // cos(2pi/5)
    const double c2 = -0.80901699437494742410;  // cos(4pi/5)
    const double s1 = 0.95105651629515357212;   // sin(2pi/5)
    const double s2 = 0.58778525229247312917;   // sin(4pi/5)
    double2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    double2 a1 = make_double2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
    double2 a2 = make_double2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
    double2 b1 = make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    double2 b2 = make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    x0 = make_double2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
    x1 = cadd(a1, cmuli(b1));
    x4 = csub(a1, cmuli(b1));
    x2 = cadd(a2, cmuli(b2));
    x3 = csub(a2, cmuli(b2));
}

template <int RA>
PSFR_HD void dft_small(double2* x);
template <>
PSFR_HD void dft_small<1>(double2*) {}
template <>
PSFR_HD void dft_small<2>(double2* x) {
    double2 a = x[0];
    x[0] = cadd(a, x[1]);
    x[1] = csub(a, x[1]);
}
template <>
PSFR_HD void dft_small<4>(double2* x) { dft4(x[0], x[1], x[2], x[3]); }
template <>
PSFR_HD void dft_small<8>(double2* x) { dft8(x); }

__host__ __device__ constexpr int modinv(int a, int m) {
    for (int i = 1; i < m; ++i)
        if ((a * i) % m == 1) return i;
    return 1;
}

// Prime-factor (Good-Thomas) DFT of length R3 = 5 * RA, RA in {1,2,4,8}; natural order in
// and out.  n = (5 na + RA nb) mod R3,  k = (5 inv5 ka + RA invRA kb) mod R3.
template <int R3>
PSFR_HD void dft_r3(double2* x) {
    constexpr int RA = R3 / 5;
    constexpr int I5 = (RA == 1) ? 0 : modinv(5 % RA == 0 ? 1 : 5 % RA, RA);
    constexpr int IA = modinv(RA % 5, 5);
    double2 y[R3];
#pragma unroll
    for (int nb = 0; nb < 5; ++nb) {
        double2 z[RA];
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

template <int R3>
PSFR_HD void fft_pass1(double2* v, const double2* tw1, int t) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        dft8(v + j * 8);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1)
            v[j * 8 + k1] = cmul(v[j * 8 + k1], tw1[(j * 7 + (k1 - 1)) * G::TL + t]);
    }
}

template <int R3>
PSFR_HD void fft_x1_store(const double2* v, double* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) sm[k1 * G::S1 + t + G::TL * j] = comp_get(v[j * 8 + k1], c);
}

template <int R3>
PSFR_HD void fft_x1_load(double2* v, const double* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int p = t + G::TL * j;
        const int base = (p % 8) * G::S1 + (p / 8);
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) comp_set(v[j * 8 + n2], c, sm[base + n2 * R3]);
    }
}

template <int R3>
PSFR_HD void fft_pass2(double2* v, const double2* tw2, int t) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int n3 = (t + G::TL * j) / 8;
        dft8(v + j * 8);
#pragma unroll
        for (int k2 = 1; k2 < 8; ++k2)
            v[j * 8 + k2] = cmul(v[j * 8 + k2], tw2[(k2 - 1) * R3 + n3]);
    }
}

template <int R3>
PSFR_HD void fft_x2_store(const double2* v, double* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int p = t + G::TL * j;
        const int n3 = p / 8;
        const int base = 64 * n3 + ((p % 8) ^ (8 * (n3 & 1)));
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) sm[base ^ (8 * k2)] = comp_get(v[j * 8 + k2], c);   // q = k1 + 8 k2
    }
}

template <int R3>
PSFR_HD void fft_x2_load(double2* v, const double* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int u = 0; u < G::NQ; ++u)
#pragma unroll
        for (int n3 = 0; n3 < R3; ++n3)
            comp_set(v[u * R3 + n3], c, sm[64 * n3 + ((t + G::TL * u) ^ (8 * (n3 & 1)))]);
}

template <int R3>
PSFR_HD void fft_pass3(double2* v) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int u = 0; u < G::NQ; ++u) dft_r3<R3>(v + u * R3);
}

// skewed natural-order address of output k (conflict-free dump and strided gathers)
PSFR_HD int nat_addr(int k) { return k + (k >> 4); }

template <int R3>
PSFR_HD void fft_dump(const double2* v, double* sm, int t, int c) {
    using G = FftGeom<R3>;
#pragma unroll
    for (int u = 0; u < G::NQ; ++u)
#pragma unroll
        for (int k3 = 0; k3 < R3; ++k3)
            // nat_addr(t + c) = nat_addr(t) + c + c/16 for c a multiple of 16: constant offsets
            sm[nat_addr(t) + (G::TL * u + 64 * k3) + ((G::TL * u + 64 * k3) >> 4)] = comp_get(v[u * R3 + k3], c);
}

// Twiddle tables (host fills them in double precision; see psfr_api.cu / host_check.cu)
//   tw1[(j*7 + k1-1)*32 + t] = exp(+2 pi i (t+32j) k1 / N),  tw2[(k2-1)*R3 + n3] = exp(+2 pi i n3 k2 / (N/8))

#ifdef __CUDACC__
// The whole transform for one warp (device).  v: 40 points in the "load" layout on entry,
// in the "result" layout on exit.  sm: warp-private buffer of FftGeom<R3>::XBUF doubles.
template <int R3>
__device__ __forceinline__ void warp_fft(double2* v, double* sm, const double2* tw1,
                                         const double2* tw2, int t) {
    fft_pass1<R3>(v, tw1, t);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        fft_x1_store<R3>(v, sm, t, c);
        __syncwarp();
        fft_x1_load<R3>(v, sm, t, c);
        __syncwarp();
    }
    fft_pass2<R3>(v, tw2, t);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        fft_x2_store<R3>(v, sm, t, c);
        __syncwarp();
        fft_x2_load<R3>(v, sm, t, c);
        __syncwarp();
    }
    fft_pass3<R3>(v);
}
#endif

}  // namespace psfr
