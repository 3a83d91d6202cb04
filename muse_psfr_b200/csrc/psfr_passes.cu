// Generic FFT passes built on the warp transform: each warp loads one length-N line
// (two real lines packed as re/im, or a Hermitian-extended pair of half columns),
// transforms it, and writes either the transposed, untangled half-spectrum or two real
// output rows.  Used for
//   stage A  : PSD -> structure function D_unit     (psfrec.py:717-722)
//   stage B  : full-grid OTF -> PSF (parity mode)   (psfrec.py:792-801)
//   init     : pupil -> telescope OTF               (psfrec.py:784-790)
// A 2-D transform is two passes; the first writes its output transposed in 32-byte
// sectors, the second reads contiguous lines again, so no stand-alone transpose exists.
// Everything is templated on NF = dim / 1280 (see Dim<NF>, pass_kernel.cuh).
#include "pass_kernel.cuh"
#include "fast_exp.cuh"

// warps per CTA of the two tile-staged stage-A passes and whether they overlap the next tile's
// load with the transform (pass_kernel.cuh): row pass 8 warps without overlap, column pass 6
#ifndef PSFR_PASS1_WARPS
#define PSFR_PASS1_WARPS 8
#define PSFR_PASS1_OVERLAP false
#endif
#ifndef PSFR_PASS2_WARPS
#define PSFR_PASS2_WARPS 6
#define PSFR_PASS2_OVERLAP false
#endif

namespace psfr {

// ------------------------------------------------------------------ loaders
// E = (P + P reflected)/2 on rows (2rp, 2rp+1) of plane f / Pairs: real-even input whose
// transform is Re of the transform of P (psfrec.py:718,721 keeps only bg.real).
template <int NF>
struct LoadEvenRows {
    using D = Dim<NF>;
    const double* P;  // [nplanes][N][N]
    __device__ void operator()(int f, int lane, double2* v, int sub) const {
        const int plane = f / D::Pairs, rp = f % D::Pairs;
        const int a1 = 2 * rp, a2 = a1 + 1;
        const double* base = P + (size_t)plane * D::N * D::N;
        const double* r1 = base + (size_t)a1 * D::N;
        const double* r1m = base + (size_t)((D::N - a1) % D::N) * D::N;
        const bool ok2 = a2 <= D::NH;
        const double* r2 = base + (size_t)(ok2 ? a2 : 0) * D::N;
        const double* r2m = base + (size_t)(ok2 ? (D::N - a2) : 0) * D::N;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_e<NF>(i, lane, sub);
            const int nm = (D::N - n) % D::N;
            const double e1 = 0.5 * (__ldg(r1 + n) + __ldg(r1m + nm));
            const double e2 = ok2 ? 0.5 * (__ldg(r2 + n) + __ldg(r2m + nm)) : 0.0;
            v[i] = make_double2(e1, e2);
        }
    }
};

// The same even-part rows from the quadrant form of the PSD (fused path): P[a][n] =
// max(fit, AO zone) * scale2 with fit = Q[min(a, N-1-a)][min(n, N-1-n)] of the plane's draw.
template <int NF>
struct LoadEvenRowsQuad {
    using D = Dim<NF>;
    const double* Q;    // [ndraw][N/2][N/2] unscaled fitting PSD
    const double* ao;   // [nplanes][80][80] AO zones (centred, reference orientation), unscaled
    int ndir;
    double scale2;
    __device__ __forceinline__ double psd(const double* q, const double* z, int a, int n) const {
        const int qa = a < D::NH ? a : D::N - 1 - a, qn = n < D::NH ? n : D::N - 1 - n;
        double val = __ldg(q + (size_t)qa * D::NH + qn);
        constexpr int lo = D::NH - kAO / 2, hi = D::NH + kAO / 2;
        if (a >= lo && a < hi && n >= lo && n < hi) val = fmax(val, __ldg(z + (a - lo) * kAO + (n - lo)));
        return __dmul_rn(val, scale2);   // never contracted into the sum below: same rounding as the stored PSD
    }
    __device__ void operator()(int f, int lane, double2* v, int sub) const {
        const int plane = f / D::Pairs, rp = f % D::Pairs;
        const int a1 = 2 * rp, a2 = a1 + 1;
        const double* q = Q + (size_t)(plane / ndir) * D::NH * D::NH;
        const double* z = ao + (size_t)plane * kAO * kAO;
        const int a1m = (D::N - a1) % D::N;
        const bool ok2 = a2 <= D::NH;
        const int a2m = ok2 ? D::N - a2 : 0;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_e<NF>(i, lane, sub);
            const int nm = (D::N - n) % D::N;
            const double e1 = 0.5 * (psd(q, z, a1, n) + psd(q, z, a1m, nm));
            const double e2 = ok2 ? 0.5 * (psd(q, z, a2, n) + psd(q, z, a2m, nm)) : 0.0;
            v[i] = make_double2(e1, e2);
        }
    }
};

// Tile sources of the tile-staged stage-A passes (dim 1280, see tiled_pass_kernel).
// Rows: E[a][n] = (P[a][n] + P[N-a][N-n]) / 2 only touches the quadrant rows a-1, a, a+1 for the
// pair (a, a+1) = (2rp, 2rp+1): min(a, N-1-a) of rows a, a+1, N-a, N-1-a.  One contiguous block.
struct SrcEvenRowsQuad {
    using D = Dim<1>;
    static constexpr int kTileRows = 3;
    static constexpr int kTileBytes = kTileRows * D::NH * sizeof(double);
    const double* Q;    // [ndraw][N/2][N/2]
    const double* ao;   // [nplanes][80][80]
    int ndir;
    double scale2;
    __device__ static int first_row(int rp) { return max(2 * rp - 1, 0); }
    __device__ const void* block(int f, uint32_t* bytes) const {
        const int plane = f / D::Pairs, rp = f % D::Pairs;
        const int r0 = first_row(rp), r1 = min(2 * rp + 1, D::NH - 1);
        *bytes = (uint32_t)((r1 - r0 + 1) * D::NH * sizeof(double));
        return Q + ((size_t)(plane / ndir) * D::NH + r0) * D::NH;
    }
    __device__ __forceinline__ double psd(const double* q, const double* z, int r0, int a, int n) const {
        const int qa = a < D::NH ? a : D::N - 1 - a, qn = n < D::NH ? n : D::N - 1 - n;
        double val = q[(qa - r0) * D::NH + qn];
        constexpr int lo = D::NH - kAO / 2, hi = D::NH + kAO / 2;
        if (a >= lo && a < hi && n >= lo && n < hi) val = fmax(val, __ldg(z + (a - lo) * kAO + (n - lo)));
        return __dmul_rn(val, scale2);
    }
    __device__ void build(int f, int lane, const unsigned char* tile, double2* v) const {
        const int plane = f / D::Pairs, rp = f % D::Pairs;
        const int a1 = 2 * rp, a2 = a1 + 1, r0 = first_row(rp);
        const double* q = reinterpret_cast<const double*>(tile);
        const double* z = ao + (size_t)plane * kAO * kAO;
        const int a1m = (D::N - a1) % D::N;
        const bool ok2 = a2 <= D::NH;
        const int a2m = ok2 ? D::N - a2 : a1;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_n(i, lane);
            const int nm = (D::N - n) % D::N;
            const double e1 = 0.5 * (psd(q, z, r0, a1, n) + psd(q, z, r0, a1m, nm));
            const double e2 = ok2 ? 0.5 * (psd(q, z, r0, a2, n) + psd(q, z, r0, a2m, nm)) : 0.0;
            v[i] = make_double2(e1, e2);
        }
    }
};

// Columns: the Hermitian pair of line (plane, m) reads the two adjacent rows y1 = (2m + N/2) % N,
// y1 + 1 of the transposed row-pass output (the partner of the last valid row is a duplicate).
struct SrcHermitianPair {
    using D = Dim<1>;
    static constexpr int kTileBytes = 2 * D::Rows * sizeof(double2);
    const double2* Bt;  // [nplanes][N][Rows]
    int npair, last_valid;
    __device__ const void* block(int f, uint32_t* bytes) const {
        const int plane = f / npair, m = f % npair;
        *bytes = kTileBytes;
        return Bt + ((size_t)plane * D::N + (2 * m + D::NH) % D::N) * D::Rows;
    }
    __device__ void build(int f, int lane, const unsigned char* tile, double2* v) const {
        const int m = f % npair;
        const double2* c1 = reinterpret_cast<const double2*>(tile);
        const double2* c2 = (2 * m + 1 <= last_valid) ? c1 + D::Rows : c1;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_n(i, lane);
            if (n <= D::NH) {
                const double2 r1 = c1[n], r2 = c2[n];
                v[i] = make_double2(r1.x - r2.y, r1.y + r2.x);
            } else {
                const double2 r1 = c1[D::N - n], r2 = c2[D::N - n];
                v[i] = make_double2(r1.x + r2.y, r2.x - r1.y);
            }
        }
    }
};

// rows (2f, 2f+1) of exp(-c D) * OTF on the transposed half-plane (psfrec.py:793-797)
template <int NF>
struct LoadOtfRows {
    using D_ = Dim<NF>;
    const double* D;   // [Rows][N] of the selected plane
    const double* T;   // [Rows][N]
    double c;
    __device__ void operator()(int f, int lane, double2* v, int sub) const {
        const double* d1 = D + (size_t)(2 * f) * D_::N;
        const double* t1 = T + (size_t)(2 * f) * D_::N;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_e<NF>(i, lane, sub);
            v[i] = make_double2(fast_exp(-c * __ldg(d1 + n)) * __ldg(t1 + n),
                                fast_exp(-c * __ldg(d1 + D_::N + n)) * __ldg(t1 + D_::N + n));
        }
    }
};

// pupil rows (2f, 2f+1), zero-padded to N columns (psfrec.py:784-787)
template <int NF>
struct LoadPupilRows {
    using D = Dim<NF>;
    const double* pup;  // [N/2][N/2]
    __device__ void operator()(int f, int lane, double2* v, int sub) const {
        const double* p1 = pup + (size_t)(2 * f) * D::NH;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_e<NF>(i, lane, sub);
            v[i] = n < D::NH ? make_double2(__ldg(p1 + n), __ldg(p1 + D::NH + n)) : make_double2(0., 0.);
        }
    }
};

// Hermitian-extended pair of half columns: line f = (plane, m) reads buffer rows
// y(2m), y(2m+1) with y(a) = (a + N/2) % N and forms R[.,y1] + i R[.,y2] over all N rows,
// R[N-a] = conj(R[a]).  The transform of that line is real-in-two-halves: Re -> output
// row 2m, Im -> output row 2m+1.
template <int NF>
struct LoadHermitianPair {
    using D = Dim<NF>;
    const double2* Bt;  // [nplanes][N][Rows]
    int npair;          // pairs per plane
    int last_valid;     // largest valid output row (the partner of an invalid row is duplicated)
    __device__ void operator()(int f, int lane, double2* v, int sub) const {
        const int plane = f / npair, m = f % npair;
        const int o1 = 2 * m, o2 = (2 * m + 1 <= last_valid) ? 2 * m + 1 : o1;
        const double2* c1 = Bt + ((size_t)plane * D::N + (o1 + D::NH) % D::N) * D::Rows;
        const double2* c2 = Bt + ((size_t)plane * D::N + (o2 + D::NH) % D::N) * D::Rows;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_e<NF>(i, lane, sub);
            if (n <= D::NH) {
                const double2 r1 = __ldg(c1 + n), r2 = __ldg(c2 + n);
                v[i] = make_double2(r1.x - r2.y, r1.y + r2.x);
            } else {
                const double2 r1 = __ldg(c1 + (D::N - n)), r2 = __ldg(c2 + (D::N - n));
                v[i] = make_double2(r1.x + r2.y, r2.x - r1.y);
            }
        }
    }
};

// single complex column y = f, rows >= nrows are zero (forward transform of the padded pupil)
template <int NF>
struct LoadColumnPadded {
    using D = Dim<NF>;
    const double2* Bt;  // [N][Rows]
    int nrows;
    __device__ void operator()(int f, int lane, double2* v, int sub) const {
        const double2* c1 = Bt + (size_t)f * D::Rows;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
            const int n = slot_e<NF>(i, lane, sub);
            v[i] = n < nrows ? __ldg(c1 + n) : make_double2(0., 0.);
        }
    }
};

// ------------------------------------------------------------------ storers
// untangle the two real lines and write them transposed: Bt[plane][y][2rp], [2rp+1].
// UPPER = true writes only the frequencies y = 0 and N/2 .. N-1: the column pass of the structure
// function (LoadHermitianPair / SrcHermitianPair with npair = Pairs) reads rows (a + N/2) % N,
// a = 0 .. N/2, and nothing else - half of the scattered 32-byte stores of this pass were never read.
template <int NF, bool UPPER = false>
struct StoreTransposedPair {
    using D = Dim<NF>;
    double2* Bt;  // [nplanes][N][Rows]
    int npair;
    __device__ __forceinline__ void put(double2* out, const double* xb, int y) const {
        const double2 za = nat_get<NF>(xb, y), zb = nat_get<NF>(xb, (D::N - y) % D::N);
        st_global_256(out + (size_t)y * D::Rows, make_double2(0.5 * (za.x + zb.x), 0.5 * (za.y - zb.y)),
                      make_double2(0.5 * (za.y + zb.y), 0.5 * (zb.x - za.x)));
    }
    __device__ void operator()(int f, int lane, const double* xb) const {
        const int plane = f / npair, rp = f % npair;
        double2* out = Bt + (size_t)plane * D::N * D::Rows + 2 * rp;
        // unrolled far enough that a store's data registers are not rewritten while it still
        // sits in the store queue (ncu: long-scoreboard stalls on the first DADD of a trip)
        if (UPPER) {
#pragma unroll 10
            for (int i = 0; i < 20 * NF; ++i) put(out, xb, D::NH + lane + 32 * i);
            if (lane == 0) put(out, xb, 0);
        } else {
#pragma unroll 10
            for (int i = 0; i < 40 * NF; ++i) put(out, xb, lane + 32 * i);
        }
    }
};

// Re -> row 2m, Im -> row 2m+1 of out[plane][nrows_out][N], column b = (x + N/2) % N,
// value * scale * (-1)^(row + b).  The sign alternation is the output-side image of an input
// that was stored centred (fftshift-ed); `alternate` = 0 for an input in natural FFT order.
template <int NF>
struct StoreRealRows {
    using D = Dim<NF>;
    double* out;
    int npair, nrows_out, last_valid;
    double scale;
    int alternate;
    __device__ void operator()(int f, int lane, const double* xb) const {
        const int plane = f / npair, m = f % npair;
        const int o1 = 2 * m, o2 = o1 + 1;
        double* r1 = out + ((size_t)plane * nrows_out + o1) * D::N;
        const bool ok2 = o2 <= last_valid;
#pragma unroll 4
        for (int i = 0; i < 40 * NF; ++i) {
            const int b = lane + 32 * i;
            const double2 z = nat_get<NF>(xb, (b + D::NH) % D::N);
            const double s1 = (alternate && ((o1 + b) & 1)) ? -scale : scale;
            r1[b] = s1 * z.x;
            if (ok2) r1[D::N + b] = (alternate ? -s1 : s1) * z.y;
        }
    }
};

// |Z[x]|^2 -> out[y = f][x]
template <int NF>
struct StoreAbs2 {
    using D = Dim<NF>;
    double* out;  // [N][N]
    __device__ void operator()(int f, int lane, const double* xb) const {
        double* r = out + (size_t)f * D::N;
#pragma unroll 4
        for (int i = 0; i < 40 * NF; ++i) {
            const int x = lane + 32 * i;
            const double2 z = nat_get<NF>(xb, x);
            r[x] = z.x * z.x + z.y * z.y;
        }
    }
};

// Structure function rows: D[a][b] = centre - value (psfrec.py:721: 2 (bg[0,0] - bg), centred),
// written straight from the column transform, with the smallest D of each row recorded for the
// underflow cut of stage B and the pad row zeroed.  centre[plane] comes from StoreCentre below,
// which runs the very same transform on the line that holds the DC term.
template <int NF>
struct StoreDphi {
    using D = Dim<NF>;
    double* out;            // [nplanes][Rows][N]
    const double* centre;   // [nplanes]
    double* dmin;           // [nplanes][Rows]
    double scale;
    float* out32;           // single-precision copy of `out` (may be NULL)
    __device__ void operator()(int f, int lane, const double* xb) const {
        const int plane = f / D::Pairs, m = f % D::Pairs;
        const int o1 = 2 * m, o2 = o1 + 1;
        double* r1 = out + ((size_t)plane * D::Rows + o1) * D::N;
        float* q1 = out32 ? out32 + ((size_t)plane * D::Rows + o1) * D::N : nullptr;
        const bool ok2 = o2 <= D::NH;
        const double c0 = __ldg(centre + plane);
        double m1 = 1e300, m2 = ok2 ? 1e300 : 0.0;
#pragma unroll 4
        for (int i = 0; i < 40 * NF; ++i) {
            const int b = lane + 32 * i;
            const double2 z = nat_get<NF>(xb, (b + D::NH) % D::N);
            const double s1 = ((o1 + b) & 1) ? -scale : scale;
            const double d1 = c0 - s1 * z.x;
            const double d2 = ok2 ? c0 + s1 * z.y : 0.0;
            r1[b] = d1;
            r1[D::N + b] = d2;
            if (q1) {
                q1[b] = (float)d1;
                q1[D::N + b] = (float)d2;
            }
            m1 = fmin(m1, d1);
            m2 = fmin(m2, d2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m1 = fmin(m1, __shfl_xor_sync(0xffffffffu, m1, o));
            m2 = fmin(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        }
        if (lane == 0) {
            dmin[(size_t)plane * D::Rows + o1] = m1;
            dmin[(size_t)plane * D::Rows + o2] = m2;
        }
    }
};

// line f = plane: the pair that holds row N/2; keeps only centre[plane] = scale * Re X[0]
template <int NF>
struct StoreCentre {
    using D = Dim<NF>;
    double* centre;
    double scale;
    __device__ void operator()(int f, int lane, const double* xb) const {
        // row N/2 is even (o1 of its pair), column b = N/2 reads output (b + N/2) % N = 0; sign (+)
        if (lane == 0) centre[f] = scale * nat_get<NF>(xb, 0).x;
    }
};

// loader wrapper: line f of the wrapped loader is line f * stride + offset
template <class L>
struct LoadStrided {
    L inner;
    int stride, offset;
    __device__ void operator()(int f, int lane, double2* v, int sub) const { inner(f * stride + offset, lane, v, sub); }
};

// ------------------------------------------------------------------ small elementwise kernels
__global__ void pupil_kernel(double* pup, int nh, double radius, double oc) {
    // pupil_mask(N/4, N/2, oc) (psfrec.py:190-203): rho = hypot(x-c, y-c)/radius
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nh * nh) return;
    const int y = idx / nh, x = idx % nh;
    const double cc = (nh - 1) / 2.0;
    const double rho = hypot(y - cc, x - cc) / radius;
    pup[idx] = (rho < 1.0 && rho >= oc) ? 1.0 : 0.0;
}

// T = rint(autocorrelation counts) / (N^2 * sum(pupil)), pad row zero
__global__ void finalize_otf_kernel(double* t, float* t32, size_t live, size_t total, double inv_norm) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // counts are >= 0; a count that the transform left at -0.3 must become +0, not -0: the
    // single-precision grade of the row kernel builds doubles from float bit patterns
    // (f2d_bits) and relies on a clear sign bit
    const double r = (idx >= live) ? 0.0 : rint(t[idx]);
    const double v = r > 0.0 ? r * inv_norm : 0.0;
    t[idx] = v;
    if (t32) t32[idx] = (float)v;
}

// ------------------------------------------------------------------ drivers
template <int NF>
static int structure_function_t(Ctx* c, int nplanes, cudaStream_t s, bool from_quadrant, int ndir) {
    using D = Dim<NF>;
    // pass 1: rows of the even part of the PSD -> transposed half spectrum (only the frequencies pass 2 reads)
    const double k = 0.5 * 1000 / (2 * 3.141592653589793);      // rad^2 -> nm^2 (psfrec.py:151), as run_psd
    const StoreTransposedPair<NF, true> store{c->d_bt, D::Pairs};
    int rc;
    if constexpr (NF == 1) {
        if (from_quadrant)
            rc = launch_tiled_pass<PSFR_PASS1_WARPS, SrcEvenRowsQuad::kTileBytes, PSFR_PASS1_OVERLAP>(
                c, SrcEvenRowsQuad{c->d_psdq, c->d_ao, ndir, k * k}, store, nplanes * D::Pairs, s);
        else
            rc = launch_pass<NF>(c, LoadEvenRows<NF>{c->d_psd}, store, nplanes * D::Pairs, s);
    } else {
        rc = from_quadrant ? launch_pass<NF>(c, LoadEvenRowsQuad<NF>{c->d_psdq, c->d_ao, ndir, k * k}, store, nplanes * D::Pairs, s)
                           : launch_pass<NF>(c, LoadEvenRows<NF>{c->d_psd}, store, nplanes * D::Pairs, s);
    }
    if (rc) return rc;
    // pass 2: columns -> rows of the transposed structure function, 2/L^2 with L = 16 m (psfrec.py:710,718)
    const double L = 16.0;
    const double scale = 2.0 / (L * L);
    double* centre = c->d_misc + kMiscCentre;  // scratch area reserved for plane centres
    const LoadHermitianPair<NF> cols{c->d_bt, D::Pairs, D::NH};
    // the DC term first (one line per plane: the pair of row N/2), then every line with the
    // subtraction fused into the store - bit-identical to subtracting after the fact
    rc = launch_pass<NF>(c, LoadStrided<LoadHermitianPair<NF>>{cols, D::Pairs, D::NH / 2},
                         StoreCentre<NF>{centre, scale}, nplanes, s);
    if (rc) return rc;
    if constexpr (NF == 1)
        rc = launch_tiled_pass<PSFR_PASS2_WARPS, SrcHermitianPair::kTileBytes, PSFR_PASS2_OVERLAP>(c, SrcHermitianPair{c->d_bt, D::Pairs, D::NH},
                                                                StoreDphi<1>{c->d_dphi, centre, c->d_dmin, scale, c->d_dphi32},
                                                                nplanes * D::Pairs, s);
    else
        rc = launch_pass<NF>(c, cols, StoreDphi<NF>{c->d_dphi, centre, c->d_dmin, scale, c->d_dphi32}, nplanes * D::Pairs, s);
    if (rc) return rc;
    c->planes_struct = nplanes;
    return PSFR_OK;
}

int run_structure_function(Ctx* c, int nplanes, cudaStream_t s, bool from_quadrant, int ndir) {
    return c->NF == 1 ? structure_function_t<1>(c, nplanes, s, from_quadrant, ndir)
                      : structure_function_t<2>(c, nplanes, s, from_quadrant, ndir);
}

template <int NF>
static int full_psf_t(Ctx* c, int plane, double clam, double* out_dev, cudaStream_t s) {
    using D = Dim<NF>;
    const double* Dp = c->d_dphi + (size_t)plane * D::Rows * D::N;
    int rc = launch_pass<NF>(c, LoadOtfRows<NF>{Dp, c->d_otf, clam}, StoreTransposedPair<NF>{c->d_bt, D::Pairs},
                             D::Pairs, s);
    if (rc) return rc;
    // psf/psf.sum(): the sum of the raw PSF is the OTF at the origin = 1/N^2 exactly (T centre);
    // the unnormalised inverse transform carries 1/N^2 as well, so the two cancel.
    return launch_pass<NF>(c, LoadHermitianPair<NF>{c->d_bt, D::NH, D::N - 1},
                           StoreRealRows<NF>{out_dev, D::NH, D::N, D::N - 1, 1.0, 1}, D::NH, s);
}

int run_full_psf(Ctx* c, int plane, double clam, double* out_dev, cudaStream_t s) {
    return c->NF == 1 ? full_psf_t<1>(c, plane, clam, out_dev, s) : full_psf_t<2>(c, plane, clam, out_dev, s);
}

__global__ void debug_exp_kernel(const double* x, double* y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fast_exp(x[i]);
}

int run_debug_exp(Ctx* c, const double* x_dev, double* y_dev, int n, cudaStream_t s) {
    debug_exp_kernel<<<(n + 255) / 256, 256, 0, s>>>(x_dev, y_dev, n);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

template <int NF>
static int build_otf_t(Ctx* c, cudaStream_t s) {
    using D = Dim<NF>;
    pupil_kernel<<<(D::NH * D::NH + 255) / 256, 256, 0, s>>>(c->d_pup, D::NH, D::N / 4.0, 0.14);
    PSFR_LAUNCH_CHECK(c);
    // forward transform of the zero-padded pupil: rows (real pairs) then complex columns -> |.|^2
    int rc = launch_pass<NF>(c, LoadPupilRows<NF>{c->d_pup}, StoreTransposedPair<NF>{c->d_bt, D::NH / 2}, D::NH / 2, s);
    if (rc) return rc;
    rc = launch_pass<NF>(c, LoadColumnPadded<NF>{c->d_bt, D::NH}, StoreAbs2<NF>{c->d_psd}, D::N, s);
    if (rc) return rc;
    // inverse transform of |P^|^2 (real, even): N^2 x the pupil autocorrelation, centred
    rc = launch_pass<NF>(c, LoadEvenRows<NF>{c->d_psd}, StoreTransposedPair<NF>{c->d_bt, D::Pairs}, D::Pairs, s);
    if (rc) return rc;
    rc = launch_pass<NF>(c, LoadHermitianPair<NF>{c->d_bt, D::Pairs, D::NH},
                         StoreRealRows<NF>{c->d_otf, D::Pairs, D::Rows, D::NH, 1.0 / ((double)D::N * D::N), 0}, D::Pairs, s);
    if (rc) return rc;
    double centre = 0;
    PSFR_CUDA(c, cudaMemcpyAsync(&centre, c->d_otf + (size_t)D::NH * D::N + D::NH, sizeof(double),
                                 cudaMemcpyDeviceToHost, s));
    PSFR_CUDA(c, cudaStreamSynchronize(s));
    c->pup_sum = rint(centre);
    if (!(c->pup_sum > 0)) return set_error(c, PSFR_E_CUDA, "telescope OTF init failed (pupil sum %g)", centre);
    const size_t total = (size_t)D::Rows * D::N;
    finalize_otf_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        c->d_otf, c->d_otf32, (size_t)(D::NH + 1) * D::N, total, 1.0 / ((double)D::N * D::N * c->pup_sum));
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

int run_build_otf(Ctx* c, cudaStream_t s) { return c->NF == 1 ? build_otf_t<1>(c, s) : build_otf_t<2>(c, s); }

}  // namespace psfr
