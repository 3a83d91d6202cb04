// Per-image tail of the path, one CTA per 40 x 40 image:
//   resample_kernel : clip >= 0, bilinear 80x80 samples -> 40x40, per-plane normalisation
//                     (psf_muse tail, psfrec.py:679-685 with interpolate :635-641)
//   (the two Moffat convolutions of convolve_final_psf live in psfr_conv.cu)
//   fit_kernel      : 5-parameter circular Moffat least squares by Levenberg-Marquardt with
//                     analytic Jacobian (fit_psf_cube -> mpdaf moffat_fit, :861-871)
//   mean / polyfit  : time mean of cubes (:1104) and polynomial smoothing (:1174-1210)
#include <algorithm>
#include "psfr_internal.h"
#include "fast_exp.cuh"

namespace psfr {

constexpr int kImg = kPSF * kPSF;  // 1600

// deterministic block sum (blockDim.x multiple of 32, <= 1024); result valid in all threads
__device__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

// ---------------------------------------------------------------- Moffat kernels
// astropy Moffat2DKernel(gamma, alpha, x_size=41, y_size=41): amplitude (alpha-1)/(pi gamma^2),
// (1 + r^2/gamma^2)^(-alpha) at integer offsets, normalised to unit sum.
__global__ void moffat_kernels_kernel(const double* __restrict__ gam, const double* __restrict__ alp,
                                      int alp_stride, double* __restrict__ out) {
    __shared__ double red[32];
    const int k = blockIdx.x;
    const double g = gam[k], a = alp[k * alp_stride];
    const double amp = (a - 1) / (3.141592653589793 * g * g);
    double vals[7];
    double part = 0.0;
    int cnt = 0;
    for (int idx = threadIdx.x; idx < kKW * kKW; idx += blockDim.x, ++cnt) {
        const int dy = idx / kKW - kKW / 2, dx = idx % kKW - kKW / 2;
        const double rr = (double)(dx * dx + dy * dy) / (g * g);
        vals[cnt] = amp * pow(1 + rr, -a);
        part += vals[cnt];
    }
    const double tot = block_sum(part, red);
    cnt = 0;
    for (int idx = threadIdx.x; idx < kKW * kKW; idx += blockDim.x, ++cnt)
        out[(size_t)k * kKW * kKW + idx] = vals[cnt] / tot;
}

// ---------------------------------------------------------------- resample
__global__ void __launch_bounds__(256)
resample_kernel(const double* __restrict__ samp, const double* __restrict__ frac, int nlam,
                double* __restrict__ cube) {
    extern __shared__ double dyn_smem[];
    double* s = dyn_smem;               // kNS*kNS
    double* red = dyn_smem + kNS * kNS; // 32
    const int img = blockIdx.x, lam = img % nlam;
    const double* src = samp + (size_t)img * kNS * kNS;
    for (int i = threadIdx.x; i < kNS * kNS; i += blockDim.x) s[i] = fmax(src[i], 0.0);   // :680
    __syncthreads();
    const double* fr = frac + (size_t)lam * kPSF;
    double vals[7];
    double part = 0.0;
    int cnt = 0;
    for (int idx = threadIdx.x; idx < kImg; idx += blockDim.x, ++cnt) {
        const int y = idx / kPSF, x = idx % kPSF;
        const double fy = fr[y], fx = fr[x];
        const double* r0 = s + (2 * y) * kNS + 2 * x;
        const double v = (1 - fy) * ((1 - fx) * r0[0] + fx * r0[1]) +
                         fy * ((1 - fx) * r0[kNS] + fx * r0[kNS + 1]);
        vals[cnt] = v;
        part += v;
    }
    const double tot = block_sum(part, red);   // :685
    cnt = 0;
    for (int idx = threadIdx.x; idx < kImg; idx += blockDim.x, ++cnt)
        cube[(size_t)img * kImg + idx] = vals[cnt] / tot;
}

// ---------------------------------------------------------------- Moffat fit
// model f = I (1 + ((p-y0)^2 + (q-x0)^2)/a^2)^(-n), parameters x = [I, y0, x0, a, n].
// One WARP per image (4 images per CTA): the image sits in shared memory, every lane owns
// npx/32 pixels, the 21 sums of the normal equations are reduced with xor-shuffles (all lanes
// end with bit-identical sums, so the 5x5 solve and the LM control flow run redundantly and
// uniformly in every lane) - no block-level synchronisation at all.
constexpr int kFitWarps = 4;
// resident CTAs per SM the fitter is compiled for (register budget 65536 / (128 * this))
#ifndef PSFR_FIT_MINBLOCKS
#define PSFR_FIT_MINBLOCKS 4
#endif
// The two-stage fit: pixel stride and squared relative-step tolerance of the coarse stage, squared
// relative-step tolerance of the final stage = (1.49e-8)^2, the xtol at which MINPACK - the reference's
// solver behind mpdaf - stops.  Measured on the 4480 images of a config-4 chunk (tools/fit_bench.py):
// stride 4 / final 1e-9: 0.646 ms; stride 8: 0.600; final 1.49e-8: 0.610; both: 0.556 ms with FWHM / beta
// within 1.5e-8 / 4.3e-8 of the tighter solution (bar 1e-5); strides 16 and 32 need more iterations than they save.
#ifndef PSFR_FIT_COARSE_STEP
#define PSFR_FIT_COARSE_STEP 8
#endif
#ifndef PSFR_FIT_COARSE_TOL2
#define PSFR_FIT_COARSE_TOL2 1e-6
#endif
#ifndef PSFR_FIT_TOL2
#define PSFR_FIT_TOL2 2.2e-16
#endif
constexpr int kNP = 5;
constexpr int kNSUM = 21;  // 15 (J^T J upper) + 5 (J^T r) + 1 (cost)

// packed-upper index of (i, j), i <= j
__device__ __forceinline__ constexpr int pk(int i, int j) { return i * kNP - i * (i - 1) / 2 + (j - i); }

__device__ __forceinline__ void warp_sum_vec(double* v) {
#pragma unroll
    for (int k = 0; k < kNSUM; ++k) {
        double t = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        v[k] = t;
    }
}

// 1/u for u in [1, 1e300): hardware seed + two Newton steps (no special cases to test for)
__device__ __forceinline__ double rcp_pos(double u) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(u));
    double e = fma(-u, r, 1.0);
    r = fma(r, e, r);
    e = fma(-u, r, 1.0);
    return fma(r, e, r);
}

// log(u) for u in [1, 2^60): u = 2^k m with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(s),
// s = (m - 1)/(m + 1), |s| <= 0.1716, series to s^23; k ln2 added in two words.  Branch-free and
// half as long as the library log, whose dependent chain dominated the per-pixel latency of the
// fitter (the kernel runs one image per warp in a single wave).  Max relative error 3.8e-16
// against a long-double log1p over [1 + 1e-8, 1e4] (the argument is 1 + rho^2 here).
// series coefficients 1/23 ... 1/3 and the two words of ln 2 in constant memory (see fast_exp.cuh: immediates
// cost two uniform-register moves per use)
static __constant__ double kLogPoly[11] = {1.0 / 23, 1.0 / 21, 1.0 / 19, 1.0 / 17, 1.0 / 15, 1.0 / 13,
                                           1.0 / 11, 1.0 / 9,  1.0 / 7,  1.0 / 5,  1.0 / 3};
static __constant__ double kLn2[2] = {6.93147180369123816490e-01, 1.90821492927058770002e-10};
__device__ __forceinline__ double log_ge1(double u) {
    int hi = __double2hiint(u);
    const int k = (hi - 0x3fe6a09e) >> 20;
    hi -= k << 20;
    const double m = __hiloint2double(hi, __double2loint(u));
    const double s = (m - 1.0) * rcp_pos(m + 1.0);
    const double z = s * s;
    double p = kLogPoly[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = fma(p, z, kLogPoly[i]);
    const double kf = (double)k;
    double r = fma(s * z, 2.0 * p, kf * kLn2[1]);
    r += 2.0 * s;
    return fma(kf, kLn2[0], r);
}

// one pixel's contribution to the normal equations
__device__ __forceinline__ void accumulate_pixel(double dp, double dq, double pix, double I, double n, double ia2,
                                                 double ia, double* sums) {
    const double rho2 = (dp * dp + dq * dq) * ia2;
    const double uu = 1.0 + rho2;
    const double lu = log_ge1(uu);
    const double m = fast_exp(-n * lu);        // u^-n
    const double f = I * m;
    const double g = 2.0 * n * f * rcp_pos(uu);  // 2 I n u^(-n-1)
    double J[kNP];
    J[0] = m;
    J[1] = g * dp * ia2;
    J[2] = g * dq * ia2;
    J[3] = g * rho2 * ia;
    J[4] = -f * lu;
    const double r = f - pix;
#pragma unroll
    for (int i = 0; i < kNP; ++i)
#pragma unroll
        for (int j = i; j < kNP; ++j) sums[pk(i, j)] = fma(J[i], J[j], sums[pk(i, j)]);
#pragma unroll
    for (int i = 0; i < kNP; ++i) sums[15 + i] = fma(J[i], r, sums[15 + i]);
    sums[20] = fma(r, r, sums[20]);
}

// accumulate normal equations at x over this lane's pixels (two independent pixel chains per
// trip so that the log / exp latencies overlap), then reduce over the warp.  STEP = 1 visits every
// pixel; STEP = 4 every fourth pixel of each lane (a quasi-uniform quarter of the image), used
// by the coarse first stage of the fit.
template <int STEP>
__device__ __forceinline__ void accumulate(const double* __restrict__ img, int npx, int nx, const double* x,
                                           double* sums) {
#pragma unroll
    for (int k = 0; k < kNSUM; ++k) sums[k] = 0.0;
    const double I = x[0], y0 = x[1], x0 = x[2], a = x[3], n = x[4];
    const double ia2 = 1.0 / (a * a), ia = 1.0 / a;
    const int lane = threadIdx.x & 31;
    int pr = lane / nx, qc = lane % nx;          // once per call; then stepped incrementally
    const int dpr = (32 * STEP) / nx, dqc = (32 * STEP) % nx;
    auto step = [&]() {
        pr += dpr;
        qc += dqc;
        if (qc >= nx) {
            qc -= nx;
            ++pr;
        }
    };
    int idx = lane;
    for (; idx + 32 * STEP < npx; idx += 64 * STEP) {
        const double dp0 = (double)pr - y0, dq0 = (double)qc - x0;
        step();
        const double dp1 = (double)pr - y0, dq1 = (double)qc - x0;
        step();
        accumulate_pixel(dp0, dq0, img[idx], I, n, ia2, ia, sums);
        accumulate_pixel(dp1, dq1, img[idx + 32 * STEP], I, n, ia2, ia, sums);
    }
    if (idx < npx) accumulate_pixel((double)pr - y0, (double)qc - x0, img[idx], I, n, ia2, ia, sums);
    warp_sum_vec(sums);
}

// Cholesky of A + mu diag(A) (A from the packed sums), fully unrolled in registers.
// L holds the factor with the RECIPROCAL of the diagonal on the diagonal.
__device__ __forceinline__ bool cholesky(const double* s, double mu, double (&L)[kNP][kNP]) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < kNP; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double t = s[pk(j, i)];
            if (i == j) t *= (1.0 + mu);
#pragma unroll
            for (int q = 0; q < j; ++q) t = fma(-L[i][q], L[j][q], t);
            if (i == j) {
                ok = ok && (t > 0.0);
                L[i][i] = rsqrt(t);
            } else {
                L[i][j] = t * L[j][j];
            }
        }
    }
    return ok;
}

// solve L L^T z = b in place
__device__ __forceinline__ void chol_solve(const double (&L)[kNP][kNP], double* z) {
#pragma unroll
    for (int i = 0; i < kNP; ++i) {
        double t = z[i];
#pragma unroll
        for (int q = 0; q < i; ++q) t = fma(-L[i][q], z[q], t);
        z[i] = t * L[i][i];
    }
#pragma unroll
    for (int i = kNP - 1; i >= 0; --i) {
        double t = z[i];
#pragma unroll
        for (int q = i + 1; q < kNP; ++q) t = fma(-L[q][i], z[q], t);
        z[i] = t * L[i][i];
    }
}

// Levenberg-Marquardt on the pixel subset STEP from the start point x (updated in place);
// sums holds the normal equations at the final x.  tol2 = square of the relative step at which
// to stop (MINPACK, the reference's solver, stops at 1.49e-8; so does the final stage).
template <int STEP>
__device__ __forceinline__ void lm_solve(const double* __restrict__ img, int npx, int nx, double* x, double* sums,
                                         double tol2, int max_iter, int& iter, int& status) {
    double trial[kNSUM], xt[kNP], d[kNP];
    accumulate<STEP>(img, npx, nx, x, sums);
    double mu = 1e-3, nu = 2.0;
    status = -1;
    constexpr int dia[kNP] = {pk(0, 0), pk(1, 1), pk(2, 2), pk(3, 3), pk(4, 4)};
    double L[kNP][kNP];
    for (iter = 1; iter <= max_iter; ++iter) {
        // every lane runs the identical solve on bit-identical sums
        if (!cholesky(sums, mu, L)) {
            mu *= nu;
            nu *= 2.0;
            if (mu > 1e15) break;
            continue;
        }
#pragma unroll
        for (int i = 0; i < kNP; ++i) d[i] = -sums[15 + i];
        chol_solve(L, d);
        double dn = 0.0, xn = 0.0, pred = 0.0;
#pragma unroll
        for (int i = 0; i < kNP; ++i) {
            xt[i] = x[i] + d[i];
            const double sc2 = sums[dia[i]];                  // |J column|^2
            dn = fma(d[i] * d[i], sc2, dn);
            xn = fma(x[i] * x[i], sc2, xn);
            // predicted decrease of r.r : d.(mu D d - g)
            pred += d[i] * (mu * sc2 * d[i] - sums[15 + i]);
        }
        if (!(xt[3] > 0.0) || !(xt[4] > 0.0) || !isfinite(xt[0])) {
            mu *= nu;
            nu *= 2.0;
            if (mu > 1e15) break;
            continue;
        }
        accumulate<STEP>(img, npx, nx, xt, trial);
        const double actual = sums[20] - trial[20];
        // near the minimum the cost is flat to rounding: tolerate a noise-level increase so that
        // Gauss-Newton steps keep contracting, and stop on the (column-scaled) step size
        // MINPACK (the reference, through mpdaf -> scipy leastsq) stops at a relative step of 1.49e-8
        const bool small = dn <= tol2 * (xn + 1e-300);            // relative step <= sqrt(tol2)
        if (isfinite(trial[20]) && actual >= -1e-13 * sums[20]) {
            const double rho = pred > 0.0 ? actual / pred : 1.0;
#pragma unroll
            for (int i = 0; i < kNP; ++i) x[i] = xt[i];
#pragma unroll
            for (int k = 0; k < kNSUM; ++k) sums[k] = trial[k];
            const double t = 2.0 * rho - 1.0;
            mu *= fmax(1.0 / 3.0, 1.0 - t * t * t);
            if (mu < 1e-15) mu = 1e-15;
            nu = 2.0;
            if (small) {
                status = iter;
                break;
            }
        } else {
            if (dn <= 1e-18 * (xn + 1e-300)) {   // step below 1e-9 relative and no decrease: rounding floor
                status = iter;
                break;
            }
            mu *= nu;
            nu *= 2.0;
            if (mu > 1e15) break;
        }
    }
}

__global__ void __launch_bounds__(kFitWarps * 32, PSFR_FIT_MINBLOCKS)
fit_kernel(const double* __restrict__ imgs, int nimg, int ny, int nx, double* __restrict__ out,
           int* __restrict__ next_image) {
    extern __shared__ double sm[];
    const int npx = ny * nx, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* img = sm + (size_t)warp * (npx + nx);    // npx pixels + nx column weights
    // images are handed to the warps by a global counter: the LM iteration count varies 3x
    // between images, so a static one-image-per-warp grid ends in a long tail
#pragma unroll 1
    for (;;) {
    int image = 0;
    if (lane == 0) image = atomicAdd(next_image, 1);
    image = __shfl_sync(0xffffffffu, image, 0);
    if (image >= nimg) return;                       // whole warp leaves; no block barrier is used below
    __syncwarp();
    double* colw = img + npx;
    const double* src = imgs + (size_t)image * npx;
    for (int i = lane; i < npx; i += 32) img[i] = src[i];
    __syncwarp();
    // ---- start point as mpdaf: peak pixel, width from Image.moments(), n = 2
    for (int q = lane; q < nx; q += 32) {
        double t = 0.0;
        for (int p = 0; p < ny; ++p) t += p * fabs(img[p * nx + q]);
        colw[q] = t;
    }
    __syncwarp();
    // first maximum of colw (column of the weighted sums) and of the image (peak pixel)
    double bestw = -1.0, bestv = -1e300;
    int qb = 0, ib = 0;
    for (int q = lane; q < nx; q += 32)
        if (colw[q] > bestw) { bestw = colw[q]; qb = q; }
    for (int i = lane; i < npx; i += 32)
        if (img[i] > bestv) { bestv = img[i]; ib = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ow = __shfl_xor_sync(0xffffffffu, bestw, o);
        const int oq = __shfl_xor_sync(0xffffffffu, qb, o);
        if (ow > bestw || (ow == bestw && oq < qb)) { bestw = ow; qb = oq; }
        const double ov = __shfl_xor_sync(0xffffffffu, bestv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ib, o);
        if (ov > bestv || (ov == bestv && oi < ib)) { bestv = ov; ib = oi; }
    }
    double num = 0.0, den = 0.0;
    for (int p = lane; p < ny; p += 32) {
        const double v = img[p * nx + qb];
        num += fabs((p - qb) * v);
        den += fabs(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    double x[kNP], sums[kNSUM];
    {
        const double wp = sqrt(num / den);
        const double fwhm0 = wp * 2.0 * sqrt(2.0 * log(2.0));
        x[0] = bestv;
        x[1] = ib / nx;
        x[2] = ib % nx;
        x[3] = fwhm0 / (2.0 * sqrt(sqrt(2.0) - 1.0));
        x[4] = 2.0;
    }
    // Stage 1: Levenberg-Marquardt on every eighth pixel down to a relative step of 1e-3 - it
    // only moves the start point (the converged minimum does not depend on it); stage 2: all
    // pixels down to MINPACK's own tolerance, 1.49e-8.
    int iter = 0, status = -1;
    lm_solve<PSFR_FIT_COARSE_STEP>(img, npx, nx, x, sums, PSFR_FIT_COARSE_TOL2, 12, iter, status);
    const int coarse_iter = iter;
    lm_solve<1>(img, npx, nx, x, sums, PSFR_FIT_TOL2, 200, iter, status);
    iter += coarse_iter;
    if (status > 0) status += coarse_iter;
    const int max_iter = 212;
    if (lane == 0) {
        double* o = out + (size_t)image * PSFR_FIT_NPAR;
        const double a = fabs(x[3]), n = x[4];
        const double kf = 2.0 * sqrt(pow(2.0, 1.0 / n) - 1.0);
        o[PSFR_FIT_PEAK] = x[0];
        o[PSFR_FIT_Y0] = x[1];
        o[PSFR_FIT_X0] = x[2];
        o[PSFR_FIT_ALPHA] = a;
        o[PSFR_FIT_N] = n;
        o[PSFR_FIT_FWHM] = a * kf;
        o[PSFR_FIT_CHISQ] = sums[20];
        o[PSFR_FIT_ITER] = (status > 0) ? (double)status : -(double)(iter > max_iter ? max_iter : iter);
        const double dof = (double)(npx - kNP);
        double L[kNP][kNP];
        if (cholesky(sums, 0.0, L)) {
            double e[kNP];
#pragma unroll
            for (int c = 0; c < kNP; ++c) {
                double z[kNP];
#pragma unroll
                for (int i = 0; i < kNP; ++i) z[i] = (i == c) ? 1.0 : 0.0;
                chol_solve(L, z);
                e[c] = sqrt(fabs(z[c]) * fabs(sums[20] / dof));
            }
            o[PSFR_FIT_ERR_PEAK] = e[0];
            o[PSFR_FIT_ERR_Y0] = e[1];
            o[PSFR_FIT_ERR_X0] = e[2];
            o[PSFR_FIT_ERR_ALPHA] = e[3];
            o[PSFR_FIT_ERR_N] = e[4];
            // mpdaf Image.moffat_fit (fit_n, circular): err_fwhm = err_a * n; err_flux = err_I err_n err_a^2 err_e
            // with err_e = 0 for a circular fit (restated from mpdaf 3.x; the package is not in the reference tree)
            o[PSFR_FIT_ERR_FWHM] = e[3] * n;
            o[PSFR_FIT_ERR_FLUX] = e[0] * e[4] * e[3] * e[3] * 0.0;
        } else {
            for (int i = PSFR_FIT_ERR_PEAK; i <= PSFR_FIT_ERR_FWHM; ++i) o[i] = nan("");
            o[PSFR_FIT_ERR_FLUX] = nan("");
        }
        o[PSFR_FIT_FLUX] = x[0] / (n - 1.0) * (3.141592653589793 * a * a);
    }
    __syncwarp();
    }   // next image
}

// ---------------------------------------------------------------- mean of cubes
__global__ void mean_kernel(const double* __restrict__ cubes, int ncube, size_t elems,
                            double* __restrict__ out, int first, int last, int ntotal) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= elems) return;
    // straight sum in input order (np.mean over axis 0 reduces plane by plane), then / ntotal;
    // `first` / `last` let the caller stream the cubes through in slabs with identical rounding
    double t = first ? 0.0 : out[i];
    for (int k = 0; k < ncube; ++k) t += cubes[(size_t)k * elems + i];
    out[i] = last ? t / ntotal : t;
}

// ---------------------------------------------------------------- polynomial smoothing
// least squares via Householder QR of the shared Vandermonde matrix (columns x^deg .. x^0,
// as np.polyfit); one thread per series applies Q^T and back-substitutes.
constexpr int kMaxLam = 256, kMaxDeg = 8;
__global__ void polyfit_kernel(const double* __restrict__ lb, int nlam, int deg, int nseries,
                               const double* __restrict__ y, double* __restrict__ coef) {
    __shared__ double V[kMaxLam * (kMaxDeg + 1)];   // column-major, overwritten by R / reflectors
    __shared__ double beta[kMaxDeg + 1], rdiag[kMaxDeg + 1];
    const int nc = deg + 1;
    for (int idx = threadIdx.x; idx < nlam * nc; idx += blockDim.x) {
        const int c = idx / nlam, r = idx % nlam;
        V[c * nlam + r] = pow(lb[r], (double)(deg - c));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < nc; ++c) {
            double nrm = 0.0;
            for (int r = c; r < nlam; ++r) nrm += V[c * nlam + r] * V[c * nlam + r];
            nrm = sqrt(nrm);
            const double alpha = V[c * nlam + c] > 0 ? -nrm : nrm;
            V[c * nlam + c] -= alpha;                    // v = x - alpha e1
            double vv = 0.0;
            for (int r = c; r < nlam; ++r) vv += V[c * nlam + r] * V[c * nlam + r];
            beta[c] = vv > 0 ? 2.0 / vv : 0.0;
            rdiag[c] = alpha;
            for (int c2 = c + 1; c2 < nc; ++c2) {
                double dot = 0.0;
                for (int r = c; r < nlam; ++r) dot += V[c * nlam + r] * V[c2 * nlam + r];
                dot *= beta[c];
                for (int r = c; r < nlam; ++r) V[c2 * nlam + r] -= dot * V[c * nlam + r];
            }
        }
    }
    __syncthreads();
    for (int sidx = threadIdx.x; sidx < nseries; sidx += blockDim.x) {
        double w[kMaxLam];
        for (int r = 0; r < nlam; ++r) w[r] = y[(size_t)sidx * nlam + r];
        for (int c = 0; c < nc; ++c) {
            double dot = 0.0;
            for (int r = c; r < nlam; ++r) dot += V[c * nlam + r] * w[r];
            dot *= beta[c];
            for (int r = c; r < nlam; ++r) w[r] -= dot * V[c * nlam + r];
        }
        double cf[kMaxDeg + 1];
        for (int c = nc - 1; c >= 0; --c) {
            double t = w[c];
            for (int c2 = c + 1; c2 < nc; ++c2) t -= V[c2 * nlam + c] * cf[c2];
            cf[c] = t / rdiag[c];
        }
        for (int c = 0; c < nc; ++c) coef[(size_t)sidx * nc + c] = cf[c];
    }
}

// ---------------------------------------------------------------- drivers
int run_build_kernels(Ctx* c, int ndraw, int nlam, const double* lambda_nm_host, bool tt, bool mu,
                      cudaStream_t s) {
    if (tt) {
        // gamma = alpha_tt per draw (slot PSFR_DRAW_ALPHA_TT), Moffat power 2
        double* two = c->d_misc + kMiscTwo;   // holds the constant 2.0
        const double h2 = 2.0;
        PSFR_CUDA(c, cudaMemcpyAsync(two, &h2, sizeof(double), cudaMemcpyHostToDevice, s));
        moffat_kernels_kernel<<<ndraw, 256, 0, s>>>(c->d_misc + misc_alpha_tt(c->max_planes), two, 0, c->d_kern_tt);
        PSFR_LAUNCH_CHECK(c);
        int rc = run_kernel_spectra(c, ndraw, c->d_kern_tt, c->d_khat_tt, s);
        if (rc) return rc;
    }
    if (mu && (int)c->lam_kernels.size() == nlam &&
        std::equal(lambda_nm_host, lambda_nm_host + nlam, c->lam_kernels.begin()))
        mu = false;   // the MUSE kernels of these wavelengths and their spectra are already on the device
    if (mu) {
        c->lam_kernels.clear();
        // muse_intrinsic_psf (psfrec.py:1160-1168, np.polyval = Horner) and alpha = fwhm/0.2/(2 sqrt(2^(1/beta)-1)) (:923-924)
        static const double pol_beta[6] = {-0.83704697, 1.1337153, 0.0609222, -1.35581762, 1.15237178, 2.2106042};
        static const double pol_fwhm[6] = {0.60467385, -1.58905792, 1.75293264, -1.0368302, 0.21487023, 0.34851139};
        double* h = static_cast<double*>(c->h_pinned);
        for (int l = 0; l < nlam; ++l) {
            const double lb = (10 * lambda_nm_host[l] - 4750) / (9350 - 4750);
            double fw = 0.0, be = 0.0;
            for (int k = 0; k < 6; ++k) {
                fw = fw * lb + pol_fwhm[k];
                be = be * lb + pol_beta[k];
            }
            fw = fw / 0.2;
            h[l] = fw / (2 * sqrt(pow(2.0, 1. / be) - 1));
            h[nlam + l] = be;
        }
        double* d = c->d_misc + kMiscMuse;    // gamma[nlam], beta[nlam]
        PSFR_CUDA(c, cudaMemcpyAsync(d, h, 2 * nlam * sizeof(double), cudaMemcpyHostToDevice, s));
        PSFR_CUDA(c, cudaStreamSynchronize(s));   // pinned bounce buffer is reused by the caller
        moffat_kernels_kernel<<<nlam, 256, 0, s>>>(d, d + nlam, 1, c->d_kern_mu);
        PSFR_LAUNCH_CHECK(c);
        int rc = run_kernel_spectra(c, nlam, c->d_kern_mu, c->d_khat_mu, s);
        if (rc) return rc;
        c->lam_kernels.assign(lambda_nm_host, lambda_nm_host + nlam);
    }
    return PSFR_OK;
}

int run_resample(Ctx* c, int nimg, int nlam, double* cube_dev, cudaStream_t s) {
    const size_t smem = (size_t)(kNS * kNS + 32) * sizeof(double);
    if (int rc = ensure_dynamic_smem(c, resample_kernel, smem)) return rc;
    resample_kernel<<<nimg, 256, smem, s>>>(c->d_samp, c->d_frac, nlam, cube_dev);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

int run_convolve(Ctx* c, int ndraw, int nlam, const double* in_dev, double* out_dev, cudaStream_t s) {
    return run_fft_convolve(c, ndraw, nlam, in_dev, out_dev, s);
}

int run_fit(Ctx* c, int nimg, int ny, int nx, const double* img_dev, double* fit_dev, cudaStream_t s) {
    const size_t smem = (size_t)kFitWarps * (ny * nx + nx) * sizeof(double);
    if (smem > 200 * 1024) return set_error(c, PSFR_E_UNSUPPORTED, "image %dx%d too large for the fitter", ny, nx);
    if (int rc = ensure_dynamic_smem(c, fit_kernel, smem)) return rc;
    int grid = (nimg + kFitWarps - 1) / kFitWarps;
    int per_sm = 2;                                  // persistent grid: as many CTAs as are resident at once
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_kernel, kFitWarps * 32, smem) != cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        per_sm = 2;
    }
    const int resident = c->sm_count * per_sm;
    if (grid > resident) grid = resident;
    PSFR_CUDA(c, cudaMemsetAsync(c->d_counter + 1, 0, sizeof(int), s));
    fit_kernel<<<grid, kFitWarps * 32, smem, s>>>(img_dev, nimg, ny, nx, fit_dev, c->d_counter + 1);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

int run_mean(Ctx* c, int ncube, int plane_elems, const double* cubes_dev, double* out_dev, cudaStream_t s,
             bool first, bool last, int ntotal) {
    mean_kernel<<<(plane_elems + 255) / 256, 256, 0, s>>>(cubes_dev, ncube, (size_t)plane_elems, out_dev,
                                                          first ? 1 : 0, last ? 1 : 0, ntotal);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

int run_polyfit(Ctx* c, int nseries, int nlam, int deg, const double* lb_dev, const double* y_dev,
                double* coef_dev, cudaStream_t s) {
    if (nlam > kMaxLam || deg > kMaxDeg || deg + 1 > nlam)
        return set_error(c, PSFR_E_UNSUPPORTED, "polyfit limits: nlam <= %d, deg <= %d", kMaxLam, kMaxDeg);
    polyfit_kernel<<<1, 128, 0, s>>>(lb_dev, nlam, deg, nseries, y_dev, coef_dev);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

}  // namespace psfr
