// exp(x) for the OTF evaluation exp(-c_lambda * D) (psfrec.py:793-794), x <= ~0.
//
// Branch-free: Cody-Waite reduction x = k ln2 + r with a two-word ln2, degree-11 near-minimax
// polynomial on |r| <= ln2/2 (approximation error 4e-18, evaluated error <= 1 ulp against
// mpmath), scale by 2^k built in the exponent field.  No range checks: k is clamped at -1000,
// so arguments below ~-693 return a value < 1e-301 instead of a denormal/zero - far below
// anything the 1e-9 parity bar can see.  The argument must not exceed ~ +700.
// Written so that several independent evaluations interleave (no branches, plain FMA chains).
#pragma once
#include <cuda_runtime.h>

namespace psfr {

// Polynomial coefficients in constant memory: as immediates the compiler re-materialises every one of them
// through two uniform-register moves per use (ncu: UMOV was 16 % of the row kernel's instructions); a
// constant-bank operand rides along with the DFMA.
static __constant__ double kExpPoly[11] = {
    2.5110049204818659793e-8, 2.763265472252779189e-7, 2.7557240887229868596e-6, 0.000024801485441561312966,
    0.00019841269890076402829, 0.0013888888952352862866, 0.0083333333333195896163, 0.04166666666648795252,
    0.1666666666666668082, 0.50000000000000184039, 1.0};
static __constant__ double kExpRed[4] = {1.44269504088896338700e+00, 6755399441055744.0, -6.93147180369123816490e-01,
                                         -1.90821492927058770002e-10};

__device__ __forceinline__ double fast_exp_imm(double x) {
    const double L2E = 1.44269504088896338700e+00;
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: round-to-nearest integer trick
    const double LN2_HI = 6.93147180369123816490e-01;   // low 21 bits zero: k * LN2_HI is exact
    const double LN2_LO = 1.90821492927058770002e-10;
    const double kd = fma(x, L2E, MAGIC);
    int k = __double2loint(kd);
    const double kf = kd - MAGIC;
    double r = fma(kf, -LN2_HI, x);
    r = fma(kf, -LN2_LO, r);
    double p = 2.5110049204818659793e-8;
    p = fma(p, r, 2.763265472252779189e-7);
    p = fma(p, r, 2.7557240887229868596e-6);
    p = fma(p, r, 0.000024801485441561312966);
    p = fma(p, r, 0.00019841269890076402829);
    p = fma(p, r, 0.0013888888952352862866);
    p = fma(p, r, 0.0083333333333195896163);
    p = fma(p, r, 0.04166666666648795252);
    p = fma(p, r, 0.1666666666666668082);
    p = fma(p, r, 0.50000000000000184039);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    k = max(k, -1000);
    // p in [0.70, 1.42]: scaling by 2^k is an integer add on the exponent field (no denormals for
    // k >= -1000), which keeps one multiply per evaluation off the FP64 pipe
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// the same evaluation with every constant read from the constant bank
__device__ __forceinline__ double fast_exp(double x) {
    const double MAGIC = kExpRed[1];
    const double kd = fma(x, kExpRed[0], MAGIC);
    int k = __double2loint(kd);
    const double kf = kd - MAGIC;
    double r = fma(kf, kExpRed[2], x);
    r = fma(kf, kExpRed[3], r);
    double p = kExpPoly[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = fma(p, r, kExpPoly[i]);
    p = fma(p, r, 1.0);
    k = max(k, -1000);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// 2^x on the special-function unit (MUFU.EX2), relative error 2^-22; flushes to zero below 2^-126.
// Used only where the result is below exp(-20) of the OTF peak (psfr_hot.cu).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// float -> double on the integer pipe (F2F issues at a quarter of the DFMA rate): exact for
// normal non-negative floats; +0 and denormals map to values below 2^-126, which is as good as
// zero for an OTF entry.  The sign bit must be clear (-0.0f would come out as 2^129): the
// operands are exp(...) >= 0 and the telescope OTF, which finalize_otf_kernel keeps at >= +0.
__device__ __forceinline__ double f2d_bits(float f) {
    const unsigned b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}

// sign-agnostic test for zero on the integer pipe
__device__ __forceinline__ bool is_zero_bits(double t) {
    return ((__double2hiint(t) << 1) | __double2loint(t)) == 0;
}

}  // namespace psfr
