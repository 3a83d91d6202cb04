// Shared pass-kernel template: one warp per length-N line, loader/storer functors.
#pragma once
#include "psfr_internal.h"
#include "warp_fft.cuh"

namespace psfr {

using G = FftGeom<kR3>;
constexpr int kPassWarps = 4;
constexpr size_t kPassSmem = (size_t)(G::TW1 + G::TW2) * sizeof(double2) +
                             (size_t)kPassWarps * 2 * G::XBUF * sizeof(double);

// index of register slot i in the load layout
__device__ __forceinline__ int slot_n(int i, int lane) { return (i & 7) * (kN / 8) + lane + 32 * (i >> 3); }

// all storers start from the natural-order dump: re in xb[0..), im in xb[XBUF..)
__device__ __forceinline__ double2 nat_get(const double* xb, int k) {
    return make_double2(xb[nat_addr(k)], xb[G::XBUF + nat_addr(k)]);
}


template <class Loader, class Storer>
__global__ void __launch_bounds__(kPassWarps * 32)
pass_kernel(Loader ld, Storer st, int nfft, const double2* __restrict__ g_tw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tw1 = reinterpret_cast<double2*>(smem_raw);
    double2* tw2 = tw1 + G::TW1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* xb = reinterpret_cast<double*>(tw2 + G::TW2) + (size_t)warp * 2 * G::XBUF;
    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    __syncthreads();
    for (int f = blockIdx.x * kPassWarps + warp; f < nfft; f += gridDim.x * kPassWarps) {
        double2 v[40];
        ld(f, lane, v);
        warp_fft<kR3>(v, xb, tw1, tw2, lane);
        fft_dump<kR3>(v, xb, lane, 0);
        fft_dump<kR3>(v, xb + G::XBUF, lane, 1);
        __syncwarp();
        st(f, lane, xb);
        __syncwarp();
    }
}

template <class Loader, class Storer>
static int launch_pass(Ctx* c, Loader ld, Storer st, int nfft, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        PSFR_CUDA(c, cudaFuncSetAttribute(pass_kernel<Loader, Storer>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassSmem));
        attr_set = true;
    }
    int grid = (nfft + kPassWarps - 1) / kPassWarps;
    const int cap = c->sm_count * 8;
    if (grid > cap) grid = cap;
    pass_kernel<Loader, Storer><<<grid, kPassWarps * 32, kPassSmem, s>>>(ld, st, nfft, c->d_tw);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

}  // namespace psfr
