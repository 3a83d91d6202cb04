// Shared pass-kernel template: one warp per length-N line, loader/storer functors.
//
// N = NF * 1280.  The warp runs NF interleaved 1280-point transforms (sub-sequence s holds
// x[NF m + s]) and dumps each in natural order; for NF = 2 one combine step
//   X[k] = F0[k] + w_N^k F1[k],   X[k + 1280] = F0[k] - w_N^k F1[k]
// turns the two dumps into the two halves of the length-2560 spectrum in place.
#pragma once
#include "psfr_internal.h"
#include "warp_fft.cuh"
#include "tma.cuh"

namespace psfr {

using G = FftGeom<kR3>;
constexpr int kPassWarps = 4;
template <int NF>
constexpr size_t pass_smem() {
    return (size_t)(G::TW1 + G::TW2) * sizeof(double2) + (size_t)kPassWarps * NF * 2 * G::XBUF * sizeof(double);
}

// index of register slot i in the load layout of the 1280-point warp transform
__device__ __forceinline__ int slot_n(int i, int lane) { return (i & 7) * (kNB / 8) + lane + 32 * (i >> 3); }
// element of the length-N line that slot i of sub-sequence `sub` holds
template <int NF>
__device__ __forceinline__ int slot_e(int i, int lane, int sub) { return NF * slot_n(i, lane) + sub; }

// all storers start from the natural-order dumps: output k of the length-N spectrum sits in
// block k / 1280 (re in [0, XBUF), im in [XBUF, 2 XBUF) of that block)
template <int NF>
__device__ __forceinline__ double2 nat_get(const double* xb, int k) {
    const int q = (NF == 1) ? 0 : k / kNB, kk = (NF == 1) ? k : k % kNB;
    const double* b = xb + (size_t)q * 2 * G::XBUF;
    return make_double2(b[nat_addr(kk)], b[G::XBUF + nat_addr(kk)]);
}

template <int NF, class Loader, class Storer>
__global__ void __launch_bounds__(kPassWarps * 32)
pass_kernel(Loader ld, Storer st, int nfft, const double2* __restrict__ g_tw, const double2* __restrict__ g_twc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tw1 = reinterpret_cast<double2*>(smem_raw);
    double2* tw2 = tw1 + G::TW1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* xb = reinterpret_cast<double*>(tw2 + G::TW2) + (size_t)warp * NF * 2 * G::XBUF;
    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    __syncthreads();
    for (int f = blockIdx.x * kPassWarps + warp; f < nfft; f += gridDim.x * kPassWarps) {
#pragma unroll 1
        for (int sub = 0; sub < NF; ++sub) {
            double2 v[40];
            ld(f, lane, v, sub);
            // exchange scratch = the last dump region, which no earlier sub-sequence has filled
            warp_fft<kR3>(v, xb + (size_t)(2 * NF - 1) * G::XBUF, tw1, tw2, lane);
            __syncwarp();
            fft_dump<kR3>(v, xb + (size_t)sub * 2 * G::XBUF, lane, 0);
            fft_dump<kR3>(v, xb + (size_t)sub * 2 * G::XBUF + G::XBUF, lane, 1);
            __syncwarp();
        }
        if (NF == 2) {
            double* b0 = xb;
            double* b1 = xb + 2 * G::XBUF;
#pragma unroll 4
            for (int i = 0; i < 40; ++i) {
                const int k = lane + 32 * i, a = nat_addr(k);
                const double2 w = __ldg(g_twc + k);
                const double2 f0 = make_double2(b0[a], b0[G::XBUF + a]);
                const double2 f1 = cmul(make_double2(b1[a], b1[G::XBUF + a]), w);
                b0[a] = f0.x + f1.x;
                b0[G::XBUF + a] = f0.y + f1.y;
                b1[a] = f0.x - f1.x;
                b1[G::XBUF + a] = f0.y - f1.y;
            }
            __syncwarp();
        }
        st(f, lane, xb);
        __syncwarp();
    }
}

// Tile-staged variant (dim 1280): the input of a line is one contiguous block of global memory
// (`src.block(f, &bytes)`) that a TMA bulk copy brings into a warp-private shared-memory tile, so
// the pass streams its input at HBM speed instead of waiting on dependent per-lane loads.  Lines
// are dealt to the warps in a fixed stride.  The tile is dead as soon as `build` has turned it
// into the 40 register values; two ways to use that:
//   OVERLAP = true : the next tile is fetched while the warp transforms the current line; the
//                    natural-order dump needs its own two regions (re, im) -> 5 warps fit;
//   OVERLAP = false: the dead tile doubles as the imaginary half of the dump and the next tile
//                    is fetched after the store -> 8 (6) warps fit.  These kernels run one
//                    long dependent instruction stream per warp, so warps per scheduler count
//                    for more than the ~1.5 us of TMA latency per line that they now hide for
//                    each other.
template <int WARPS, int TILE_BYTES, bool OVERLAP>
constexpr size_t tiled_pass_smem() {
    return 128 + (size_t)(G::TW1 + G::TW2) * sizeof(double2) +
           (size_t)WARPS * (TILE_BYTES + (OVERLAP ? 2 : 1) * G::XBUF * sizeof(double));
}

template <int WARPS, int TILE_BYTES, bool OVERLAP, class Src, class Storer>
__global__ void __launch_bounds__(WARPS * 32, 1)
tiled_pass_kernel(Src src, Storer st, int nfft, const double2* __restrict__ g_tw) {
    static_assert(TILE_BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
    static_assert(OVERLAP || TILE_BYTES >= G::XBUF * sizeof(double), "the tile must hold one dump region");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    double2* tw1 = reinterpret_cast<double2*>(smem_raw + 128);
    double2* tw2 = tw1 + G::TW1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: [re dump = exchange scratch : XBUF][im dump : XBUF (OVERLAP only)][tile]; the storers
    // read im at xb + XBUF, which is the tile itself when it doubles as the im dump
    constexpr size_t kWarpBytes = TILE_BYTES + (OVERLAP ? 2 : 1) * G::XBUF * sizeof(double);
    unsigned char* mine = reinterpret_cast<unsigned char*>(tw2 + G::TW2) + (size_t)warp * kWarpBytes;
    double* xb = reinterpret_cast<double*>(mine);
    unsigned char* tile = mine + (OVERLAP ? 2 : 1) * G::XBUF * sizeof(double);
    uint64_t* bar = bars + warp;
    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
    auto fetch = [&](int f) {
        if (lane == 0) {
            uint32_t bytes;
            const void* g = src.block(f, &bytes);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, bytes);
            tma_load_1d(tile, g, bytes, bar);
        }
    };
    if (gw < nfft) fetch(gw);
    uint32_t phase = 0;
#pragma unroll 1
    for (int f = gw; f < nfft; f += nw) {
        double2 v[40];
        mbar_wait(bar, phase);
        phase ^= 1;
        src.build(f, lane, tile, v);
        __syncwarp();
        if (OVERLAP && f + nw < nfft) fetch(f + nw);
        warp_fft<kR3>(v, xb, tw1, tw2, lane);
        __syncwarp();
        fft_dump<kR3>(v, xb, lane, 0);
        fft_dump<kR3>(v, xb + G::XBUF, lane, 1);
        __syncwarp();
        st(f, lane, xb);
        __syncwarp();
        if (!OVERLAP && f + nw < nfft) fetch(f + nw);
    }
}

template <int WARPS, int TILE_BYTES, bool OVERLAP, class Src, class Storer>
static int launch_tiled_pass(Ctx* c, Src src, Storer st, int nfft, cudaStream_t s) {
    constexpr size_t smem = tiled_pass_smem<WARPS, TILE_BYTES, OVERLAP>();
    static_assert(smem <= 232448, "tiled pass shared memory exceeds the 227 KB per-CTA limit");
    if (int rc = ensure_dynamic_smem(c, tiled_pass_kernel<WARPS, TILE_BYTES, OVERLAP, Src, Storer>, smem)) return rc;
    int grid = (nfft + WARPS - 1) / WARPS;
    if (grid > c->sm_count) grid = c->sm_count;
    tiled_pass_kernel<WARPS, TILE_BYTES, OVERLAP, Src, Storer><<<grid, WARPS * 32, smem, s>>>(src, st, nfft, c->d_tw);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

template <int NF, class Loader, class Storer>
static int launch_pass(Ctx* c, Loader ld, Storer st, int nfft, cudaStream_t s) {
    if (int rc = ensure_dynamic_smem(c, pass_kernel<NF, Loader, Storer>, pass_smem<NF>())) return rc;
    int grid = (nfft + kPassWarps - 1) / kPassWarps;
    const int cap = c->sm_count * 8;
    if (grid > cap) grid = cap;
    pass_kernel<NF, Loader, Storer><<<grid, kPassWarps * 32, pass_smem<NF>(), s>>>(ld, st, nfft, c->d_tw, c->d_twc);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

}  // namespace psfr
