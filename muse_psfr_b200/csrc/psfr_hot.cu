// Stage B, pruned: the hot path of psf_muse / psd_to_psf (psfrec.py:667-685, 792-801).
//
// For every (plane, wavelength) the reference builds OTF = exp(-Dphi/2) * OTF_tel on the
// full N x N grid, inverse-transforms it and then reads only 80 rows x 80 columns of the
// result (bilinear resampling to 40 x 40).  Here:
//
//  * hot_rows_kernel  - persistent, one CTA per SM.  Row pairs of the structure function D
//    (and of the telescope OTF) stream into a shared-memory ring with TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx), issued by whichever warp releases a stage
//    last; eight warps each take one wavelength at a time: OTF rows = exp(-c_lambda D) * T
//    evaluated straight from shared memory into registers, one 1280-point warp FFT for the two packed real rows,
//    and only the 80 sampled frequencies (+ mirrors) are untangled and written, as one
//    32-byte sector per frequency.  D is read from HBM/L2 once per row pair for ALL
//    wavelengths; the N x N OTF and PSF grids never exist in memory.
//  * hot_cols pass    - 40 Hermitian column-pair transforms per PSF (summing the field
//    directions of a draw before the transform: the mean over directions, psfrec.py:674,
//    commutes with the linear transform), keeping the 80 sampled outputs -> 80x80 samples.
#include "pass_kernel.cuh"
#include "fast_exp.cuh"
#include "tma.cuh"

#ifndef PSFR_HOT_BLK
#define PSFR_HOT_BLK 4
#endif

namespace psfr {

// Launch shape per grid size: the ring stage holds two rows of D and two of the telescope OTF
// (4 N doubles), so dim 2560 affords two stages and four transform warps in 227 KB.
template <int NF>
struct HotCfg {
    static constexpr int Warps = NF == 1 ? 8 : 4;    // consumer warps
    static constexpr int Stages = NF == 1 ? 3 : 2;   // ring depth
    static constexpr int Tile = 2 * Dim<NF>::N;      // doubles per tile (two rows)
    static constexpr uint32_t TileBytes = Tile * sizeof(double);
    static constexpr size_t Smem = 128 + (size_t)(G::TW1 + G::TW2) * sizeof(double2) +
                                   (size_t)Stages * 2 * TileBytes + (size_t)Warps * G::XBUF * sizeof(double);
    static_assert(Smem <= 232448, "hot kernel shared memory exceeds the 227 KB per-CTA limit");
};

int hot_event(Ctx* c, int which, cudaStream_t s);   // psfr_api.cu: CUDA-event bracket of the row kernel

struct HotParams {
    const double* D;       // [nplanes][kRows][N]
    const double* T;       // [kRows][N]
    double2* Y;            // [nplanes][nlam][kNS][kRows]
    const double* clam;    // [nlam]
    const uint16_t* kidx;  // [nlam][kNS]
    const double2* wsamp;  // [nlam][2][kNS] NF = 2: w_N^k of the sampled outputs and of their mirrors
    const double* dmin;    // [nplanes][kRows] smallest D of each row (finalize_dphi_kernel)
    int* next_item;        // work counter, zeroed before the launch
    double cut;            // OTF entries with c*D > cut (exp < e^-cut) are flushed to zero
    int nplanes, nlam;
};

template <int NF>
__global__ void __launch_bounds__(HotCfg<NF>::Warps * 32, 1)
hot_rows_kernel(HotParams p, const double2* __restrict__ g_tw) {
    using D = Dim<NF>;
    using C = HotCfg<NF>;
    constexpr int kStages = C::Stages, kHotWarps = C::Warps, kTile = C::Tile, kN = D::N, kRows = D::Rows,
                  kPairs = D::Pairs;
    constexpr uint32_t kTileBytes = C::TileBytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    int* released = reinterpret_cast<int*>(full + kStages);   // per-stage count of warps done with it
    volatile int* item_of = released + kStages;               // per-stage work item (-1: no more work)
    double2* tw1 = reinterpret_cast<double2*>(smem_raw + 128);
    double2* tw2 = tw1 + G::TW1;
    double* ring = reinterpret_cast<double*>(tw2 + G::TW2);   // [stage][D tile | T tile]
    double* xall = ring + (size_t)kStages * 2 * kTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.nplanes * kPairs;

    // Work items (plane, row pair) are handed out by a global counter: with the underflow cut
    // the cost of an item ranges from "write zeros" to nlam full transforms, so a static
    // partition would leave most CTAs idle.  The fetching thread stages the two rows of D and
    // of the telescope OTF with TMA bulk loads; the item id travels through shared memory and
    // is published by the mbarrier phase (arrive has release, try_wait acquire semantics).
    auto issue = [&](int s) {
        const int item = atomicAdd(p.next_item, 1);
        if (item < items) {
            const int plane = item / kPairs, rp = item % kPairs;
            double* dst = ring + (size_t)s * 2 * kTile;
            item_of[s] = item;
            mbar_expect_tx(full + s, 2 * kTileBytes);
            tma_load_1d(dst, p.D + ((size_t)plane * kRows + 2 * rp) * kN, kTileBytes, full + s);
            tma_load_1d(dst + kTile, p.T + (size_t)(2 * rp) * kN, kTileBytes, full + s);
        } else {
            item_of[s] = -1;
            mbar_arrive(full + s);
        }
    };

    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            released[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < kStages; ++s) issue(s);
    }
    __syncthreads();

    double* xb = xall + (size_t)warp * G::XBUF;
#pragma unroll 1
    for (int it = 0;; ++it) {
        const int s = it % kStages, u = it / kStages;
        // every warp observes every fill, also when it has no wavelength in this item: that keeps
        // all warps within kStages items of each other, which the per-stage release counter and
        // the phase parity rely on.  Fills are issued in iteration order, so the first -1 a warp
        // sees is followed by -1 only and no TMA is in flight when the CTA retires.
        mbar_wait(full + s, u & 1);
        const int item = item_of[s];
        if (item < 0) break;
        // flat (item, wavelength) index g = it*nlam + lam is dealt round-robin to the warps
        int lam = (warp - (int)(((long long)it * p.nlam) % kHotWarps) + kHotWarps) % kHotWarps;
        if (lam < p.nlam) {
            const double* sD = ring + (size_t)s * 2 * kTile;
            const double* sT = sD + kTile;
            const int plane = item / kPairs, rp = item % kPairs;
            const double dm = fmin(__ldg(p.dmin + (size_t)plane * kRows + 2 * rp),
                                   __ldg(p.dmin + (size_t)plane * kRows + 2 * rp + 1));
#pragma unroll 1
            for (; lam < p.nlam; lam += kHotWarps) {
                const double cl = __ldg(p.clam + lam);
                double2* out = p.Y + ((size_t)plane * p.nlam + lam) * kNS * kRows + 2 * rp;
                if (cl * dm > p.cut) {
                    // both rows are below the cut everywhere: their transform is zero
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int y = lane + 32 * i;
                        if (y < kNS) {
                            double2* o = out + (size_t)y * kRows;
                            o[0] = make_double2(0.0, 0.0);
                            o[1] = make_double2(0.0, 0.0);
                        }
                    }
                    continue;
                }
                const double negc = -cl;
                const unsigned cut_hi = (unsigned)__double2hiint(-p.cut);
                // sampled frequencies kA and their mirrors kB = -kA (indices into the length-N spectrum)
                const uint16_t* kx = p.kidx + (size_t)lam * kNS;
                int ka[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) ka[i] = (lane + 32 * i < kNS) ? (int)__ldg(kx + lane + 32 * i) : 0;
                double2 za[3], zb[3];   // X[kA], X[kB] accumulated over the NF interleaved sub-sequences
#pragma unroll 1
                for (int sub = 0; sub < NF; ++sub) {
                    double2 v[40];
                    // kBlk slots (= 2 kBlk independent exp chains) per basic block
                    constexpr int kBlk = PSFR_HOT_BLK;
#pragma unroll
                    for (int i = 0; i < 40; i += kBlk) {
                        double tt[2 * kBlk], xx[2 * kBlk];
                        bool dead = true;
#pragma unroll
                        for (int q = 0; q < kBlk; ++q) {
                            const int n = slot_e<NF>(i + q, lane, sub);
                            tt[2 * q] = sT[n];
                            tt[2 * q + 1] = sT[kN + n];
                            xx[2 * q] = negc * sD[n];
                            xx[2 * q + 1] = negc * sD[kN + n];
                            // outside the pupil-autocorrelation support the OTF is exactly zero;
                            // below the underflow cut it is flushed to zero
                            dead = dead & (is_zero_bits(tt[2 * q]) | below_cut(xx[2 * q], cut_hi)) &
                                   (is_zero_bits(tt[2 * q + 1]) | below_cut(xx[2 * q + 1], cut_hi));
                        }
                        if (__all_sync(0xffffffffu, dead)) {
#pragma unroll
                            for (int q = 0; q < kBlk; ++q) v[i + q] = make_double2(0.0, 0.0);
                        } else {
                            double ee[2 * kBlk];
#pragma unroll
                            for (int q = 0; q < 2 * kBlk; ++q) ee[q] = fast_exp(xx[q]);
#pragma unroll
                            for (int q = 0; q < kBlk; ++q)
                                v[i + q] = make_double2(ee[2 * q] * tt[2 * q], ee[2 * q + 1] * tt[2 * q + 1]);
                        }
                    }
                    warp_fft<kR3>(v, xb, tw1, tw2, lane);
                    double2 fa[3], fb[3];
#pragma unroll
                    for (int cpt = 0; cpt < 2; ++cpt) {
                        fft_dump<kR3>(v, xb, lane, cpt);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            comp_set(fa[i], cpt, xb[nat_addr(ka[i] % kNB)]);
                            comp_set(fb[i], cpt, xb[nat_addr(((kN - ka[i]) % kN) % kNB)]);
                        }
                        __syncwarp();
                    }
                    if (NF == 1) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            za[i] = fa[i];
                            zb[i] = fb[i];
                        }
                    } else if (sub == 0) {
                        // park F0 in the output slots (L2-resident) instead of 12 more live registers
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const int y = lane + 32 * i;
                            if (y < kNS) {
                                double2* o = out + (size_t)y * kRows;
                                o[0] = fa[i];
                                o[1] = fb[i];
                            }
                        }
                    } else {
                        // X[k] = F0[k mod 1280] + w_N^k F1[k mod 1280]
                        const double2* ws = p.wsamp + (size_t)lam * 2 * kNS;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const int y = lane + 32 * i;
                            if (y < kNS) {
                                const double2* o = out + (size_t)y * kRows;
                                za[i] = cadd(o[0], cmul(fa[i], __ldg(ws + y)));
                                zb[i] = cadd(o[1], cmul(fb[i], __ldg(ws + kNS + y)));
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int y = lane + 32 * i;
                    if (y < kNS) {
                        double2* o = out + (size_t)y * kRows;
                        o[0] = make_double2(0.5 * (za[i].x + zb[i].x), 0.5 * (za[i].y - zb[i].y));
                        o[1] = make_double2(0.5 * (za[i].y + zb[i].y), 0.5 * (zb[i].x - za[i].x));
                    }
                }
            }
        }
        // release the stage; the last warp to do so refills it with the next work item
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int old = atomicAdd(released + s, 1);
            if (old == kHotWarps - 1) {
                atomicExch(released + s, 0);
                __threadfence_block();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(s);
            }
        }
    }
}

// ---------------- pruned column pass
// line f = ((draw*nlam + lam)*40 + m): sampled rows 2m, 2m+1 of that PSF (adjacent in Y, one
// contiguous 2*Rows*16-byte block), summed over the ndir planes of the draw, Hermitian-extended
// along the half-plane row index, transformed, and only the 80 sampled outputs kept:
//   S[img][2m + c][j] = scale * (-1)^(k_{2m+c} + k_j) * {Re, Im}(X[k_j]).
// Every warp owns a private tile that a TMA bulk copy fills while the warp transforms the
// previous line (the tile is dead as soon as its values sit in registers), so the kernel
// streams Y at HBM speed instead of waiting on 80 dependent 16-byte loads per lane.
template <int NF>
struct ColCfg {
    static constexpr int Warps = NF == 1 ? 6 : 4;
    static constexpr int TileElems = 2 * Dim<NF>::Rows;                 // double2 per tile
    static constexpr uint32_t TileBytes = TileElems * sizeof(double2);
    static constexpr size_t Smem = 128 + (size_t)(G::TW1 + G::TW2) * sizeof(double2) +
                                   (size_t)Warps * (TileBytes + G::XBUF * sizeof(double));
    static_assert(TileBytes % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
    static_assert(Smem <= 232448, "column kernel shared memory exceeds the 227 KB per-CTA limit");
};

struct ColParams {
    const double2* Y;      // [nplanes][nlam][kNS][Rows]
    double* S;             // [nimg][kNS][kNS]
    const uint16_t* kidx;  // [nlam][kNS]
    const double2* wsamp;  // [nlam][2][kNS] (NF = 2)
    int nlines, nlam, ndir;
    double scale;
};

template <int NF>
__global__ void __launch_bounds__(ColCfg<NF>::Warps * 32, 1)
hot_cols_kernel(ColParams p, const double2* __restrict__ g_tw) {
    using D = Dim<NF>;
    using C = ColCfg<NF>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);              // one mbarrier per warp
    double2* tw1 = reinterpret_cast<double2*>(smem_raw + 128);
    double2* tw2 = tw1 + G::TW1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* tile = tw2 + G::TW2 + (size_t)warp * C::TileElems;
    double* xb = reinterpret_cast<double*>(tw2 + G::TW2 + (size_t)C::Warps * C::TileElems) + (size_t)warp * G::XBUF;
    uint64_t* bar = bars + warp;

    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int gw = blockIdx.x * C::Warps + warp, nw = gridDim.x * C::Warps;
    // tile of (line f, direction d)
    auto src_of = [&](int f, int d) {
        const int m = f % (kNS / 2), img = f / (kNS / 2);
        const int lam = img % p.nlam, draw = img / p.nlam;
        return p.Y + (((size_t)(draw * p.ndir + d) * p.nlam + lam) * kNS + 2 * m) * D::Rows;
    };
    auto fetch = [&](int f, int d) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, C::TileBytes);
            tma_load_1d(tile, src_of(f, d), C::TileBytes, bar);
        }
    };
    if (gw < p.nlines) fetch(gw, 0);
    uint32_t phase = 0;
#pragma unroll 1
    for (int f = gw; f < p.nlines; f += nw) {
        const int m = f % (kNS / 2), img = f / (kNS / 2), lam = img % p.nlam;
        const uint16_t* kx = p.kidx + (size_t)lam * kNS;
        int kj[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) kj[i] = (lane + 32 * i < kNS) ? (int)__ldg(kx + lane + 32 * i) : 0;
        double2 z[3];
#pragma unroll 1
        for (int sub = 0; sub < NF; ++sub) {
            double2 v[40];
#pragma unroll
            for (int i = 0; i < 40; ++i) v[i] = make_double2(0.0, 0.0);
#pragma unroll 1
            for (int d = 0; d < p.ndir; ++d) {
                // NF = 2 walks the tiles of the line twice (even, then odd elements)
                mbar_wait(bar, phase);
                phase ^= 1;
                const double2* c1 = tile;
                const double2* c2 = tile + D::Rows;
#pragma unroll
                for (int i = 0; i < 40; ++i) {
                    const int n = slot_e<NF>(i, lane, sub);
                    if (n <= D::NH) {
                        const double2 r1 = c1[n], r2 = c2[n];
                        v[i].x += r1.x - r2.y;
                        v[i].y += r1.y + r2.x;
                    } else {
                        const double2 r1 = c1[D::N - n], r2 = c2[D::N - n];
                        v[i].x += r1.x + r2.y;
                        v[i].y += r2.x - r1.y;
                    }
                }
                __syncwarp();
                // the tile is dead: prefetch the next one of this warp's sequence
                if (d + 1 < p.ndir) fetch(f, d + 1);
                else if (sub + 1 < NF) fetch(f, 0);
                else if (f + nw < p.nlines) fetch(f + nw, 0);
            }
            warp_fft<kR3>(v, xb, tw1, tw2, lane);
            double2 fz[3];
#pragma unroll
            for (int cpt = 0; cpt < 2; ++cpt) {
                fft_dump<kR3>(v, xb, lane, cpt);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 3; ++i) comp_set(fz[i], cpt, xb[nat_addr(kj[i] % kNB)]);
                __syncwarp();
            }
            if (NF == 1 || sub == 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) z[i] = fz[i];
            } else {
                const double2* ws = p.wsamp + (size_t)lam * 2 * kNS;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (lane + 32 * i < kNS) z[i] = cadd(z[i], cmul(fz[i], __ldg(ws + lane + 32 * i)));
            }
        }
        const int k1 = __ldg(kx + 2 * m), k2 = __ldg(kx + 2 * m + 1);
        double* o = p.S + ((size_t)img * kNS + 2 * m) * kNS;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int j = lane + 32 * i;
            if (j < kNS) {
                o[j] = (((k1 + kj[i]) & 1) ? -p.scale : p.scale) * z[i].x;
                o[kNS + j] = (((k2 + kj[i]) & 1) ? -p.scale : p.scale) * z[i].y;
            }
        }
    }
}

template <int NF>
static int pruned_psf_t(Ctx* c, int ndraw, int ndir, int nlam, cudaStream_t s) {
    using D = Dim<NF>;
    using C = HotCfg<NF>;
    if (int rc = ensure_dynamic_smem(c, hot_rows_kernel<NF>, C::Smem)) return rc;
    const int nplanes = ndraw * ndir;
    HotParams p{c->d_dphi, c->d_otf, c->d_ybuf, c->d_lam, c->d_kidx, c->d_wsamp, c->d_dmin, c->d_counter,
                c->exp_cut, nplanes, nlam};
    int grid = c->sm_count;
    if (grid > nplanes * D::Pairs) grid = nplanes * D::Pairs;
    PSFR_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int), s));
    int rc = hot_event(c, 0, s);
    if (rc) return rc;
    hot_rows_kernel<NF><<<grid, C::Warps * 32, C::Smem, s>>>(p, c->d_tw);
    PSFR_LAUNCH_CHECK(c);
    if ((rc = hot_event(c, 1, s))) return rc;
    c->hot_launches += 1;
    c->hot_psfs += (long long)nplanes * nlam;
    // psd_to_psf divides by the PSF sum (= T centre = 1/N^2, cancelling the 1/N^2 of the
    // inverse transform); psf_muse averages the directions.
    if ((rc = ensure_dynamic_smem(c, hot_cols_kernel<NF>, ColCfg<NF>::Smem))) return rc;
    ColParams q{c->d_ybuf, c->d_samp, c->d_kidx, c->d_wsamp, ndraw * nlam * (kNS / 2), nlam, ndir, 1.0 / ndir};
    int cgrid = (q.nlines + ColCfg<NF>::Warps - 1) / ColCfg<NF>::Warps;
    if (cgrid > c->sm_count) cgrid = c->sm_count;
    hot_cols_kernel<NF><<<cgrid, ColCfg<NF>::Warps * 32, ColCfg<NF>::Smem, s>>>(q, c->d_tw);
    PSFR_LAUNCH_CHECK(c);
    return PSFR_OK;
}

int run_pruned_psf(Ctx* c, int ndraw, int ndir, int nlam, cudaStream_t s) {
    return c->NF == 1 ? pruned_psf_t<1>(c, ndraw, ndir, nlam, s) : pruned_psf_t<2>(c, ndraw, ndir, nlam, s);
}

}  // namespace psfr
