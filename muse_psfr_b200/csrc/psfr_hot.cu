// Stage B, pruned: the hot path of psf_muse / psd_to_psf (psfrec.py:667-685, 792-801).
//
// For every (plane, wavelength) the reference builds OTF = exp(-Dphi/2) * OTF_tel on the
// full N x N grid, inverse-transforms it and then reads only 80 rows x 80 columns of the
// result (bilinear resampling to 40 x 40).  Here:
//
//  * row pass - group_rows_kernel (psfr_hot2.cu, the default) or hot_rows_kernel below
//    (PSFR_OPT_ROW_KERNEL = 1): persistent, one CTA per SM.  Row pairs of the structure function D
//    and of the telescope OTF (FP64 and FP32 copies) stream into a shared-memory ring with TMA
//    bulk copies (cp.async.bulk + mbarrier complete_tx), issued by whichever warp releases a
//    stage last.  A unit = one row pair at one wavelength: OTF rows = exp(-c_lambda D) * T
//    evaluated straight from shared memory into registers (graded precision),
//    one 1280-point warp FFT for the two packed real rows, and only the 40 kept frequencies
//    (+ mirrors) are untangled and written, as one 32-byte sector per frequency.
//    The eight warps run their units in lockstep rounds so that they share instruction fetches.
//    D is read from HBM/L2 once per row pair for ALL wavelengths; the N x N OTF and PSF grids
//    never exist in memory.
//  * hot_cols pass    - 20 Hermitian column-pair transforms per PSF (summing the field
//    directions of a draw before the transform: the mean over directions, psfrec.py:674,
//    commutes with the linear transform), keeping the 80 sampled outputs and their mirrors ->
//    80x80 samples (the mirrored outputs fill the sample rows of the frequencies -kc).
#include "pass_kernel.cuh"
#include "fast_exp.cuh"
#include "tma.cuh"

// warps per lockstep group of the row kernel (see hot_rows_kernel)
#ifndef PSFR_HOT_GROUP
#define PSFR_HOT_GROUP 8
#endif

namespace psfr {

// Launch shape per grid size.  dim 1280: a ring stage holds two rows of D and of the telescope
// OTF in double AND in single precision (6 N doubles = 60 KB); two stages, eight transform warps.
// The single-precision copies feed the block grading and the FP32 row pairs without any
// double -> float conversion in the kernel (F2F issues at a quarter of the DFMA rate).
// dim 2560: FP64 only (4 N doubles = 80 KB per stage), two stages, four warps, no grading.
template <int NF>
struct HotCfg {
    static constexpr bool F32 = NF == 1;             // single-precision copies staged, grading on
    static constexpr int Warps = NF == 1 ? 8 : 4;    // consumer warps
    static constexpr int Stages = 2;                 // ring depth
    static constexpr int Tile = 2 * Dim<NF>::N;      // elements per tile (two rows)
    static constexpr uint32_t TileBytes = Tile * sizeof(double);
    static constexpr uint32_t TileBytes32 = Tile * sizeof(float);
    static constexpr uint32_t StageBytes = 2 * TileBytes + (F32 ? 2 * TileBytes32 : 0);
    static constexpr int TabMax = 64;                // wavelength tables up to this size are staged in shared memory
    static constexpr size_t TabBytes = TabMax * (2 * sizeof(double) + sizeof(int));
    static constexpr size_t Smem = 128 + (size_t)(G::TW1 + G::TW2) * sizeof(double2) +
                                   (size_t)Stages * StageBytes + (size_t)Warps * G::XBUF * sizeof(double) + TabBytes;
    static_assert(Smem <= 232448, "hot kernel shared memory exceeds the 227 KB per-CTA limit");
};

int hot_event(Ctx* c, int which, cudaStream_t s);   // psfr_api.cu: CUDA-event bracket of the row kernel

struct HotParams {
    const double* D;       // [nplanes][kRows][N]
    const double* T;       // [kRows][N]
    double2* Y;            // [nplanes][nlam][kNC][kRows]
    const ushort2* kaddr;  // [nlam][kNC] where X[k], X[-k] of the kept frequencies sit in the natural-order dump
    const double2* wsamp;  // [nlam][2][kNC] NF = 2: w_N^k of the kept frequencies and of their mirrors
    const double* dmin;    // [nplanes][kRows] smallest D of each row (StoreDphi)
    const float* D32;      // single-precision copies of D and T (dim 1280)
    const float* T32;
    const float2* tw32;    // single-precision twiddles (global memory, L1-resident)
    const double* csort;   // [nlam] c_lambda in descending order
    const int* lorder;     // [nlam] wavelength index of sorted position i
    int* next_item;        // work counter, zeroed before the launch
    double cut;            // OTF entries with c*D > cut (exp < e^-cut) are flushed to zero
    double grade;          // blocks whose live entries all have c*D >= grade take the single-precision exp
    double f32_min;        // row pairs with c*min(D) >= f32_min run entirely in single precision (NF = 1)
    int nplanes, nlam;
};

template <int NF>
__global__ void __launch_bounds__(HotCfg<NF>::Warps * 32, 1)
hot_rows_kernel(HotParams p, const double2* __restrict__ g_tw) {
    using D = Dim<NF>;
    using C = HotCfg<NF>;
    constexpr int kStages = C::Stages, kHotWarps = C::Warps, kTile = C::Tile, kN = D::N, kRows = D::Rows,
                  kPairs = D::Pairs;
    constexpr uint32_t kTileBytes = C::TileBytes, kTileBytes32 = C::TileBytes32;
    constexpr int kQ = (kNC + 31) / 32;   // kept frequencies per lane
    constexpr size_t kStageDoubles = C::StageBytes / sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    int* released = reinterpret_cast<int*>(full + kStages);   // per-stage count of warps done with it
    volatile int* item_of = released + kStages;               // per-stage work item (-1: no more work)
    volatile int* la_of = item_of + kStages;                  // per-stage: sorted positions [0, la) are dead,
    volatile int* lb_of = la_of + kStages;                    //   [la, lb) single precision, [lb, nlam) FP64
    volatile int* ns_of = lb_of + kStages;                    // per-stage number of stream slots of the item
    static_assert(kStages == 2, "the 128-byte header is laid out for two stages");
    double2* tw1 = reinterpret_cast<double2*>(smem_raw + 128);
    double2* tw2 = tw1 + G::TW1;
    double* ring = reinterpret_cast<double*>(tw2 + G::TW2);   // [stage][D | T | D32 | T32]
    double* xall = ring + (size_t)kStages * kStageDoubles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.nplanes * kPairs;
    // per-wavelength scalars in sorted order (c, 1/c, wavelength index): every unit starts with
    // them, and in lockstep nobody hides a trip to L2 - staged here when they fit
    double* tab_c = xall + (size_t)kHotWarps * G::XBUF;
    double* tab_rc = tab_c + C::TabMax;
    int* tab_lo = reinterpret_cast<int*>(tab_rc + C::TabMax);
    const bool tabbed = p.nlam <= C::TabMax;
    auto c_of = [&](int pos) { return tabbed ? tab_c[pos] : __ldg(p.csort + pos); };

    // Work items (plane, row pair) are handed out by a global counter: with the underflow cut
    // the cost of an item ranges from "write zeros" to nlam full transforms, so a static
    // partition would leave most CTAs idle.  The fetching thread stages the two rows of D and
    // of the telescope OTF with TMA bulk loads; the item id travels through shared memory and
    // is published by the mbarrier phase (arrive has release, try_wait acquire semantics).
    auto issue = [&](int s) {
        const int item = atomicAdd(p.next_item, 1);
        if (item < items) {
            const int plane = item / kPairs, rp = item % kPairs;
            double* dst = ring + (size_t)s * kStageDoubles;
            item_of[s] = item;
            const double dm = fmin(__ldg(p.dmin + (size_t)plane * kRows + 2 * rp),
                                   __ldg(p.dmin + (size_t)plane * kRows + 2 * rp + 1));
            // Classes of the item's nlam units, by binary search on the descending c_lambda:
            // c * min(D) > cut -> dead (the transform of both rows is zero), >= f32_min ->
            // single precision, else FP64.  Both predicates are monotone along the sorted order.
            int lo = 0, hi = p.nlam;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (c_of(mid) * dm > p.cut) lo = mid + 1; else hi = mid;
            }
            const int la = lo;
            hi = p.nlam;
            while (C::F32 && lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (c_of(mid) * dm >= p.f32_min) lo = mid + 1; else hi = mid;
            }
            la_of[s] = la;
            lb_of[s] = lo;
            // stream slots: one FP64 unit, or TWO single-precision units that run as the two
            // halves of one packed transform
            ns_of[s] = (lo - la + 1) / 2 + (p.nlam - lo);
            // dead at every wavelength: the consumers only write zeros and never look at the stage
            if (la == p.nlam) {
                mbar_arrive(full + s);
                return;
            }
            const size_t off = ((size_t)plane * kRows + 2 * rp) * kN, offT = (size_t)(2 * rp) * kN;
            mbar_expect_tx(full + s, C::StageBytes);
            tma_load_1d(dst, p.D + off, kTileBytes, full + s);
            tma_load_1d(dst + kTile, p.T + offT, kTileBytes, full + s);
            if (C::F32) {
                float* dst32 = reinterpret_cast<float*>(dst + 2 * kTile);
                tma_load_1d(dst32, p.D32 + off, kTileBytes32, full + s);
                tma_load_1d(dst32 + kTile, p.T32 + offT, kTileBytes32, full + s);
            }
        } else {
            item_of[s] = -1;
            mbar_arrive(full + s);
        }
    };

    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    if (tabbed)
        for (int i = threadIdx.x; i < p.nlam; i += blockDim.x) {
            const double cv = __ldg(p.csort + i);
            tab_c[i] = cv;
            tab_rc[i] = 1.0 / cv;
            tab_lo[i] = __ldg(p.lorder + i);
        }
    __syncthreads();   // the first issue() below reads the table
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
    // The LIVE units of this CTA form a stream: item after item in the order in which the CTA
    // receives them (seq = 0, 1, ...), within an item the sorted positions la .. nlam-1
    // (reversed for odd seq, so that the class changes only once per item along the stream).
    // The stream is dealt to the warps in ROUNDS of eight consecutive units and all warps
    // enter a round together (one CTA barrier): the unit body is ~70 KB of straight-line code
    // (+ 40 KB for the single-precision class), far more than the instruction caches hold;
    // warps that run it side by side share every fetched line, warps that drift apart each
    // stream it from L2 on their own (ncu: `no_instruction` was the top stall of the
    // free-running version).  Dead units are not part of the stream: whoever passes an item
    // zeroes its share of them.
    //
    // A warp passes through EVERY item in order - wait for its fill, release it when its next
    // unit lies in a later item - also when it has no unit in it: that keeps all warps within
    // kStages items of each other, which the per-stage release counter and the phase parity
    // rely on.  Fills are issued in order, so the first -1 is followed by -1 only and no TMA
    // is in flight when the CTA retires.
    int cur = 0;            // sequence number of the item this warp is in
    bool seen = false;      // its fill has been observed (and this warp's dead units zeroed)
    int base = 0;           // stream offset of the round's first unit, relative to item cur's first live unit
    // move to the item that holds stream offset `rel` (relative to item cur); false: the stream ended
    auto seek = [&](int& rel) -> bool {
        while (true) {
            const int s = cur % kStages;
            if (!seen) {
                mbar_wait(full + s, (cur / kStages) & 1);
                seen = true;
                const int item = item_of[s];
                if (item >= 0) {
                    const int plane = item / kPairs, rp = item % kPairs, la = la_of[s];
                    for (int i = warp; i < la; i += kHotWarps) {
                        double2* out = p.Y + ((size_t)plane * p.nlam + (tabbed ? tab_lo[i] : __ldg(p.lorder + i))) * kNC * kRows + 2 * rp;
#pragma unroll
                        for (int k = 0; k < kQ; ++k) {
                            const int y = lane + 32 * k;
                            if (y < kNC) {
                                st_global_256(out + (size_t)y * kRows, make_double2(0.0, 0.0), make_double2(0.0, 0.0));
                            }
                        }
                    }
                }
            }
            if (item_of[s] < 0) return false;
            const int n = ns_of[s];
            if (rel < n) return true;
            rel -= n;
            base -= n;
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
            __syncwarp();
            ++cur;
            seen = false;
        }
    };
    constexpr int kGroup = PSFR_HOT_GROUP < kHotWarps ? PSFR_HOT_GROUP : kHotWarps;
    base = (warp / kGroup) * kGroup;   // stream offset of the group's first unit of the round
    int rel0 = base;
    bool more = seek(rel0);
#pragma unroll 1
    for (;;) {
        // one barrier per lockstep group (kGroup consecutive warps = one warp per scheduler)
        asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / kGroup), "n"(kGroup * 32) : "memory");
        if (!more) break;   // the group's first unit decides for the group: the stream only gets later
        int rel = base + warp % kGroup;
        if (seek(rel)) {
            const int s = cur % kStages;
            const int item = item_of[s];
            // slot -> units (sorted positions): the first npair slots of an item hold two
            // single-precision units each (the last one may hold one), the others one FP64 unit;
            // odd items run their slots backwards
            const int la = la_of[s], lb = lb_of[s], npair = (lb - la + 1) / 2;
            const int slot = (cur & 1) ? ns_of[s] - 1 - rel : rel;
            const double* sD = ring + (size_t)s * kStageDoubles;
            const double* sT = sD + kTile;
            const float* sD32 = reinterpret_cast<const float*>(sD + 2 * kTile);   // dim 1280 only
            const float* sT32 = sD32 + kTile;
            const int plane = item / kPairs, rp = item % kPairs;
            auto lam_of = [&](int pos) { return tabbed ? tab_lo[pos] : __ldg(p.lorder + pos); };
            auto out_of = [&](int lam) { return p.Y + ((size_t)plane * p.nlam + lam) * kNC * kRows + 2 * rp; };
            // untangle the two packed real rows: row 2rp from the even, row 2rp+1 from the odd part
            auto store_rows = [&](double2* out, int i, double2 za, double2 zb) {
                if (lane + 32 * i < kNC)
                    st_global_256(out + (size_t)(lane + 32 * i) * kRows,
                                  make_double2(0.5 * (za.x + zb.x), 0.5 * (za.y - zb.y)),
                                  make_double2(0.5 * (za.y + zb.y), 0.5 * (zb.x - za.x)));
            };
            if (C::F32 && slot < npair) {
                // ---- single-precision pair: every entry of both rows is below exp(-f32_min) (default
                // e^-25 = 1.4e-11) of the OTF peak at BOTH wavelengths, so a relative error of 1e-6 in
                // their contribution is < 1e-17 of the peak.  exp (MUFU ex2), the products and the
                // transform run in FP32 from the FP32 copies of D and T, and the two wavelengths are
                // the two halves of ONE packed transform (Z2: FADD2 / FMUL2 / FFMA2).
                const int posA = la + 2 * slot;
                const bool two = posA + 1 < lb;
                const int posB = two ? posA + 1 : posA;
                const int lamA = lam_of(posA), lamB = lam_of(posB);
                const double cB = c_of(posB);
                const float nA = (float)(-c_of(posA) * 1.44269504088896338700);   // exp(-c D) = 2^(n D)
                const float nB = (float)(-cB * 1.44269504088896338700);
                // c_B <= c_A: an entry below the cut at B is below it at A
                const int cut32 = __float_as_int((float)(p.cut * (tabbed ? tab_rc[posB] : 1.0 / cB)));
                ushort2 kaA[kQ], kaB[kQ];   // dump addresses of X[k], X[-k] for this lane's sampled frequencies
#pragma unroll
                for (int i = 0; i < kQ; ++i) {
                    const bool in = lane + 32 * i < kNC;
                    kaA[i] = in ? __ldg(p.kaddr + (size_t)lamA * kNC + lane + 32 * i) : make_ushort2(0, 0);
                    kaB[i] = in ? __ldg(p.kaddr + (size_t)lamB * kNC + lane + 32 * i) : make_ushort2(0, 0);
                }
                Z2 vf[40];
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) {
                    float df[10], tf[10];
                    bool dead = true;
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const int n = slot_e<NF>(j * 8 + n1, lane, 0);
                        df[2 * j] = sD32[n];
                        df[2 * j + 1] = sD32[kN + n];
                        tf[2 * j] = sT32[n];
                        tf[2 * j + 1] = sT32[kN + n];
                        dead = dead & ((tf[2 * j] == 0.f) | (__float_as_int(df[2 * j]) >= cut32)) &
                               ((tf[2 * j + 1] == 0.f) | (__float_as_int(df[2 * j + 1]) >= cut32));
                    }
                    if (__all_sync(0xffffffffu, dead)) {
#pragma unroll
                        for (int j = 0; j < 5; ++j) {
                            vf[j * 8 + n1].x = F2(0.f, 0.f);
                            vf[j * 8 + n1].y = F2(0.f, 0.f);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 5; ++j) {
                            vf[j * 8 + n1].x = F2(ex2_approx(nA * df[2 * j]) * tf[2 * j], ex2_approx(nB * df[2 * j]) * tf[2 * j]);
                            vf[j * 8 + n1].y = F2(ex2_approx(nA * df[2 * j + 1]) * tf[2 * j + 1],
                                                  ex2_approx(nB * df[2 * j + 1]) * tf[2 * j + 1]);
                        }
                    }
                }
                warp_fft<kR3>(vf, xb, p.tw32, p.tw32 + G::TW1, lane);
                float2* xf = reinterpret_cast<float2*>(xb);
                // natural-order dump of one component (both wavelengths) per round; wavelength A reads
                // the first halves at its own sampled frequencies, B the second halves at its own
                float2 fa[kQ], fb[kQ], ga[kQ], gb[kQ];   // A: X[kA], X[-kA]; B likewise
#pragma unroll
                for (int cpt = 0; cpt < 2; ++cpt) {
                    fft_dump<kR3>(vf, xf, lane, cpt);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < kQ; ++i) {
                        const float a = xf[kaA[i].x].x, am = xf[kaA[i].y].x;
                        const float b = xf[kaB[i].x].y, bm = xf[kaB[i].y].y;
                        if (cpt == 0) {
                            fa[i].x = a; fb[i].x = am; ga[i].x = b; gb[i].x = bm;
                        } else {
                            fa[i].y = a; fb[i].y = am; ga[i].y = b; gb[i].y = bm;
                        }
                    }
                    __syncwarp();
                }
                double2* outA = out_of(lamA);
                double2* outB = out_of(lamB);
#pragma unroll
                for (int i = 0; i < kQ; ++i) {
                    store_rows(outA, i, make_double2((double)fa[i].x, (double)fa[i].y),
                               make_double2((double)fb[i].x, (double)fb[i].y));
                    if (two)
                        store_rows(outB, i, make_double2((double)ga[i].x, (double)ga[i].y),
                                   make_double2((double)gb[i].x, (double)gb[i].y));
                }
            } else {
                const int pos = lb + (slot - npair);   // without the single-precision class lb = la, npair = 0
                const int lam = lam_of(pos);
                const double cl = c_of(pos), rcl = tabbed ? tab_rc[pos] : 1.0 / cl;
                double2* out = out_of(lam);
                const double negc = -cl;
                // where the sampled frequencies kA and their mirrors kB = -kA sit in the natural-order dump
                const ushort2* kx = p.kaddr + (size_t)lam * kNC;
                ushort2 ka[kQ];
#pragma unroll
                for (int i = 0; i < kQ; ++i) ka[i] = (lane + 32 * i < kNC) ? __ldg(kx + lane + 32 * i) : make_ushort2(0, 0);
                double2 za[kQ], zb[kQ];   // X[kA], X[kB] accumulated over the NF interleaved sub-sequences
                // Cut and grade thresholds on D itself, tested on the integer pipe: a non-negative
                // float (double) orders like its bit pattern (high word); the SIGNED compare keeps a
                // D rounded slightly below zero alive.
                const float negc2f = (float)(negc * 1.44269504088896338700);   // exp(-c D) = 2^(negc2f D)
                const int cut32 = __float_as_int((float)(p.cut * rcl));
                const int grade32 = __float_as_int((float)(p.grade * rcl));
#pragma unroll 1
                for (int sub = 0; sub < NF; ++sub) {
                    double2 v[40];
                    // One block = the 160 contiguous cells n1*160 .. n1*160+159 of both rows (slots
                    // j*8 + n1, j = 0..4): ten independent exp chains per lane.  Three grades,
                    // chosen per block by a warp vote on the FP32 copies: dead (outside the pupil-
                    // autocorrelation support, where the OTF is exactly zero, or below the underflow
                    // cut) -> zeros; every live entry below exp(-grade) (default e^-20 = 2.1e-9 of
                    // the peak) -> MUFU ex2 and the product in single precision, absolute error
                    // < 1e-14 of the peak; else FP64 values from the ring and the full FP64 exp.
#pragma unroll
                    for (int n1 = 0; n1 < 8; ++n1) {
                        bool dead = true, cheap = true;
                        float df[10], tf[10];
                        double dd[10], tt[10];
                        if (C::F32) {
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                const int n = slot_e<NF>(j * 8 + n1, lane, sub);
                                df[2 * j] = sD32[n];
                                df[2 * j + 1] = sD32[kN + n];
                                tf[2 * j] = sT32[n];
                                tf[2 * j + 1] = sT32[kN + n];
#pragma unroll
                                for (int r = 0; r < 2; ++r) {
                                    const bool z = tf[2 * j + r] == 0.f;
                                    const int h = __float_as_int(df[2 * j + r]);
                                    dead = dead & (z | (h >= cut32));
                                    cheap = cheap & (z | (h >= grade32));
                                }
                            }
                        } else {
                            const int cut_hi = __double2hiint(p.cut * rcl);
                            cheap = false;
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                const int n = slot_e<NF>(j * 8 + n1, lane, sub);
                                dd[2 * j] = sD[n];
                                dd[2 * j + 1] = sD[kN + n];
                                tt[2 * j] = sT[n];
                                tt[2 * j + 1] = sT[kN + n];
                                dead = dead & (is_zero_bits(tt[2 * j]) | (__double2hiint(dd[2 * j]) >= cut_hi)) &
                                       (is_zero_bits(tt[2 * j + 1]) | (__double2hiint(dd[2 * j + 1]) >= cut_hi));
                            }
                        }
                        if (__all_sync(0xffffffffu, dead)) {
#pragma unroll
                            for (int j = 0; j < 5; ++j) v[j * 8 + n1] = make_double2(0.0, 0.0);
                        } else if (C::F32 && __all_sync(0xffffffffu, cheap)) {
#pragma unroll
                            for (int j = 0; j < 5; ++j)
                                v[j * 8 + n1] = make_double2(f2d_bits(ex2_approx(negc2f * df[2 * j]) * tf[2 * j]),
                                                             f2d_bits(ex2_approx(negc2f * df[2 * j + 1]) * tf[2 * j + 1]));
                        } else {
                            if (C::F32) {
#pragma unroll
                                for (int j = 0; j < 5; ++j) {
                                    const int n = slot_e<NF>(j * 8 + n1, lane, sub);
                                    dd[2 * j] = sD[n];
                                    dd[2 * j + 1] = sD[kN + n];
                                    tt[2 * j] = sT[n];
                                    tt[2 * j + 1] = sT[kN + n];
                                }
                            }
                            double ee[10];
#pragma unroll
                            for (int q = 0; q < 10; ++q) ee[q] = fast_exp(negc * dd[q]);
#pragma unroll
                            for (int j = 0; j < 5; ++j)
                                v[j * 8 + n1] = make_double2(ee[2 * j] * tt[2 * j], ee[2 * j + 1] * tt[2 * j + 1]);
                        }
                    }
                    warp_fft<kR3>(v, xb, tw1, tw2, lane);
                    double2 fa[kQ], fb[kQ];
#pragma unroll
                    for (int cpt = 0; cpt < 2; ++cpt) {
                        fft_dump<kR3>(v, xb, lane, cpt);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < kQ; ++i) {
                            comp_set(fa[i], cpt, xb[ka[i].x]);
                            comp_set(fb[i], cpt, xb[ka[i].y]);
                        }
                        __syncwarp();
                    }
                    if (NF == 1) {
#pragma unroll
                        for (int i = 0; i < kQ; ++i) {
                            za[i] = fa[i];
                            zb[i] = fb[i];
                        }
                    } else if (sub == 0) {
                        // park F0 in the output slots (L2-resident) instead of 12 more live registers
#pragma unroll
                        for (int i = 0; i < kQ; ++i) {
                            const int y = lane + 32 * i;
                            if (y < kNC) {
                                double2* o = out + (size_t)y * kRows;
                                o[0] = fa[i];
                                o[1] = fb[i];
                            }
                        }
                    } else {
                        // X[k] = F0[k mod 1280] + w_N^k F1[k mod 1280]
                        const double2* ws = p.wsamp + (size_t)lam * 2 * kNC;
#pragma unroll
                        for (int i = 0; i < kQ; ++i) {
                            const int y = lane + 32 * i;
                            if (y < kNC) {
                                const double2* o = out + (size_t)y * kRows;
                                za[i] = cadd(o[0], cmul(fa[i], __ldg(ws + y)));
                                zb[i] = cadd(o[1], cmul(fb[i], __ldg(ws + kNC + y)));
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < kQ; ++i) store_rows(out, i, za[i], zb[i]);
            }
        }
        // move on to the next round's first unit BEFORE its barrier: the items that lie wholly
        // before it are released and can be refilled while the slower warps finish this round
        base += kHotWarps;
        rel0 = base;
        more = seek(rel0);
    }
}

// ---------------- pruned column pass
// line f = ((draw*nlam + lam)*20 + m): kept row-pass frequencies 2m, 2m+1 of that PSF (adjacent in Y,
// one contiguous 2*Rows*16-byte block), summed over the ndir planes of the draw, Hermitian-extended
// along the half-plane row index and transformed as the real and imaginary part of one complex
// line.  Only the 80 sampled outputs k_j AND their mirrors -k_j are kept: with P the (real,
// point-symmetric) PSF and kc the line's row-pass frequency,
//   S[img][x][j] = scale * (-1)^(kc + k_j) * P[k_j][kc]    for the sample x whose frequency is +kc,
//   S[img][x'][j] = scale * (-1)^(kc + k_j) * P[-k_j][kc]  for the sample x' whose frequency is -kc
// (P[k_j][-kc] = P[-k_j][kc]): 20 transforms per PSF give all 80 x 80 samples (xmap of set_lambda_tables).
// Every warp owns a private tile that a TMA bulk copy fills while the warp transforms the
// previous line (the tile is dead as soon as its values sit in registers), so the kernel
// streams Y at HBM speed instead of waiting on 80 dependent 16-byte loads per lane.
template <int NF>
struct ColCfg {
    // dim 1280: the tile is dead once its values sit in registers, so it doubles as the exchange
    // buffer of the transform and the next line is fetched after the gather: 8 warps fit instead
    // of 6 (these kernels are one long dependent instruction stream per warp - warps per
    // scheduler count for more than the load latency the warps now hide for each other).
    // dim 2560 keeps a separate exchange buffer and fetches ahead (4 warps either way).
    static constexpr bool SharedTile = NF == 1;
    static constexpr int Warps = NF == 1 ? 8 : 4;
    static constexpr int TileElems = 2 * Dim<NF>::Rows;                 // double2 per tile
    static constexpr uint32_t TileBytes = TileElems * sizeof(double2);
    static constexpr size_t XbufBytes = SharedTile ? 0 : G::XBUF * sizeof(double);
    static constexpr size_t Smem = 128 + (size_t)(G::TW1 + G::TW2) * sizeof(double2) +
                                   (size_t)Warps * (TileBytes + XbufBytes);
    static_assert(TileBytes % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
    static_assert(!SharedTile || TileBytes >= G::XBUF * sizeof(double), "the tile must hold the exchange buffer");
    static_assert(Smem <= 232448, "column kernel shared memory exceeds the 227 KB per-CTA limit");
};

struct ColParams {
    const double2* Y;      // [nplanes][nlam][kNC][Rows]
    double* S;             // [nimg][kNS][kNS]
    const uint16_t* kidx;  // [nlam][kNS] sampled frequencies
    const uint16_t* kcol;  // [nlam][kNC] kept row-pass frequencies
    const short2* xmap;    // [nlam][kNC] sample index of +kcol / -kcol (-1: none)
    const double2* wsamp;  // [nlam][2][kNS] (NF = 2)
    int nlines, nlam, ndir;
    double scale;
};

template <int NF>
__global__ void __launch_bounds__(ColCfg<NF>::Warps * 32, 1)
hot_cols_kernel(ColParams p, const double2* __restrict__ g_tw) {
    using D = Dim<NF>;
    using C = ColCfg<NF>;
    constexpr int kLines = kNC / 2;   // lines (pairs of kept frequencies) per PSF
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);              // one mbarrier per warp
    double2* tw1 = reinterpret_cast<double2*>(smem_raw + 128);
    double2* tw2 = tw1 + G::TW1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* tile = tw2 + G::TW2 + (size_t)warp * C::TileElems;
    double* xb = C::SharedTile ? reinterpret_cast<double*>(tile)
                               : reinterpret_cast<double*>(tw2 + G::TW2 + (size_t)C::Warps * C::TileElems) +
                                     (size_t)warp * G::XBUF;
    uint64_t* bar = bars + warp;

    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int gw = blockIdx.x * C::Warps + warp, nw = gridDim.x * C::Warps;
    // tile of (line f, direction d)
    auto src_of = [&](int f, int d) {
        const int m = f % kLines, img = f / kLines;
        const int lam = img % p.nlam, draw = img / p.nlam;
        return p.Y + (((size_t)(draw * p.ndir + d) * p.nlam + lam) * kNC + 2 * m) * D::Rows;
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
        const int m = f % kLines, img = f / kLines, lam = img % p.nlam;
        const uint16_t* kx = p.kidx + (size_t)lam * kNS;
        int kj[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) kj[i] = (lane + 32 * i < kNS) ? (int)__ldg(kx + lane + 32 * i) : 0;
        double2 z[3], zm[3];   // X[k_j], X[-k_j]
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
                // the tile is dead: prefetch the next one of this warp's sequence (unless the
                // transform is about to use the tile as its exchange buffer)
                if (d + 1 < p.ndir) fetch(f, d + 1);
                else if (sub + 1 < NF) fetch(f, 0);
                else if (!C::SharedTile && f + nw < p.nlines) fetch(f + nw, 0);
            }
            warp_fft<kR3>(v, xb, tw1, tw2, lane);
            double2 fz[3], fm[3];
#pragma unroll
            for (int cpt = 0; cpt < 2; ++cpt) {
                fft_dump<kR3>(v, xb, lane, cpt);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    comp_set(fz[i], cpt, xb[nat_addr(kj[i] % kNB)]);
                    comp_set(fm[i], cpt, xb[nat_addr((D::N - kj[i]) % kNB)]);
                }
                __syncwarp();
            }
            if (NF == 1 || sub == 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    z[i] = fz[i];
                    zm[i] = fm[i];
                }
            } else {
                const double2* ws = p.wsamp + (size_t)lam * 2 * kNS;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (lane + 32 * i < kNS) {
                        z[i] = cadd(z[i], cmul(fz[i], __ldg(ws + lane + 32 * i)));
                        zm[i] = cadd(zm[i], cmul(fm[i], __ldg(ws + kNS + lane + 32 * i)));
                    }
            }
        }
        if (C::SharedTile && f + nw < p.nlines) fetch(f + nw, 0);   // the gathers above are done (__syncwarp)
        // real part = kept frequency 2m, imaginary part = 2m + 1; each fills the sample row of +kc from
        // the outputs at k_j and the sample row of -kc from the outputs at -k_j
        const int kc1 = __ldg(p.kcol + (size_t)lam * kNC + 2 * m), kc2 = __ldg(p.kcol + (size_t)lam * kNC + 2 * m + 1);
        const short2 x1 = __ldg(p.xmap + (size_t)lam * kNC + 2 * m), x2 = __ldg(p.xmap + (size_t)lam * kNC + 2 * m + 1);
        double* o = p.S + (size_t)img * kNS * kNS;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int j = lane + 32 * i;
            if (j < kNS) {
                const double s1 = ((kc1 + kj[i]) & 1) ? -p.scale : p.scale;
                const double s2 = ((kc2 + kj[i]) & 1) ? -p.scale : p.scale;
                if (x1.x >= 0) o[x1.x * kNS + j] = s1 * z[i].x;
                if (x1.y >= 0) o[x1.y * kNS + j] = s1 * zm[i].x;
                if (x2.x >= 0) o[x2.x * kNS + j] = s2 * z[i].y;
                if (x2.y >= 0) o[x2.y * kNS + j] = s2 * zm[i].y;
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
    HotParams p{c->d_dphi, c->d_otf, c->d_ybuf, c->d_kaddr, c->d_wcol, c->d_dmin, c->d_dphi32,
                c->d_otf32, c->d_tw32, c->d_csort, c->d_lorder, c->d_counter, c->exp_cut, c->exp_grade, c->f32_rows, nplanes, nlam};
    int grid = c->sm_count;
    if (grid > nplanes * D::Pairs) grid = nplanes * D::Pairs;
    int rc;
    if (c->row_kernel == 2) {
        if ((rc = run_group_rows(c, nplanes, nlam, s))) return rc;
    } else {
        PSFR_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int), s));
        if ((rc = hot_event(c, 0, s))) return rc;
        hot_rows_kernel<NF><<<grid, C::Warps * 32, C::Smem, s>>>(p, c->d_tw);
        PSFR_LAUNCH_CHECK(c);
        if ((rc = hot_event(c, 1, s))) return rc;
        c->hot_launches += 1;
        c->hot_psfs += (long long)nplanes * nlam;
    }
    // psd_to_psf divides by the PSF sum (= T centre = 1/N^2, cancelling the 1/N^2 of the
    // inverse transform); psf_muse averages the directions.
    if ((rc = ensure_dynamic_smem(c, hot_cols_kernel<NF>, ColCfg<NF>::Smem))) return rc;
    ColParams q{c->d_ybuf, c->d_samp, c->d_kidx, c->d_kcol, c->d_xmap, c->d_wsamp, ndraw * nlam * (kNC / 2), nlam, ndir, 1.0 / ndir};
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
