// Row kernel of the pruned stage B at dim 1280 (PSFR_OPT_ROW_KERNEL = 2, the default; 1 selects
// hot_rows_kernel of psfr_hot.cu, which also serves dim 2560).  Same work, same stream of units,
// same ring; what differs is who runs a transform:
//
//   hot_rows_kernel : one WARP = one 1280-point transform, 40 points per lane in registers,
//                     ~70 KB of unrolled code per unit, 8 warps per SM (253 registers);
//   group_rows_kernel: one GROUP of 160 threads = one transform that lives in shared memory
//                     (22.5 KB), three radix passes (8, 8, 20) separated by group barriers,
//                     8 or 20 points per thread, a few KB of code, 15 warps per SM.
//
// Index maps (those of warp_fft.cuh): n = n1*160 + n2*20 + n3,  k = k1 + 8*k2 + 64*k3.
//   pass 1 : thread b = n2*20 + n3 evaluates its own eight inputs x[n1*160 + b] (both rows:
//            exp(-c D) * T, graded like hot_rows_kernel but per 32-cell segment), radix-8 over
//            n1, times w_N^(b k1)                       -> buf[k1*172 + n2*21 + n3]
//   pass 2 : thread (k1, n3) radix-8 over n2, times w_160^(n3 k2), in place
//   pass 3 : threads (k1, k2) < 64: radix-20 over n3 -> natural order buf[k + (k >> 3)]
//   gather : thread pairs (2y, 2y+1) read X[k_y], X[-k_y], untangle the two real rows, store.
// The strides 172 / 21 and the skews make every 16-byte access pattern above conflict-free
// (quarter-warps hit eight distinct 16-byte slots).
//
// Row pairs below exp(-f32_min) run the same passes in single precision, two wavelengths at a time
// as the halves of one packed transform (Z2, warp_fft.cuh).
#include "pass_kernel.cuh"
#include "fast_exp.cuh"
#include "tma.cuh"

namespace psfr {

int hot_event(Ctx* c, int which, cudaStream_t s);

namespace {

constexpr int kGroups = 3;                // transforms in flight per CTA
constexpr int kGT = 160;                  // threads per group
constexpr int kBuf = 1440;                // double2 per transform buffer (layouts above)
constexpr int kN = kNB, kRows = kNB / 2 + 2, kPairs = kRows / 2, kTile = 2 * kNB;
constexpr uint32_t kTileBytes = kTile * sizeof(double), kTileBytes32 = kTile * sizeof(float);
constexpr uint32_t kStageBytes = 2 * kTileBytes + 2 * kTileBytes32;
constexpr size_t kStageDoubles = kStageBytes / sizeof(double);
constexpr int kStages = 2, kTabMax = 64;
constexpr size_t kSmem2 = 128 + (size_t)(G::TW1 + G::TW2) * sizeof(double2) + (size_t)kStages * kStageBytes +
                          (size_t)kGroups * kBuf * sizeof(double2) + kTabMax * (2 * sizeof(double) + 4 * sizeof(int));
static_assert(kSmem2 <= 232448, "group kernel shared memory exceeds the 227 KB per-CTA limit");

struct Rows2Params {
    const double* D;       // [nplanes][kRows][N]
    const double* T;       // [kRows][N]
    const float* D32;
    const float* T32;
    double2* Y;            // [nplanes][nlam][kNC][kRows]
    const uint16_t* kcol;  // [nlam][kNC] row-pass frequencies kept per PSF
    const double* dmin;    // [nplanes][kRows]
    const double* csort;   // [nlam] descending
    const int* lorder;     // [nlam]
    const float2* tw32;    // single-precision twiddles (global memory, L1-resident)
    int* next_item;
    double cut, grade, f32_min;
    int nplanes, nlam;
};

__device__ __forceinline__ void group_bar(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kGT) : "memory");
}

// Layout buf[k1 * S + m * 21 + n3] of passes 1/2 and who does what in passes 2 and 3.  Both element
// types are 16 bytes (a double2, or the packed pair Z2): a quarter-warp must hit 8 distinct
// 16-byte slots, which S = 172, pass 2 with n3 fastest along the lanes and pass 3 with k2 fastest
// achieve.
template <class Z>
struct Lay {
    static_assert(sizeof(Z) == 16, "the layouts are chosen for 16-byte elements");
    static constexpr int S = 172;
    __device__ static int p2_k1(int b) { return b / 20; }
    __device__ static int p2_n3(int b) { return b % 20; }
    __device__ static int p3_k1(int b) { return b >> 3; }
    __device__ static int p3_k2(int b) { return b & 7; }
};

// twiddles of a thread's two radix-8 butterflies: w_N^(b k1) and w_160^(n3 k2), k = 1..7: the FP64
// ones from the shared-memory tables, the single-precision ones from the L1-resident float table
struct TwSmem {
    const double2* p1;  // + (j*7)*32 + t, shared memory
    const double2* p2;  // + n3
    __device__ __forceinline__ double2 tw1(int k1) const { return p1[(k1 - 1) * 32]; }
    __device__ __forceinline__ double2 tw2(int k2) const { return p2[(k2 - 1) * kR3]; }
};
struct TwMem32Pair {   // the same float table, broadcast into both halves of a packed pair
    const float2* p1;
    const float2* p2;
    __device__ __forceinline__ Z2 tw1(int k1) const { return ztw<Z2>(__ldg(p1 + (k1 - 1) * 32)); }
    __device__ __forceinline__ Z2 tw2(int k2) const { return ztw<Z2>(__ldg(p2 + (k2 - 1) * kR3)); }
};

// one half (wavelength) of a packed pair / the value itself, as FP64
__device__ __forceinline__ double2 half_of(const Z2& m, int h) {
    return h ? make_double2((double)m.x.v.y, (double)m.y.v.y) : make_double2((double)m.x.v.x, (double)m.y.v.x);
}
__device__ __forceinline__ double2 half_of(const double2& m, int) { return m; }

// passes of one transform on the group's buffer, from the eight pass-1 inputs of every thread to the
// store of the 80 sampled frequencies of both rows (see the file header for the index maps).
template <class Z, class TW>
__device__ __forceinline__ void group_transform(Z (&x)[8], Z* buf, const TW& tw, int b, int grp,
                                                const uint16_t* __restrict__ kidx, double2* __restrict__ out,
                                                const uint16_t* __restrict__ kidx2 = nullptr,
                                                double2* __restrict__ out2 = nullptr) {
    using L = Lay<Z>;
    constexpr int S = L::S;
    const int n2 = b / 20, n3 = b % 20;
    dft8(x);
    group_bar(grp);   // the gather of the previous unit is done with the buffer (its inputs were evaluated meanwhile)
    buf[n2 * 21 + n3] = x[0];
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) buf[k1 * S + n2 * 21 + n3] = cmul(x[k1], tw.tw1(k1));
    group_bar(grp);
    // ---- pass 2: radix-8 over n2 (thread = (k1, n3), see Lay), twiddle, in place
    {
        Z* col = buf + L::p2_k1(b) * S + L::p2_n3(b);
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = col[m * 21];
        dft8(x);
        col[0] = x[0];
#pragma unroll
        for (int k2 = 1; k2 < 8; ++k2) col[k2 * 21] = cmul(x[k2], tw.tw2(k2));
    }
    group_bar(grp);
    // ---- pass 3: radix-20 over n3 by the first two warps of the group (rows k2-fastest: a quarter-
    // warp hits eight distinct 16-byte slots), natural-order output.  Splitting it into 4 x 5
    // sub-passes over all 160 threads was measured SLOWER (3.63 vs 3.38 ms per chunk: one more
    // barrier and one more trip through shared memory cost more than the idle warps), and so was
    // giving a row to two threads (even / odd outputs, each a 10-point transform of the folded
    // row: 3.36 vs 3.21 ms - every row is then read twice, and the kernel is shared-memory-bound).
    if (b < 64) {
        const int k2 = L::p3_k2(b), k1 = L::p3_k1(b);
        Z z[20];
        const Z* row = buf + k1 * S + k2 * 21;
#pragma unroll
        for (int i = 0; i < 20; ++i) z[i] = row[i];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + kGroups + grp) : "memory");   // all rows read before any is overwritten
        dft_r3<kR3>(z);
        const int k0 = k1 + 8 * k2;
#pragma unroll
        for (int k3 = 0; k3 < 20; ++k3) {
            const int k = k0 + 64 * k3;
            buf[k + (k >> 3)] = z[k3];
        }
    }
    group_bar(grp);
    // ---- gather + untangle + store: thread pair (2y, 2y+1) holds X[k_y] and X[-k_y] of kept frequency y
    // (the first three warps; warp 2 runs with its upper half clamped so that the shuffle is warp-wide);
    // a packed pair does it once per wavelength (each has its own frequencies), reading its half
    if (b < 96) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint16_t* kx = h ? kidx2 : kidx;
            double2* o = h ? out2 : out;
            if (h && o == nullptr) break;   // a single transform, or a pair with only one wavelength
            const int y = min(b >> 1, kNC - 1);
            const int k = (int)__ldg(kx + y);
            const int kk = (b & 1) ? (kN - k) % kN : k;
            const double2 mine = half_of(buf[kk + (kk >> 3)], h);
            double2 other;
            other.x = __shfl_xor_sync(0xffffffffu, mine.x, 1);
            other.y = __shfl_xor_sync(0xffffffffu, mine.y, 1);
            if (!(b & 1) && b < 2 * kNC) {
                const double2 za = mine, zb = other;
                st_global_256(o + (size_t)y * kRows, make_double2(0.5 * (za.x + zb.x), 0.5 * (za.y - zb.y)),
                              make_double2(0.5 * (za.y + zb.y), 0.5 * (zb.x - za.x)));
            }
        }
    }
}

__global__ void __launch_bounds__(kGroups* kGT, 1)
group_rows_kernel(Rows2Params p, const double2* __restrict__ g_tw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    int* released = reinterpret_cast<int*>(full + kStages);
    volatile int* item_of = released + kStages;
    volatile int* la_of = item_of + kStages;   // sorted positions [0, la) dead, [la, lb) single precision,
    volatile int* lb_of = la_of + kStages;     //   [lb, nlam) FP64 (as in hot_rows_kernel)
    volatile int* ns_of = lb_of + kStages;     // stream slots of the item: pairs of single-precision units, FP64 units
    double2* tw1 = reinterpret_cast<double2*>(smem_raw + 128);
    double2* tw2 = tw1 + G::TW1;
    double* ring = reinterpret_cast<double*>(tw2 + G::TW2);
    double2* bufs = reinterpret_cast<double2*>(ring + (size_t)kStages * kStageDoubles);
    double* tab_c = reinterpret_cast<double*>(bufs + (size_t)kGroups * kBuf);
    double* tab_rc = tab_c + kTabMax;
    int* tab_lo = reinterpret_cast<int*>(tab_rc + kTabMax);
    int* tab_cut = tab_lo + kTabMax;      // per sorted wavelength: float bit patterns of cut / c, grade / c
    int* tab_grade = tab_cut + kTabMax;   //   and of -c log2(e) (the single-precision exp is 2^(that * D))
    int* tab_n2f = tab_grade + kTabMax;
    const int grp = threadIdx.x / kGT, b = threadIdx.x % kGT;
    double2* buf = bufs + (size_t)grp * kBuf;
    const int items = p.nplanes * kPairs;
    const bool tabbed = p.nlam <= kTabMax;
    auto c_of = [&](int pos) { return tabbed ? tab_c[pos] : __ldg(p.csort + pos); };

    auto issue = [&](int s) {
        const int item = atomicAdd(p.next_item, 1);
        if (item < items) {
            const int plane = item / kPairs, rp = item % kPairs;
            double* dst = ring + (size_t)s * kStageDoubles;
            item_of[s] = item;
            const double dm = fmin(__ldg(p.dmin + (size_t)plane * kRows + 2 * rp),
                                   __ldg(p.dmin + (size_t)plane * kRows + 2 * rp + 1));
            int lo = 0, hi = p.nlam;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (c_of(mid) * dm > p.cut) lo = mid + 1; else hi = mid;
            }
            la_of[s] = lo;
            const int la = lo;
            hi = p.nlam;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (c_of(mid) * dm >= p.f32_min) lo = mid + 1; else hi = mid;
            }
            lb_of[s] = lo;
            ns_of[s] = (lo - la + 1) / 2 + (p.nlam - lo);
            if (la == p.nlam) {
                mbar_arrive(full + s);
                return;
            }
            const size_t off = ((size_t)plane * kRows + 2 * rp) * kN, offT = (size_t)(2 * rp) * kN;
            mbar_expect_tx(full + s, kStageBytes);
            tma_load_1d(dst, p.D + off, kTileBytes, full + s);
            tma_load_1d(dst + kTile, p.T + offT, kTileBytes, full + s);
            float* dst32 = reinterpret_cast<float*>(dst + 2 * kTile);
            tma_load_1d(dst32, p.D32 + off, kTileBytes32, full + s);
            tma_load_1d(dst32 + kTile, p.T32 + offT, kTileBytes32, full + s);
        } else {
            item_of[s] = -1;
            mbar_arrive(full + s);
        }
    };

    for (int i = threadIdx.x; i < G::TW1 + G::TW2; i += blockDim.x) tw1[i] = g_tw[i];
    const TwSmem twr{tw1 + ((b >> 5) * 7) * 32 + (b & 31), tw2 + Lay<double2>::p2_n3(b)};
    const TwMem32Pair twp{p.tw32 + ((b >> 5) * 7) * 32 + (b & 31), p.tw32 + G::TW1 + Lay<Z2>::p2_n3(b)};
    if (tabbed)
        for (int i = threadIdx.x; i < p.nlam; i += blockDim.x) {
            const double cv = __ldg(p.csort + i);
            tab_c[i] = cv;
            tab_rc[i] = 1.0 / cv;
            tab_lo[i] = __ldg(p.lorder + i);
            tab_cut[i] = __float_as_int((float)(p.cut * (1.0 / cv)));
            tab_grade[i] = __float_as_int((float)(p.grade * (1.0 / cv)));
            tab_n2f[i] = __float_as_int((float)(-cv * 1.44269504088896338700));
        }
    __syncthreads();
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

    // The stream of live units and the walk through the items are those of hot_rows_kernel with
    // "warp" replaced by "group": slot base + grp of every round of kGroups slots, every group
    // passes through every item (wait for its fill, zero its share of the dead units, release
    // it), no barrier between the groups.
    int cur = 0, base = 0;
    bool seen = false;
    auto seek = [&](int& rel) -> bool {
        while (true) {
            const int s = cur % kStages;
            if (!seen) {
                mbar_wait(full + s, (cur / kStages) & 1);
                seen = true;
                const int item = item_of[s];
                if (item >= 0 && b < kNC) {
                    const int plane = item / kPairs, rp = item % kPairs, la = la_of[s];
                    for (int i = grp; i < la; i += kGroups) {
                        const int lam = tabbed ? tab_lo[i] : __ldg(p.lorder + i);
                        st_global_256(p.Y + (((size_t)plane * p.nlam + lam) * kNC + b) * kRows + 2 * rp,
                                      make_double2(0.0, 0.0), make_double2(0.0, 0.0));
                    }
                }
            }
            if (item_of[s] < 0) return false;
            const int n = ns_of[s];
            if (rel < n) return true;
            rel -= n;
            base -= n;
            group_bar(grp);   // every thread of the group is done with the stage
            if (b == 0) {
                __threadfence_block();
                const int old = atomicAdd(released + s, 1);
                if (old == kGroups - 1) {
                    atomicExch(released + s, 0);
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(s);
                }
            }
            ++cur;
            seen = false;
        }
    };

#pragma unroll 1
    for (;;) {
        int rel = base + grp;
        if (!seek(rel)) break;
        const int s = cur % kStages;
        const int item = item_of[s];
        // slot -> units (sorted positions): the first npair slots of an item hold two single-precision
        // units each (the last one may hold one), the others one FP64 unit; odd items run backwards
        const int la = la_of[s], lb = lb_of[s], npair = (lb - la + 1) / 2;
        const int slot = (cur & 1) ? ns_of[s] - 1 - rel : rel;
        const double* sD = ring + (size_t)s * kStageDoubles;
        const double* sT = sD + kTile;
        const float* sD32 = reinterpret_cast<const float*>(sD + 2 * kTile);
        const float* sT32 = sD32 + kTile;
        const int plane = item / kPairs, rp = item % kPairs;
        auto lam_of = [&](int pos) { return tabbed ? tab_lo[pos] : __ldg(p.lorder + pos); };
        auto out_of = [&](int lam) { return p.Y + ((size_t)plane * p.nlam + lam) * kNC * kRows + 2 * rp; };
        auto n2f_of = [&](int pos) {
            return tabbed ? __int_as_float(tab_n2f[pos]) : (float)(-c_of(pos) * 1.44269504088896338700);
        };
        auto cut_of = [&](int pos) {
            return tabbed ? tab_cut[pos] : __float_as_int((float)(p.cut / c_of(pos)));
        };
        if (slot < npair) {
            // ---- single-precision pair (every entry below exp(-f32_min) of the OTF peak at both
            // wavelengths): inputs, transform and buffer in FP32, the two wavelengths as the two halves
            // of one packed transform (Z2: FADD2 / FMUL2 / FFMA2)
            const int posA = la + 2 * slot;
            const bool two = posA + 1 < lb;
            const int posB = two ? posA + 1 : posA;
            const int lamA = lam_of(posA), lamB = lam_of(posB);
            const float nA = n2f_of(posA), nB = n2f_of(posB);
            const int cut32 = cut_of(posB);   // c_B <= c_A: an entry below the cut at B is below it at A
            Z2 x[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const int n = n1 * 160 + b;
                const float d0 = sD32[n], d1 = sD32[kN + n], t0 = sT32[n], t1 = sT32[kN + n];
                const bool dead = ((t0 == 0.f) | (__float_as_int(d0) >= cut32)) & ((t1 == 0.f) | (__float_as_int(d1) >= cut32));
                if (__all_sync(0xffffffffu, dead)) {
                    x[n1].x = F2(0.f, 0.f);
                    x[n1].y = F2(0.f, 0.f);
                } else {
                    x[n1].x = F2(ex2_approx(nA * d0) * t0, ex2_approx(nB * d0) * t0);
                    x[n1].y = F2(ex2_approx(nA * d1) * t1, ex2_approx(nB * d1) * t1);
                }
            }
            group_transform(x, reinterpret_cast<Z2*>(buf), twp, b, grp, p.kcol + (size_t)lamA * kNC, out_of(lamA),
                            p.kcol + (size_t)lamB * kNC, two ? out_of(lamB) : nullptr);
        } else {
            // ---- FP64 unit; the exp is graded per 32-cell segment of both rows
            const int pos = lb + (slot - npair);
            const int lam = lam_of(pos);
            const double cl = c_of(pos), rcl = tabbed ? tab_rc[pos] : 1.0 / cl;
            const double negc = -cl;
            const float negc2f = n2f_of(pos);
            const int cut32 = cut_of(pos);
            const int grade32 = tabbed ? tab_grade[pos] : __float_as_int((float)(p.grade * rcl));
            double2 x[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const int n = n1 * 160 + b;
                const float d0 = sD32[n], d1 = sD32[kN + n], t0 = sT32[n], t1 = sT32[kN + n];
                const bool z0 = t0 == 0.f, z1 = t1 == 0.f;
                const int h0 = __float_as_int(d0), h1 = __float_as_int(d1);
                const bool dead = (z0 | (h0 >= cut32)) & (z1 | (h1 >= cut32));
                const bool cheap = (z0 | (h0 >= grade32)) & (z1 | (h1 >= grade32));
                if (__all_sync(0xffffffffu, dead)) {
                    x[n1] = make_double2(0.0, 0.0);
                } else if (__all_sync(0xffffffffu, cheap)) {
                    x[n1] = make_double2(f2d_bits(ex2_approx(negc2f * d0) * t0), f2d_bits(ex2_approx(negc2f * d1) * t1));
                } else {
                    x[n1] = make_double2(fast_exp(negc * sD[n]) * sT[n], fast_exp(negc * sD[kN + n]) * sT[kN + n]);
                }
            }
            group_transform(x, buf, twr, b, grp, p.kcol + (size_t)lam * kNC, out_of(lam));
        }
        base += kGroups;
    }
}

}  // namespace

int run_group_rows(Ctx* c, int nplanes, int nlam, cudaStream_t s) {
    if (c->NF != 1) return set_error(c, PSFR_E_UNSUPPORTED, "the group row kernel is dim-1280 only");
    if (int rc = ensure_dynamic_smem(c, group_rows_kernel, kSmem2)) return rc;
    Rows2Params p{c->d_dphi, c->d_otf, c->d_dphi32, c->d_otf32, c->d_ybuf, c->d_kcol, c->d_dmin, c->d_csort,
                  c->d_lorder, c->d_tw32, c->d_counter, c->exp_cut, c->exp_grade, c->f32_rows, nplanes, nlam};
    int grid = c->sm_count;
    if (grid > nplanes * kPairs) grid = nplanes * kPairs;
    PSFR_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int), s));
    int rc = hot_event(c, 0, s);
    if (rc) return rc;
    group_rows_kernel<<<grid, kGroups * kGT, kSmem2, s>>>(p, c->d_tw);
    PSFR_LAUNCH_CHECK(c);
    if ((rc = hot_event(c, 1, s))) return rc;
    c->hot_launches += 1;
    c->hot_psfs += (long long)nplanes * nlam;
    return PSFR_OK;
}

}  // namespace psfr
