// Row kernel of the pruned stage B (PSFR_OPT_ROW_KERNEL = 2, the default; 1 selects hot_rows_kernel of
// psfr_hot.cu).  Same work, same stream of units, same ring; what differs is who runs a transform:
//
//   hot_rows_kernel : one WARP = one 1280-point transform, 40 points per lane in registers,
//                     ~70 KB of unrolled code per unit, 8 warps per SM (253 registers);
//   group_rows_kernel: one GROUP of 128 threads = one transform that lives in shared memory
//                     (21.9 KB), two full radix passes (10, 16) and a PRUNED third one that evaluates
//                     only the 80 outputs the column pass needs (the 40 kept frequencies and their
//                     mirrors), a few KB of code, four groups = 16 warps per SM.
//
// Index maps (psfr_internal.h; kG1 x kG2 x 8 = 10 x 16 x 8, or 8 x 20 x 8 with PSFR_G_GEOM = 0):
//   n = n1*kGThreads + n2*8 + n3,  k = k1 + kG1*k2 + 160*k3.
//   pass 1 : thread b = n2*8 + n3 evaluates its own kG1 inputs x[n1*kGThreads + b] (both rows:
//            exp(-c D) * T, graded like hot_rows_kernel but per 32-cell segment), radix-kG1 over
//            n1, times w_160^(n2 k1)                    -> buf[k1*kGS1 + n3*kGS2 + n2]
//   pass 2 : threads (k1, n3) < 8 kG1: radix-kG2 over n2 (no twiddles), in place, storing only the
//            rows pass 3 reads                          -> buf[k1*kGS1 + n3*kGS2 + k2]
//   pass 3 : the last 80 threads, one per needed output k: X[k] = sum_n3 buf[k1*kGS1 + n3*kGS2 + k2] (w_N^k)^n3
//            by Horner's rule - the twiddle between passes 2 and 3 and the radix-8 phase are one
//            power series in w_N^k, held per (wavelength, thread) in a table (set_lambda_tables);
//            adjacent threads hold X[k], X[-k]: one shuffle, untangle the two real rows, one
//            32-byte store per kept frequency.
// The full transform's third pass (64 x radix-20, natural-order dump, gather of 160 values) took two
// more trips through shared memory and the kernel was shared-memory-bound (profiles/).
// The 10 x 16 x 8 split keeps 128 / 80 / 80 of a group's 128 threads busy in the three passes (8 x 20 x 8:
// 160 / 64 / 80 of 160) and fits four transforms per SM within 128 registers per thread.
// Strides kGS1 / kGS2 (odd n3 stride, kGS1 = 1 mod 8) make the 16-byte accesses of passes 1 and 2
// conflict-free (a quarter-warp hits eight distinct 16-byte slots); in pass 3 the slot of a thread is
// (k1 + k2 + kGS2 n3) mod 8 and the host orders the outputs so that the eight threads of a quarter-warp
// differ in (k1 + k2) mod 8 wherever the wavelength's frequency set allows it.
//
// Row pairs below exp(-f32_min) run the same passes in single precision, two wavelengths at a time
// as the halves of one packed transform (Z2, warp_fft.cuh); each half has its own pass-3 outputs.
//
// dim 2560 (NF = 2): a line is two interleaved 1280-point transforms (even / odd samples) that run one
// after the other through the same passes; pass 3 keeps F0[k mod 1280] in its registers and finishes
// X[k] = F0 + w_2560^k F1 after the second.  A ring stage then holds two rows of D and of the
// telescope OTF in FP64 only (80 KB), which leaves room for three groups; units are all FP64
// (underflow cut per 32-cell segment, no single-precision grades).
#include "pass_kernel.cuh"
#include "fast_exp.cuh"
#include "tma.cuh"

// tuning switches of the experiments recorded in DESIGN.md 3.11
#ifndef PSFR_G_P2MASK
#define PSFR_G_P2MASK 1    // pass 2 stores only the rows pass 3 reads
#endif

namespace psfr {

int hot_event(Ctx* c, int which, cudaStream_t s);

namespace {

constexpr int kGT = kGThreads;            // threads per group
constexpr int kS2 = kGS2, kS1 = kGS1;     // strides of n3 and k1 in a transform buffer
constexpr int kBuf = kG1 * kS1;           // double2 per transform buffer
constexpr int kP2Threads = 8 * kG1;       // rows (k1, n3) of pass 2
constexpr int kP3First = kGT - 2 * kNC;   // first pass-3 thread of a group
constexpr int kStages = 2, kTabMax = 64;
static_assert(kP3First >= 16 && kP3First % 32 == 16, "pass 3 starts in the upper half of a warp");
static_assert(kGT % 32 == 0, "groups are made of whole warps");

// launch shape and shared-memory budget per grid size
template <int NF>
struct GC {
    using D = Dim<NF>;
    static constexpr bool F32 = NF == 1;                 // single-precision copies staged, grades on
    static constexpr int Groups = NF == 1 ? kGGroups : 3;   // transforms in flight per CTA
    static constexpr int N = D::N, Rows = D::Rows, Pairs = D::Pairs, Tile = 2 * D::N;
    static constexpr uint32_t TileBytes = Tile * sizeof(double), TileBytes32 = Tile * sizeof(float);
    static constexpr uint32_t StageBytes = 2 * TileBytes + (F32 ? 2 * TileBytes32 : 0);
    static constexpr size_t StageDoubles = StageBytes / sizeof(double);
    static constexpr bool Tabs = NF == 1;                // per-wavelength scalars staged in shared memory
    static constexpr size_t Smem = 128 + (size_t)kGroupTw * sizeof(double2) + (size_t)kStages * StageBytes +
                                   (size_t)Groups * kBuf * sizeof(double2) +
                                   (Tabs ? kTabMax * (2 * sizeof(double) + 4 * sizeof(int)) : 0);
    static_assert(Smem <= 232448, "group kernel shared memory exceeds the 227 KB per-CTA limit");
    static_assert(Groups * kGT <= 1024, "too many threads");
};

// radix-R butterfly of a thread's R values, natural order in and out
template <int R, class Z>
__device__ __forceinline__ void dft_any(Z* x) {
    static_assert(R == 8 || R == 10 || R == 16 || R == 20, "unsupported pass radix");
    if constexpr (R == 8) dft8(x);
    else if constexpr (R == 16) dft16(x);
    else dft_r3<R>(x);
}

struct Rows2Params {
    const double* D;       // [nplanes][Rows][N]
    const double* T;       // [Rows][N]
    const float* D32;      // single-precision copies (dim 1280)
    const float* T32;
    double2* Y;            // [nplanes][nlam][kNC][Rows]
    const GroupP3* p3;     // [nlam][2 kNC] pass-3 records per wavelength
    const uint32_t* rows;  // [nlam][kGMaskStride] rows (k1, k2) pass 3 reads
    const double* dmin;    // [nplanes][Rows]
    const double* csort;   // [nlam] descending
    const int* lorder;     // [nlam]
    const float2* tw32;    // [kG2][kG1 - 1] single-precision pass-1 twiddles (global memory, L1-resident)
    int* next_item;
    double cut, grade, f32_min;
    int nplanes, nlam;
};

__device__ __forceinline__ void group_bar(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kGT) : "memory");
}

// pass-1 twiddles of a thread, w_160^(n2 k1), k1 = 1..kG1-1: the FP64 ones from the shared-memory table,
// the single-precision ones from the L1-resident float table
struct TwSmem {
    const double2* p1;  // + n2 * (kG1 - 1), shared memory
    __device__ __forceinline__ double2 tw1(int k1) const { return p1[k1 - 1]; }
};
struct TwMem32Pair {   // the float table, broadcast into both halves of a packed pair
    const float2* p1;
    __device__ __forceinline__ Z2 tw1(int k1) const { return ztw<Z2>(__ldg(p1 + (k1 - 1))); }
};

// Pass 3 of one thread: X[k] = sum_n3 v[n3] w^n3 by Horner's rule over the eight values of its row -
// FP64 for a double2 buffer, FP32 on one half (wavelength) of a packed-pair buffer.  The thread's record
// is fetched (fetch_p3) while pass 2 runs: its L2 latency must not sit between the barriers.
struct P3Reg {
    double2 w, wc;
    float2 w32;
    int base, col;
};
template <int NF>
__device__ __forceinline__ P3Reg fetch_p3(const GroupP3* __restrict__ tab) {
    const uint4* q = reinterpret_cast<const uint4*>(tab);
    const uint4 a = __ldg(q), c = __ldg(q + 2);
    P3Reg e;
    e.w = make_double2(__hiloint2double((int)a.y, (int)a.x), __hiloint2double((int)a.w, (int)a.z));
    if (NF == 2) {
        const uint4 m = __ldg(q + 1);
        e.wc = make_double2(__hiloint2double((int)m.y, (int)m.x), __hiloint2double((int)m.w, (int)m.z));
    }
    e.w32 = make_float2(__uint_as_float(c.x), __uint_as_float(c.y));
    e.base = (int)c.z;
    e.col = (int)c.w;
    return e;
}
__device__ __forceinline__ double2 pass3(const double2* buf, const P3Reg& e, int) {
    const double2* r = buf + e.base;
    double2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = r[i * kS2];
    double2 a = v[7];
#pragma unroll
    for (int i = 6; i >= 0; --i)
        a = make_double2(fma(a.x, e.w.x, fma(-a.y, e.w.y, v[i].x)), fma(a.x, e.w.y, fma(a.y, e.w.x, v[i].y)));
    return a;
}
__device__ __forceinline__ double2 pass3(const Z2* buf, const P3Reg& e, int h) {
    float2 v[8];
    const float4* r = reinterpret_cast<const float4*>(buf + e.base);   // (x.A, x.B, y.A, y.B) per element
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 q = r[i * kS2];
        v[i] = h ? make_float2(q.y, q.w) : make_float2(q.x, q.z);
    }
    float2 a = v[7];
#pragma unroll
    for (int i = 6; i >= 0; --i)
        a = make_float2(fmaf(a.x, e.w32.x, fmaf(-a.y, e.w32.y, v[i].x)), fmaf(a.x, e.w32.y, fmaf(a.y, e.w32.x, v[i].y)));
    return make_double2((double)a.x, (double)a.y);
}

// Passes 1 and 2 of one transform on the group's buffer, from the kG1 pass-1 inputs of every thread to
// the rows pass 3 reads (see the file header for the index maps).  rowsA / rowsB: needed-row masks of the
// transform (of the two halves of a packed pair).  Ends with the barrier that publishes pass 2.
template <class Z, class TW>
__device__ __forceinline__ void group_passes12(Z (&x)[kG1], Z* buf, const TW& tw, int b, int br, int grp,
                                               const uint32_t* __restrict__ rowsA, const uint32_t* __restrict__ rowsB) {
    dft_any<kG1>(x);
    group_bar(grp);   // pass 3 of the previous transform is done with the buffer (these inputs were evaluated meanwhile)
    {
        Z* dst = buf + (b & 7) * kS2 + (b >> 3);
        dst[0] = x[0];
#pragma unroll
        for (int k1 = 1; k1 < kG1; ++k1) dst[k1 * kS1] = cmul(x[k1], tw.tw1(k1));
    }
    group_bar(grp);
    // ---- pass 2: radix-kG2 over n2 by the first 8 kG1 threads of the group (rows n3-fastest: a quarter-
    // warp hits eight distinct 16-byte slots), in place - a row belongs to one thread, so no barrier
    // inside - and only the outputs k2 that pass 3 will read are stored (the mask is the same for the
    // eight lanes of a quarter-warp: whole wavefronts are saved).
    if (br < kP2Threads) {
        Z* row = buf + (br >> 3) * kS1 + (br & 7) * kS2;
        Z z[kG2];
#pragma unroll
        for (int i = 0; i < kG2; ++i) z[i] = row[i];
        dft_any<kG2>(z);
#if PSFR_G_P2MASK
        uint32_t need = __ldg(rowsA + (br >> 3));
        if (rowsB != nullptr) need |= __ldg(rowsB + (br >> 3));
#pragma unroll
        for (int i = 0; i < kG2; ++i)
            if ((need >> i) & 1) row[i] = z[i];
#else
#pragma unroll
        for (int i = 0; i < kG2; ++i) row[i] = z[i];
#endif
    }
}

// the pair (X[k], X[-k]) on adjacent lanes untangles the two packed real rows: one 32-byte store per
// kept frequency.  Whole warps enter so that the shuffle has a compile-time full mask.
__device__ __forceinline__ void untangle_store(double2 mine, bool active, int b, double2* dst) {
    double2 other;
    other.x = __shfl_xor_sync(0xffffffffu, mine.x, 1);
    other.y = __shfl_xor_sync(0xffffffffu, mine.y, 1);
    if (active && !(b & 1)) {
        const double2 za = mine, zb = other;
        st_global_256(dst, make_double2(0.5 * (za.x + zb.x), 0.5 * (za.y - zb.y)),
                      make_double2(0.5 * (za.y + zb.y), 0.5 * (zb.x - za.x)));
    }
}

template <int NF>
__global__ void __launch_bounds__(GC<NF>::Groups* kGT, 1)
group_rows_kernel(Rows2Params p, const double2* __restrict__ g_tw) {
    using C = GC<NF>;
    constexpr int kGroups = C::Groups, kN = C::N, kRows = C::Rows, kPairs = C::Pairs, kTile = C::Tile;
    constexpr size_t kStageDoubles = C::StageDoubles;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    int* released = reinterpret_cast<int*>(full + kStages);
    volatile int* item_of = released + kStages;
    volatile int* la_of = item_of + kStages;   // sorted positions [0, la) dead, [la, lb) single precision,
    volatile int* lb_of = la_of + kStages;     //   [lb, nlam) FP64 (as in hot_rows_kernel)
    volatile int* ns_of = lb_of + kStages;     // stream slots of the item: pairs of single-precision units, FP64 units
    double2* tw1 = reinterpret_cast<double2*>(smem_raw + 128);   // [kG2][kG1 - 1] pass-1 twiddles w_160^(n2 k1)
    double* ring = reinterpret_cast<double*>(tw1 + kGroupTw);
    double2* bufs = reinterpret_cast<double2*>(ring + (size_t)kStages * kStageDoubles);
    double* tab_c = reinterpret_cast<double*>(bufs + (size_t)kGroups * kBuf);   // C::Tabs only
    double* tab_rc = tab_c + kTabMax;
    int* tab_lo = reinterpret_cast<int*>(tab_rc + kTabMax);
    int* tab_cut = tab_lo + kTabMax;      // per sorted wavelength: float bit patterns of cut / c, grade / c
    int* tab_grade = tab_cut + kTabMax;   //   and of -c log2(e) (the single-precision exp is 2^(that * D))
    int* tab_n2f = tab_grade + kTabMax;
    const int grp = threadIdx.x / kGT, b = threadIdx.x % kGT;
    double2* buf = bufs + (size_t)grp * kBuf;
    const int items = p.nplanes * kPairs;
    const bool tabbed = C::Tabs && p.nlam <= kTabMax;
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
            while (C::F32 && lo < hi) {
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
            mbar_expect_tx(full + s, C::StageBytes);
            tma_load_1d(dst, p.D + off, C::TileBytes, full + s);
            tma_load_1d(dst + kTile, p.T + offT, C::TileBytes, full + s);
            if (C::F32) {
                float* dst32 = reinterpret_cast<float*>(dst + 2 * kTile);
                tma_load_1d(dst32, p.D32 + off, C::TileBytes32, full + s);
                tma_load_1d(dst32 + kTile, p.T32 + offT, C::TileBytes32, full + s);
            }
        } else {
            item_of[s] = -1;
            mbar_arrive(full + s);
        }
    };

    for (int i = threadIdx.x; i < kGroupTw; i += blockDim.x) tw1[i] = g_tw[i];
    const TwSmem twr{tw1 + (b >> 3) * (kG1 - 1)};
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

    // Role index of passes 2 and 3: the thread index rotated by one warp per group.  Pass 2 (the heavy one)
    // runs on role threads 0..79, pass 3 on 48..127; a scheduler holds the same warp of every group, so
    // without the rotation two of the four schedulers would carry every group's pass 2.
    const int br = (b + 32 * grp) % kGT;
    const bool p3 = br >= kP3First, p3warp = br >= kP3First - 16;
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
        const int plane = item / kPairs, rp = item % kPairs;
        auto lam_of = [&](int pos) { return tabbed ? tab_lo[pos] : __ldg(p.lorder + pos); };
        auto out_of = [&](int lam) { return p.Y + ((size_t)plane * p.nlam + lam) * kNC * kRows + 2 * rp; };
        if constexpr (C::F32) {
            const float* sD32 = reinterpret_cast<const float*>(sD + 2 * kTile);
            const float* sT32 = sD32 + kTile;
            auto n2f_of = [&](int pos) {
                return tabbed ? __int_as_float(tab_n2f[pos]) : (float)(-c_of(pos) * 1.44269504088896338700);
            };
            auto cut_of = [&](int pos) { return tabbed ? tab_cut[pos] : __float_as_int((float)(p.cut / c_of(pos))); };
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
                const TwMem32Pair twp{p.tw32 + (b >> 3) * (kG1 - 1)};
                // The inputs are fetched and classified in two batches of kG1 / 2 cells: all loads of a batch
                // are in flight together and ONE warp reduction (REDUX.AND on the packed per-cell bits) replaces
                // a vote + branch per cell, whose latencies used to add up cell after cell.
                Z2 x[kG1];
#pragma unroll
                for (int hb = 0; hb < kG1; hb += kG1 / 2) {
                    float d0[kG1 / 2], d1[kG1 / 2], t0[kG1 / 2], t1[kG1 / 2];
                    unsigned bits = 0;
#pragma unroll
                    for (int j = 0; j < kG1 / 2; ++j) {
                        const int n = (hb + j) * kGT + b;
                        d0[j] = sD32[n];
                        d1[j] = sD32[kN + n];
                        t0[j] = sT32[n];
                        t1[j] = sT32[kN + n];
                        const bool dead = ((t0[j] == 0.f) | (__float_as_int(d0[j]) >= cut32)) &
                                          ((t1[j] == 0.f) | (__float_as_int(d1[j]) >= cut32));
                        bits |= (unsigned)dead << j;
                    }
                    const unsigned all = __reduce_and_sync(0xffffffffu, bits);
#pragma unroll
                    for (int j = 0; j < kG1 / 2; ++j) {
                        if ((all >> j) & 1) {
                            x[hb + j].x = F2(0.f, 0.f);
                            x[hb + j].y = F2(0.f, 0.f);
                        } else {
                            x[hb + j].x = F2(ex2_approx(nA * d0[j]) * t0[j], ex2_approx(nB * d0[j]) * t0[j]);
                            x[hb + j].y = F2(ex2_approx(nA * d1[j]) * t1[j], ex2_approx(nB * d1[j]) * t1[j]);
                        }
                    }
                }
                Z2* zbuf = reinterpret_cast<Z2*>(buf);
                group_passes12(x, zbuf, twp, b, br, grp, p.rows + lamA * kGMaskStride, p.rows + lamB * kGMaskStride);
                // the pass-3 threads fetch their records while pass 2 runs
                P3Reg ea, eb;
                ea.col = eb.col = 0;
                if (p3) {
                    ea = fetch_p3<1>(p.p3 + (size_t)lamA * 2 * kNC + (br - kP3First));
                    eb = fetch_p3<1>(p.p3 + (size_t)lamB * 2 * kNC + (br - kP3First));
                }
                group_bar(grp);
                if (p3warp) {
                    untangle_store(p3 ? pass3(zbuf, ea, 0) : make_double2(0.0, 0.0), p3, b, out_of(lamA) + (size_t)ea.col * kRows);
                    if (two)
                        untangle_store(p3 ? pass3(zbuf, eb, 1) : make_double2(0.0, 0.0), p3, b,
                                       out_of(lamB) + (size_t)eb.col * kRows);
                }
                base += kGroups;
                continue;
            }
        }
        // ---- FP64 unit; dim 1280: the exp is graded per 32-cell segment of both rows; dim 2560: two
        // interleaved sub-transforms, pass 3 combines them
        const int pos = lb + (slot - npair);
        const int lam = lam_of(pos);
        const double cl = c_of(pos), rcl = tabbed ? tab_rc[pos] : 1.0 / cl;
        const double negc = -cl;
        const uint32_t* need = p.rows + lam * kGMaskStride;
        P3Reg e;
        e.col = 0;
        double2 f0 = make_double2(0.0, 0.0), mine = make_double2(0.0, 0.0);
#pragma unroll 1
        for (int sub = 0; sub < NF; ++sub) {
            double2 x[kG1];
            if constexpr (C::F32) {
                const float* sD32 = reinterpret_cast<const float*>(sD + 2 * kTile);
                const float* sT32 = sD32 + kTile;
                const float negc2f = tabbed ? __int_as_float(tab_n2f[pos]) : (float)(negc * 1.44269504088896338700);
                const int cut32 = tabbed ? tab_cut[pos] : __float_as_int((float)(p.cut * rcl));
                const int grade32 = tabbed ? tab_grade[pos] : __float_as_int((float)(p.grade * rcl));
#pragma unroll
                for (int hb = 0; hb < kG1; hb += kG1 / 2) {   // two batches, one warp reduction each (see the pair path)
                    float d0[kG1 / 2], d1[kG1 / 2], t0[kG1 / 2], t1[kG1 / 2];
                    unsigned bits = 0;
#pragma unroll
                    for (int j = 0; j < kG1 / 2; ++j) {
                        const int n = (hb + j) * kGT + b;
                        d0[j] = sD32[n];
                        d1[j] = sD32[kN + n];
                        t0[j] = sT32[n];
                        t1[j] = sT32[kN + n];
                        const bool z0 = t0[j] == 0.f, z1 = t1[j] == 0.f;
                        const int h0 = __float_as_int(d0[j]), h1 = __float_as_int(d1[j]);
                        const bool dead = (z0 | (h0 >= cut32)) & (z1 | (h1 >= cut32));
                        const bool cheap = (z0 | (h0 >= grade32)) & (z1 | (h1 >= grade32));
                        bits |= ((unsigned)dead << j) | ((unsigned)cheap << (8 + j));
                    }
                    const unsigned all = __reduce_and_sync(0xffffffffu, bits);
#pragma unroll
                    for (int j = 0; j < kG1 / 2; ++j) {
                        const int n = (hb + j) * kGT + b;
                        if ((all >> j) & 1) {
                            x[hb + j] = make_double2(0.0, 0.0);
                        } else if ((all >> (8 + j)) & 1) {
                            x[hb + j] = make_double2(f2d_bits(ex2_approx(negc2f * d0[j]) * t0[j]),
                                                     f2d_bits(ex2_approx(negc2f * d1[j]) * t1[j]));
                        } else {
                            x[hb + j] = make_double2(fast_exp(negc * sD[n]) * sT[n], fast_exp(negc * sD[kN + n]) * sT[kN + n]);
                        }
                    }
                }
            } else {
                // cut threshold on D itself, tested on the integer pipe (a non-negative double orders like
                // its high word; the signed compare keeps a D rounded slightly below zero alive)
                const int cut_hi = __double2hiint(p.cut * rcl);
#pragma unroll
                for (int hb = 0; hb < kG1; hb += kG1 / 2) {
                    double d0[kG1 / 2], d1[kG1 / 2], t0[kG1 / 2], t1[kG1 / 2];
                    unsigned bits = 0;
#pragma unroll
                    for (int j = 0; j < kG1 / 2; ++j) {
                        const int n = NF * ((hb + j) * kGT + b) + sub;
                        d0[j] = sD[n];
                        d1[j] = sD[kN + n];
                        t0[j] = sT[n];
                        t1[j] = sT[kN + n];
                        const bool dead = (is_zero_bits(t0[j]) | (__double2hiint(d0[j]) >= cut_hi)) &
                                          (is_zero_bits(t1[j]) | (__double2hiint(d1[j]) >= cut_hi));
                        bits |= (unsigned)dead << j;
                    }
                    const unsigned all = __reduce_and_sync(0xffffffffu, bits);
#pragma unroll
                    for (int j = 0; j < kG1 / 2; ++j) {
                        if ((all >> j) & 1) x[hb + j] = make_double2(0.0, 0.0);
                        else x[hb + j] = make_double2(fast_exp(negc * d0[j]) * t0[j], fast_exp(negc * d1[j]) * t1[j]);
                    }
                }
            }
            group_passes12(x, buf, twr, b, br, grp, need, nullptr);
            if (sub == 0 && p3) e = fetch_p3<NF>(p.p3 + (size_t)lam * 2 * kNC + (br - kP3First));
            group_bar(grp);
            if (p3) {
                const double2 a = pass3(buf, e, 0);
                if (NF == 1) mine = a;
                else if (sub == 0) f0 = a;
                else mine = cadd(f0, cmul(a, e.wc));   // X[k] = F0[k mod 1280] + w_N^k F1[k mod 1280]
            }
        }
        if (p3warp) untangle_store(mine, p3, b, out_of(lam) + (size_t)e.col * kRows);
        base += kGroups;
    }
}

}  // namespace

template <int NF>
static int group_rows_t(Ctx* c, int nplanes, int nlam, cudaStream_t s) {
    using C = GC<NF>;
    if (int rc = ensure_dynamic_smem(c, group_rows_kernel<NF>, C::Smem)) return rc;
    Rows2Params p{c->d_dphi, c->d_otf, c->d_dphi32, c->d_otf32, c->d_ybuf, c->d_p3, c->d_p2mask, c->d_dmin, c->d_csort,
                  c->d_lorder, c->d_twg32, c->d_counter, c->exp_cut, c->exp_grade, c->f32_rows, nplanes, nlam};
    int grid = c->sm_count;
    if (grid > nplanes * C::Pairs) grid = nplanes * C::Pairs;
    PSFR_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int), s));
    int rc = hot_event(c, 0, s);
    if (rc) return rc;
    group_rows_kernel<NF><<<grid, C::Groups * kGT, C::Smem, s>>>(p, c->d_twg);
    PSFR_LAUNCH_CHECK(c);
    if ((rc = hot_event(c, 1, s))) return rc;
    c->hot_launches += 1;
    c->hot_psfs += (long long)nplanes * nlam;
    return PSFR_OK;
}

int run_group_rows(Ctx* c, int nplanes, int nlam, cudaStream_t s) {
    return c->NF == 1 ? group_rows_t<1>(c, nplanes, nlam, s) : group_rows_t<2>(c, nplanes, nlam, s);
}

}  // namespace psfr
