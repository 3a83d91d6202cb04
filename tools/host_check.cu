// CPU emulation of the warp FFT in csrc/warp_fft.cuh: runs the same __host__ __device__
// phase functions lane by lane and compares with a direct O(N^2) DFT.  Build + run:
//   nvcc -O2 -I muse_psfr_b200/csrc tools/host_check.cu -o /tmp/host_check && /tmp/host_check
// (add -DPSFR_G_GEOM=0 for the 8 x 20 x 8 geometry of the group transform)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "psfr_internal.h"
#include "warp_fft.cuh"
#include "fft_tables.h"
using namespace psfr;

// Z = double2: the product transform (two exchange rounds of one component each);
// Z = float2: the single-precision instantiation (one round of 8-byte float2 words).
template <int R3, class Z>
int check(double tol) {
    using G = FftGeom<R3>;
    using S = typename ZTraits<Z>::S;
    using W = typename ZTraits<Z>::W;
    constexpr int RND = ZTraits<Z>::Rounds;
    const int N = G::N;
    std::vector<double2> tw1, tw2;
    build_twiddles<R3>(tw1, tw2);
    std::vector<Z> x(N);
    srand(1);
    for (auto& z : x) z = mkz<Z>((S)(rand() / (double)RAND_MAX - 0.5), (S)(rand() / (double)RAND_MAX - 0.5));
    std::vector<std::vector<Z>> v(32, std::vector<Z>(40));
    std::vector<W> sm(G::XBUF);
    for (int t = 0; t < 32; ++t)
        for (int j = 0; j < 5; ++j)
            for (int n1 = 0; n1 < 8; ++n1) v[t][j * 8 + n1] = x[n1 * (N / 8) + t + 32 * j];
    for (int t = 0; t < 32; ++t) fft_pass1<R3>(v[t].data(), tw1.data(), t);
    for (int c = 0; c < RND; ++c) {
        for (int t = 0; t < 32; ++t) fft_x1_store<R3>(v[t].data(), sm.data(), t, c);
        for (int t = 0; t < 32; ++t) fft_x1_load<R3>(v[t].data(), sm.data(), t, c);
    }
    for (int t = 0; t < 32; ++t) fft_pass2<R3>(v[t].data(), tw2.data(), t);
    for (int c = 0; c < RND; ++c) {
        for (int t = 0; t < 32; ++t) fft_x2_store<R3>(v[t].data(), sm.data(), t, c);
        for (int t = 0; t < 32; ++t) fft_x2_load<R3>(v[t].data(), sm.data(), t, c);
    }
    for (int t = 0; t < 32; ++t) fft_pass3<R3>(v[t].data());
    std::vector<Z> X(N);
    for (int c = 0; c < RND; ++c) {
        for (int t = 0; t < 32; ++t) fft_dump<R3>(v[t].data(), sm.data(), t, c);
        for (int k = 0; k < N; ++k) word_set(X[k], c, sm[nat_addr(k)]);
    }
    double maxerr = 0, maxref = 0;
    for (int k = 0; k < N; k += 7) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            double2 w = unit_root((long long)n * k, N);
            sr += (long double)x[n].x * w.x - (long double)x[n].y * w.y;
            si += (long double)x[n].x * w.y + (long double)x[n].y * w.x;
        }
        double e = fabs((double)(sr - X[k].x)) + fabs((double)(si - X[k].y));
        if (e > maxerr) maxerr = e;
        double r = fabs((double)sr) + fabs((double)si);
        if (r > maxref) maxref = r;
    }
    printf("N=%d %s max err %.3e (ref scale %.3e) rel %.3e\n", N, sizeof(S) == 8 ? "double" : "float", maxerr,
           maxref, maxerr / maxref);
    return maxerr / maxref < tol ? 0 : 1;
}

// Z2: two independent single-precision transforms packed into one (f32x2 lanes on the device,
// emulated with scalar floats on the host); both halves are checked against the direct DFT.
template <int R3>
int check_packed(double tol) {
    using G = FftGeom<R3>;
    const int N = G::N;
    std::vector<double2> tw1d, tw2d;
    build_twiddles<R3>(tw1d, tw2d);
    std::vector<float2> tw1(tw1d.size()), tw2(tw2d.size());
    for (size_t i = 0; i < tw1d.size(); ++i) tw1[i] = make_float2((float)tw1d[i].x, (float)tw1d[i].y);
    for (size_t i = 0; i < tw2d.size(); ++i) tw2[i] = make_float2((float)tw2d[i].x, (float)tw2d[i].y);
    std::vector<Z2> x(N);
    srand(2);
    auto rnd = [] { return (float)(rand() / (double)RAND_MAX - 0.5); };
    for (auto& z : x) {
        z.x = F2(rnd(), rnd());
        z.y = F2(rnd(), rnd());
    }
    std::vector<std::vector<Z2>> v(32, std::vector<Z2>(40));
    std::vector<float2> sm(G::XBUF);
    for (int t = 0; t < 32; ++t)
        for (int j = 0; j < 5; ++j)
            for (int n1 = 0; n1 < 8; ++n1) v[t][j * 8 + n1] = x[n1 * (N / 8) + t + 32 * j];
    for (int t = 0; t < 32; ++t) fft_pass1<R3>(v[t].data(), tw1.data(), t);
    for (int c = 0; c < 2; ++c) {
        for (int t = 0; t < 32; ++t) fft_x1_store<R3>(v[t].data(), sm.data(), t, c);
        for (int t = 0; t < 32; ++t) fft_x1_load<R3>(v[t].data(), sm.data(), t, c);
    }
    for (int t = 0; t < 32; ++t) fft_pass2<R3>(v[t].data(), tw2.data(), t);
    for (int c = 0; c < 2; ++c) {
        for (int t = 0; t < 32; ++t) fft_x2_store<R3>(v[t].data(), sm.data(), t, c);
        for (int t = 0; t < 32; ++t) fft_x2_load<R3>(v[t].data(), sm.data(), t, c);
    }
    for (int t = 0; t < 32; ++t) fft_pass3<R3>(v[t].data());
    std::vector<Z2> X(N);
    for (int c = 0; c < 2; ++c) {
        for (int t = 0; t < 32; ++t) fft_dump<R3>(v[t].data(), sm.data(), t, c);
        for (int k = 0; k < N; ++k) word_set(X[k], c, sm[nat_addr(k)]);
    }
    double maxerr = 0, maxref = 0;
    for (int half = 0; half < 2; ++half)
        for (int k = 0; k < N; k += 7) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; ++n) {
                double2 w = unit_root((long long)n * k, N);
                const double xr = half ? x[n].x.v.y : x[n].x.v.x, xi = half ? x[n].y.v.y : x[n].y.v.x;
                sr += (long double)xr * w.x - (long double)xi * w.y;
                si += (long double)xr * w.y + (long double)xi * w.x;
            }
            const double gr = half ? X[k].x.v.y : X[k].x.v.x, gi = half ? X[k].y.v.y : X[k].y.v.x;
            double e = fabs((double)(sr - gr)) + fabs((double)(si - gi));
            if (e > maxerr) maxerr = e;
            double r = fabs((double)sr) + fabs((double)si);
            if (r > maxref) maxref = r;
        }
    printf("N=%d packed float pair max err %.3e (ref scale %.3e) rel %.3e\n", N, maxerr, maxref, maxerr / maxref);
    return maxerr / maxref < tol ? 0 : 1;
}

// The group transform of csrc/psfr_hot2.cu (one transform held in a complex buffer of kG1 * kGS1 entries:
// radix-kG1 and radix-kG2 passes, then the pruned third pass - Horner in w_N^k over the eight values of
// row (k mod kG1, (k div kG1) mod kG2)): same index maps and buffer layout, threads run one after the
// other with the group barriers as phase boundaries; every output k is checked.
template <int R>
void dft_host(double2* x) {
    if (R == 8) dft8(x);
    else if (R == 16) dft16(x);
    else dft_r3<(R == 8 || R == 16) ? 20 : R>(x);
}
int check_group() {
    constexpr int S2 = kGS2, S1 = kGS1, GT = kGThreads;
    const int N = 1280;
    std::vector<double2> x(N), buf(kG1 * S1);
    srand(3);
    for (auto& z : x) z = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    for (int b = 0; b < GT; ++b) {                        // pass 1: thread b = n2*8 + n3
        const int n2 = b >> 3, n3 = b & 7;
        double2 v[kG1];
        for (int n1 = 0; n1 < kG1; ++n1) v[n1] = x[n1 * GT + b];
        dft_host<kG1>(v);
        buf[n3 * S2 + n2] = v[0];
        for (int k1 = 1; k1 < kG1; ++k1) buf[k1 * S1 + n3 * S2 + n2] = cmul(v[k1], unit_root((long long)n2 * k1, 160));
    }
    for (int b = 0; b < 8 * kG1; ++b) {                   // pass 2: thread (k1, n3), radix-kG2 over n2, in place
        double2* row = buf.data() + (b >> 3) * S1 + (b & 7) * S2;
        double2 z[kG2];
        for (int i = 0; i < kG2; ++i) z[i] = row[i];
        dft_host<kG2>(z);
        for (int i = 0; i < kG2; ++i) row[i] = z[i];
    }
    double maxerr = 0, maxref = 0;
    for (int k = 0; k < N; ++k) {                         // pass 3 for every output
        const double2 w = unit_root(k, N);
        const double2* r = buf.data() + (k % kG1) * S1 + (k / kG1) % kG2;
        double2 acc = r[7 * S2];
        for (int n3 = 6; n3 >= 0; --n3) acc = cadd(cmul(acc, w), r[n3 * S2]);
        if (k % 3) continue;
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            double2 wn = unit_root((long long)n * k, N);
            sr += (long double)x[n].x * wn.x - (long double)x[n].y * wn.y;
            si += (long double)x[n].x * wn.y + (long double)x[n].y * wn.x;
        }
        maxerr = fmax(maxerr, fabs((double)(sr - acc.x)) + fabs((double)(si - acc.y)));
        maxref = fmax(maxref, fabs((double)sr) + fabs((double)si));
    }
    printf("N=%d group transform %d x %d x 8 (pruned pass 3) max err %.3e (ref scale %.3e) rel %.3e\n", N, kG1, kG2,
           maxerr, maxref, maxerr / maxref);
    return maxerr / maxref < 1e-14 ? 0 : 1;
}

int main() {
    // small DFT sanity: dft_r3 for 5,10,20,40
    int bad = 0;
    {
        double2 x[40], y[40];
        for (int R : {5, 10, 20, 40}) {
            for (int i = 0; i < R; ++i) x[i] = y[i] = make_double2(0.3 * i - 1, 0.1 * i * i - 2);
            if (R == 5) dft_r3<5>(y);
            if (R == 10) dft_r3<10>(y);
            if (R == 20) dft_r3<20>(y);
            if (R == 40) dft_r3<40>(y);
            double me = 0;
            for (int k = 0; k < R; ++k) {
                double sr = 0, si = 0;
                for (int n = 0; n < R; ++n) {
                    double2 w = unit_root(n * k, R);
                    sr += x[n].x * w.x - x[n].y * w.y;
                    si += x[n].x * w.y + x[n].y * w.x;
                }
                me = fmax(me, fabs(sr - y[k].x) + fabs(si - y[k].y));
            }
            printf("dft_r3<%d> err %.3e\n", R, me);
            if (me > 1e-11) bad = 1;
        }
    }
    {
        double2 x[16], y[16];
        for (int i = 0; i < 16; ++i) x[i] = y[i] = make_double2(0.37 * i - 1.3, 0.11 * i * i - 2.1);
        dft16(y);
        double me = 0;
        for (int k = 0; k < 16; ++k) {
            double sr = 0, si = 0;
            for (int n = 0; n < 16; ++n) {
                double2 w = unit_root(n * k, 16);
                sr += x[n].x * w.x - x[n].y * w.y;
                si += x[n].x * w.y + x[n].y * w.x;
            }
            me = fmax(me, fabs(sr - y[k].x) + fabs(si - y[k].y));
        }
        printf("dft16 err %.3e\n", me);
        if (me > 1e-12) bad = 1;
    }
    bad |= check<20, double2>(1e-14);
    bad |= check<20, float2>(2e-6);
    bad |= check_packed<20>(2e-6);
    bad |= check_group();
    return bad;
}
