// Host-side builders for the twiddle tables consumed by warp_fft.cuh.
#pragma once
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

namespace psfr {

// exp(+2 pi i e / M) with exact octant reduction (|angle| <= pi/4 before calling libm)
inline double2 unit_root(long long e, long long M) {
    e %= M;
    if (e < 0) e += M;
    // angle = 2 pi e / M = (pi/4) * (8e/M); split into octant o and remainder r in [-1/2, 1/2] octant units
    long long num = 8 * e;
    long long o = (num + M / 2) / M;      // nearest octant
    long long rem = num - o * M;          // in [-M/2, M/2]
    const double pi4 = 0.78539816339744830962;
    double a = pi4 * (double)rem / (double)M;
    double c = std::cos(a), s = std::sin(a);
    const double h = 0.70710678118654752440;
    double co, so;  // cos/sin of o*pi/4
    switch (o & 7) {
        case 0: co = 1; so = 0; break;
        case 1: co = h; so = h; break;
        case 2: co = 0; so = 1; break;
        case 3: co = -h; so = h; break;
        case 4: co = -1; so = 0; break;
        case 5: co = -h; so = -h; break;
        case 6: co = 0; so = -1; break;
        default: co = h; so = -h; break;
    }
    if (rem == 0) return make_double2(co, so);
    return make_double2(co * c - so * s, so * c + co * s);
}

template <int R3>
inline void build_twiddles(std::vector<double2>& tw1, std::vector<double2>& tw2) {
    const int N = 64 * R3, TL = 32;
    tw1.resize(5 * 7 * TL);
    tw2.resize(7 * R3);
    for (int j = 0; j < 5; ++j)
        for (int k1 = 1; k1 < 8; ++k1)
            for (int t = 0; t < TL; ++t)
                tw1[(j * 7 + (k1 - 1)) * TL + t] = unit_root((long long)(t + TL * j) * k1, N);
    for (int k2 = 1; k2 < 8; ++k2)
        for (int n3 = 0; n3 < R3; ++n3) tw2[(k2 - 1) * R3 + n3] = unit_root((long long)n3 * k2, N / 8);
}

}  // namespace psfr
