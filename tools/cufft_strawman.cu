// cuFFT strawman of stage A + stage B (PSD -> structure function -> PSF samples): the GPU pipeline a
// library user would compose - batched cuFFT D2Z / Z2D on the full dim x dim grids with elementwise
// kernels in between - used ONLY as a correctness and performance oracle (BASELINE.json north_star,
// SURVEY F1).  It is never linked into libpsfr_b200.so; tools/cufft_check.py builds it into its own
// shared object, checks it against the CPU oracle and times it next to the product path.
//
// It keeps the algebra of the product (one PSD -> D_unit transform per plane, the telescope
// OTF a constant) so that the comparison is "hand-written pruned, fused kernels" vs "library FFTs on
// full grids", not smart vs naive maths.  The four transforms of the reference it restates:
//   psfrec.py:718  bg = ifft2(fftshift(psd * convnm^2))           -> cufftExecD2Z per plane (stage A)
//   psfrec.py:789  dlFTO = fft2(|ifft2(pupil)|^2)  (two)          -> constant, passed in by the caller
//   psfrec.py:800  psf = Re fftshift(ifft2(fftshift(sysFTO)))     -> cufftExecZ2D per (plane, wavelength)
// All fftshifts cancel between :722 and :797-800 except the final one, which is an index offset of
// the sampled pixels.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC
// tools/cufft_strawman.cu -lcufft -o tools/libcufft_strawman.so
#include <cuda_runtime.h>
#include <cufft.h>
#include <cmath>
#include <cstdio>
#include <vector>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                 \
            return -1;                                                              \
        }                                                                           \
    } while (0)
#define CF(x)                                                                       \
    do {                                                                            \
        cufftResult r = (x);                                                        \
        if (r != CUFFT_SUCCESS) {                                                   \
            fprintf(stderr, "%s: cufft error %d\n", #x, (int)r);                    \
            return -2;                                                              \
        }                                                                           \
    } while (0)

namespace {

// in[y][x] = psd[(y + N/2) % N][(x + N/2) % N]   (the fftshift of psfrec.py:718)
__global__ void shift_kernel(const double* __restrict__ psd, double* __restrict__ out, int N) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= N) return;
    const size_t plane = (size_t)blockIdx.z * N * N;
    out[plane + (size_t)y * N + x] = psd[plane + (size_t)((y + N / 2) % N) * N + (x + N / 2) % N];
}

// D_unit[ky][kx] = 2 (Re B[0][0] - Re B[ky][kx]) / L^2, kx <= N/2 (unshifted half plane).
// ifft2 of a real array = conj(fft2) / N^2 and psfrec.py:718 multiplies by N^2 / L^2.
__global__ void dphi_kernel(const cufftDoubleComplex* __restrict__ B, double* __restrict__ D, int N, double inv_l2) {
    const int NH = N / 2 + 1;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= NH) return;
    const size_t plane = (size_t)blockIdx.z * N * NH;
    D[plane + (size_t)y * NH + x] = 2.0 * (B[plane].x - B[plane + (size_t)y * NH + x].x) * inv_l2;
}

// X[l][ky][kx] = exp(-c_l D_unit) * T   (psfrec.py:793-797), real values in a complex half plane
__global__ void otf_kernel(const double* __restrict__ D, const double* __restrict__ T,
                           const double* __restrict__ clam, cufftDoubleComplex* __restrict__ X, int N) {
    const int NH = N / 2 + 1;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, l = blockIdx.z;
    if (x >= NH) return;
    const size_t i = (size_t)y * NH + x;
    const double t = T[i];
    X[(size_t)l * N * NH + i] = make_double2(t == 0.0 ? 0.0 : exp(-clam[l] * D[i]) * t, 0.0);
}

// psf_muse tail (psfrec.py:672-685) on the 80 x 80 pixels the bilinear resampling reads: crop origin,
// clip >= 0, bilinear point sampling at stride npix / 40, per-plane normalisation.
__global__ void sample_kernel(const double* __restrict__ psf, const int* __restrict__ npix, double* __restrict__ out, int N) {
    __shared__ double red[256];
    const int l = blockIdx.x;
    const double* P = psf + (size_t)l * N * N;
    const int np = npix[l], origin = N / 2 - np / 2;
    double vals[7], sum = 0.0;
    int cnt = 0;
    for (int i = threadIdx.x; i < 1600; i += blockDim.x, ++cnt) {
        const int oy = i / 40, ox = i % 40;
        const double py = (double)(oy * np) / 40, px = (double)(ox * np) / 40;
        const int y0 = (int)floor(py), x0 = (int)floor(px);
        const double fy = py - y0, fx = px - x0;
        auto at = [&](int yy, int xx) {   // centred pixel (yy, xx) of the fftshifted PSF = unshifted index + N/2
            const double v = P[(size_t)((origin + yy + N / 2) % N) * N + (origin + xx + N / 2) % N];
            return v > 0.0 ? v : 0.0;
        };
        double v = (1 - fy) * ((1 - fx) * at(y0, x0) + (fx > 0 ? fx * at(y0, x0 + 1) : 0.0));
        if (fy > 0) v += fy * ((1 - fx) * at(y0 + 1, x0) + (fx > 0 ? fx * at(y0 + 1, x0 + 1) : 0.0));
        vals[cnt] = v;
        sum += v;
    }
    red[threadIdx.x] = sum;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const double tot = red[0];
    cnt = 0;
    for (int i = threadIdx.x; i < 1600; i += blockDim.x, ++cnt) out[(size_t)l * 1600 + i] = vals[cnt] / tot;
}

}  // namespace

extern "C" __attribute__((visibility("default")))
int strawman_run(int N, int nplanes, const double* psd_host, const double* t_half_host, int nlam,
                 const double* lam_nm, double* out_cube_host, int reps, double* ms_stage_a, double* ms_stage_b) {
    const int NH = N / 2 + 1;
    const size_t plane = (size_t)N * N, half = (size_t)N * NH;
    double *d_psd, *d_shift, *d_D, *d_T, *d_clam, *d_psf, *d_out;
    cufftDoubleComplex *d_B, *d_X;
    int* d_npix;
    CK(cudaMalloc(&d_psd, nplanes * plane * 8));
    CK(cudaMalloc(&d_shift, nplanes * plane * 8));
    CK(cudaMalloc(&d_B, nplanes * half * 16));
    CK(cudaMalloc(&d_D, nplanes * half * 8));
    CK(cudaMalloc(&d_T, half * 8));
    CK(cudaMalloc(&d_clam, nlam * 8));
    CK(cudaMalloc(&d_npix, nlam * 4));
    CK(cudaMalloc(&d_X, nlam * half * 16));
    CK(cudaMalloc(&d_psf, nlam * plane * 8));
    CK(cudaMalloc(&d_out, (size_t)nplanes * nlam * 1600 * 8));
    CK(cudaMemcpy(d_psd, psd_host, nplanes * plane * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_T, t_half_host, half * 8, cudaMemcpyHostToDevice));
    std::vector<double> cl(nlam);
    std::vector<int> np(nlam);
    for (int l = 0; l < nlam; ++l) {
        const double conv = 2 * 3.141592653589793 / lam_nm[l];
        cl[l] = 0.5 * conv * conv;                                          // exp(-Dphi/2), Dphi = conv^2 D_unit
        np[l] = (int)(std::nearbyint(((40 * 0.2 * 2 * 8 * 4.85 * 1000) / lam_nm[l]) / 2) * 2);   // psfrec.py:663-664
    }
    CK(cudaMemcpy(d_clam, cl.data(), nlam * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_npix, np.data(), nlam * 4, cudaMemcpyHostToDevice));
    cufftHandle plan_a, plan_b;
    int dims[2] = {N, N};
    CF(cufftPlanMany(&plan_a, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, nplanes));
    CF(cufftPlanMany(&plan_b, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, nlam));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&e2));
    const double L = 16.0;   // 2 * Dpup
    float acc_a = 0.f, acc_b = 0.f;
    for (int r = 0; r < reps + 1; ++r) {   // rep 0 is the warm-up
        CK(cudaEventRecord(e0));
        shift_kernel<<<dim3((N + 255) / 256, N, nplanes), 256>>>(d_psd, d_shift, N);
        CF(cufftExecD2Z(plan_a, d_shift, d_B));
        dphi_kernel<<<dim3((NH + 255) / 256, N, nplanes), 256>>>(d_B, d_D, N, 1.0 / (L * L));
        CK(cudaEventRecord(e1));
        for (int p = 0; p < nplanes; ++p) {
            otf_kernel<<<dim3((NH + 255) / 256, N, nlam), 256>>>(d_D + p * half, d_T, d_clam, d_X, N);
            CF(cufftExecZ2D(plan_b, d_X, d_psf));
            sample_kernel<<<nlam, 256>>>(d_psf, d_npix, d_out + (size_t)p * nlam * 1600, N);
        }
        CK(cudaEventRecord(e2));
        CK(cudaEventSynchronize(e2));
        float a, b;
        CK(cudaEventElapsedTime(&a, e0, e1));
        CK(cudaEventElapsedTime(&b, e1, e2));
        if (r) {
            acc_a += a;
            acc_b += b;
        }
    }
    CK(cudaGetLastError());
    *ms_stage_a = acc_a / reps;
    *ms_stage_b = acc_b / reps;
    CK(cudaMemcpy(out_cube_host, d_out, (size_t)nplanes * nlam * 1600 * 8, cudaMemcpyDeviceToHost));
    cufftDestroy(plan_a);
    cufftDestroy(plan_b);
    cudaFree(d_psd); cudaFree(d_shift); cudaFree(d_B); cudaFree(d_D); cudaFree(d_T); cudaFree(d_clam);
    cudaFree(d_npix); cudaFree(d_X); cudaFree(d_psf); cudaFree(d_out);
    return 0;
}
