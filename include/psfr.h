/* psfr.h - C ABI of the B200-native PSF-reconstruction hot path (libpsfr_b200.so).
 *
 * The reference (musevlt/muse-psfr) has no FFI: the path sits behind plain Python
 * functions in muse_psfr/psfrec.py.  Each entry point below cites the reference
 * function it replaces; the Python host (muse_psfr_b200/psfrec.py) binds them with
 * ctypes and keeps the reference's signatures.  INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative PSFR_E_* code on failure and
 *    never throws; psfr_last_error() gives the message of the last failure;
 *  - all arrays are C-contiguous FP64 unless stated; every data pointer may be a HOST
 *    pointer (pageable or pinned) or a DEVICE pointer on the context's GPU - the
 *    library detects which and stages copies on the given stream;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all
 *    work is stream-ordered, host outputs are complete when the call returns,
 *    device outputs when the stream reaches that point;
 *  - one context per GPU, not thread-safe; the context owns twiddles, the telescope
 *    OTF, geometry tables and all workspaces;
 *  - there is no CPU fallback: without a CUDA device psfr_create fails.
 */
#ifndef PSFR_H
#define PSFR_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PSFR_API __attribute__((visibility("default")))
#else
#define PSFR_API
#endif

typedef struct psfr_ctx psfr_ctx;

enum {
    PSFR_OK = 0,
    PSFR_E_CUDA = -1,      /* CUDA runtime error (message in psfr_last_error) */
    PSFR_E_ARG = -2,       /* invalid argument */
    PSFR_E_UNSUPPORTED = -3, /* e.g. dim other than 1280 / 2560, more than PSFR_MAX_LAYERS layers */
    PSFR_E_CAPACITY = -4,  /* batch larger than the context was created for */
    PSFR_E_STATE = -5      /* call order (e.g. geometry not set) */
};

/* Size of the AO-corrected zone (Dimpup*2, psfrec.py:103,138) and of the final PSF. */
#define PSFR_AO_DIM 80
#define PSFR_PSF_DIM 40

/* Per-draw parameter record (array of doubles, one row per draw).
 * Filled by the host exactly as simul_psd_wfm / convolve_final_psf derive them. */
#define PSFR_MAX_LAYERS 8    /* the reference works for 1 or 2 layers only (wind directions are a hard-coded
                              * 2-vector, psfrec.py:66,594); more need explicit wind directions from the host */
enum {
    PSFR_DRAW_R0 = 0,        /* r0 at 0.5 um on the line of sight (seeing2r01, psfrec.py:108,183-187; holds the zenith angle) */
    PSFR_DRAW_L0 = 1,        /* outer scale [m] */
    PSFR_DRAW_FITC = 2,      /* cst r0^(-5/3) of the fitting PSD (psfrec.py:622-625) */
    PSFR_DRAW_ALPHA_TT = 3,  /* Moffat alpha of the tip-tilt kernel [px] (psfrec.py:881-905) */
    PSFR_DRAW_NLAYERS = 4,   /* 1 .. PSFR_MAX_LAYERS */
    PSFR_DRAW_LAYER0 = 8,    /* layer l occupies PSFR_DRAW_LAYER0 + PSFR_LAYER_NPAR * l + PSFR_LAYER_* */
    PSFR_DRAW_NPAR = 40
};
enum {
    PSFR_LAYER_CPHI = 0,     /* 0.0229 (Cn2_l^(-3/5) r0)^(-5/3) (psfrec.py:569-571) */
    PSFR_LAYER_H = 1,        /* altitude [m] */
    PSFR_LAYER_WX = 2,       /* wind vector [m/s] (psfrec.py:61,66,594) */
    PSFR_LAYER_WY = 3,
    PSFR_LAYER_NPAR = 4
};

/* Per-image fit record written by the Moffat fitter (mpdaf Image.moffat_fit columns,
 * psfrec.py:861-871; fwhm still in pixels).  The err_* entries are sqrt(|diag((J^T J)^-1)| chi^2/dof), the
 * estimate scipy.optimize.leastsq's cov_x gives mpdaf. */
enum {
    PSFR_FIT_PEAK = 0,   /* I */
    PSFR_FIT_Y0 = 1,     /* centre, first (row) axis */
    PSFR_FIT_X0 = 2,     /* centre, second (column) axis */
    PSFR_FIT_ALPHA = 3,  /* a */
    PSFR_FIT_N = 4,      /* beta */
    PSFR_FIT_FWHM = 5,   /* 2 a sqrt(2^(1/n)-1) [px] */
    PSFR_FIT_CHISQ = 6,
    PSFR_FIT_ITER = 7,   /* LM iterations used; negative = not converged */
    PSFR_FIT_ERR_PEAK = 8, PSFR_FIT_ERR_Y0 = 9, PSFR_FIT_ERR_X0 = 10,
    PSFR_FIT_ERR_ALPHA = 11, PSFR_FIT_ERR_N = 12,
    PSFR_FIT_ERR_FWHM = 13,  /* mpdaf's expression: err_a * n [px] */
    PSFR_FIT_FLUX = 14,      /* pi a^2 I / (n - 1) */
    PSFR_FIT_ERR_FLUX = 15,  /* mpdaf's expression err_I err_n err_a^2 err_e with err_e = 0 (circular fit) */
    PSFR_FIT_NPAR = 16
};

/* Context ------------------------------------------------------------------------- */

/* dim: PSD grid size (1280; the reference hard-codes it in compute_psf, psfrec.py:955).
 * max_planes: largest number of (draw x direction) planes processed at once;
 * max_lambda: largest number of wavelengths per call.  Builds twiddles, the pupil
 * (pupil_mask, psfrec.py:190-203,656) and the telescope OTF (psfrec.py:784-790) on the GPU. */
PSFR_API int psfr_create(int device, int dim, int max_planes, int max_lambda, psfr_ctx** out);
PSFR_API void psfr_destroy(psfr_ctx* ctx);
PSFR_API const char* psfr_last_error(const psfr_ctx* ctx);   /* ctx may be NULL: last create error */
PSFR_API int psfr_version(void);

/* AO-zone frequency tables f, f_x = f cos(arg f), f_y = f sin(arg f) on the 80x80 grid,
 * formed by the host exactly as psfrec.py:548-554,241-242 (the cutoff masks depend on
 * their 1-ulp rounding). */
PSFR_API int psfr_set_geometry(psfr_ctx* ctx, const double* f, const double* f_x, const double* f_y);

/* Stages (each mirrors one reference function) -------------------------------------- */

/* simul_psd_wfm (psfrec.py:36-151) incl. dsp4muse / calc_mat_rec_glao_finale /
 * calc_dsp_res_glao_finale / psd_fit.  Planes are ordered draw-major: plane = draw*ndir+dir.
 * dirs[2*ndir]: field directions [arcsec] (direction_perf); poslgs[2*ngs]: LGS positions
 * [arcsec], x then y.  The PSD stays in the context workspace; if out_psd != NULL it is
 * also copied there ([ndraw*ndir][dim][dim], nm^2). */
PSFR_API int psfr_psd(psfr_ctx* ctx, int ndraw, const double* draws, int ndir, const double* dirs,
             int ngs, const double* poslgs, double* out_psd, void* stream);

/* Load user PSDs into the workspace instead (psf_muse / psd_to_psf called on an array). */
PSFR_API int psfr_load_psd(psfr_ctx* ctx, int nplanes, const double* psd, void* stream);

/* PSD -> wavelength-free structure function D_unit on the workspace planes
 * (psfrec.py:717-722 with the (2 pi/lambda)^2 factor left out). */
PSFR_API int psfr_structure_function(psfr_ctx* ctx, int nplanes, void* stream);

/* psd_to_psf (psfrec.py:689-807, live branch): full dim x dim PSF of workspace plane
 * `plane` at wavelength lambda_m [m], normalised to unit sum.  Parity mode. */
PSFR_API int psfr_psd_to_psf(psfr_ctx* ctx, int plane, double lambda_m, double* out_psf, void* stream);

/* psf_muse (psfrec.py:644-686): 40x40 PSFs at 0.2"/px for ndraw draws x nlam wavelengths,
 * averaging the ndir planes of each draw; pruned transform (only the 80x80 samples the
 * bilinear resampling reads).  out_cube: [ndraw][nlam][40][40]. */
PSFR_API int psfr_psf_cube(psfr_ctx* ctx, int ndraw, int ndir, int nlam, const double* lambda_nm,
                  double* out_cube, void* stream);

/* convolve_final_psf (psfrec.py:874-930): tip-tilt Moffat (beta=2, alpha_tt[draw] in px)
 * then MUSE intrinsic Moffat (muse_intrinsic_psf, psfrec.py:1144-1171) per wavelength.
 * cube [ndraw][nlam][40][40] in -> out (may alias). */
PSFR_API int psfr_convolve(psfr_ctx* ctx, int ndraw, int nlam, const double* lambda_nm,
                  const double* alpha_tt, const double* cube, double* out_cube, void* stream);

/* fit_psf_cube (psfrec.py:861-871 -> mpdaf Image.moffat_fit, circular, no background):
 * batched Levenberg-Marquardt, one CTA per image.  imgs [nimg][ny][nx], params
 * [nimg][PSFR_FIT_NPAR]. */
PSFR_API int psfr_moffat_fit(psfr_ctx* ctx, int nimg, int ny, int nx, const double* imgs,
                    double* params, void* stream);

/* compute_psf (psfrec.py:933-978) for a batch of draws, fused on the device:
 * PSD -> structure function -> pruned PSFs -> resample -> convolutions -> fit.
 * out_cube [ndraw][nlam][40][40] and out_fit [ndraw][nlam][PSFR_FIT_NPAR]; either may be NULL. */
PSFR_API int psfr_compute_batch(psfr_ctx* ctx, int ndraw, const double* draws, int ndir, const double* dirs,
                       int ngs, const double* poslgs, int nlam, const double* lambda_nm,
                       double* out_cube, double* out_fit, void* stream);

/* Time-mean of ncube cubes + refit (psfrec.py:1104-1105). cubes [ncube][nlam][40][40]. */
PSFR_API int psfr_mean_refit(psfr_ctx* ctx, int ncube, int nlam, const double* cubes,
                    double* out_mean, double* out_fit, void* stream);

/* fit_psf_with_polynom (psfrec.py:1174-1210): least-squares polynomials of degree `deg`
 * in the normalised wavelength for nseries series of nlam values each (one shared design
 * matrix).  y [nseries][nlam] -> coef [nseries][deg+1], highest power first (np.polyfit). */
PSFR_API int psfr_polyfit(psfr_ctx* ctx, int nseries, int nlam, const double* lambda_nm, int deg,
                 const double* y, double* coef, void* stream);

/* Options.  PSFR_OPT_EXP_CUT (default 45): in the pruned stage-B row pass (psfr_psf_cube,
 * psfr_compute_batch) entries of exp(-Dphi/2) smaller than exp(-cut) are flushed to zero and
 * row pairs that are below the cut everywhere are not transformed.  The OTF peak is 1, so the
 * default drops terms below 2.9e-20 of it - four orders of magnitude under the FP64 rounding of a
 * single kept term; beyond the cut radius exp(-Dphi/2) falls off faster than exponentially, so even
 * a coherent sum of everything dropped stays below ~1e-15 of the OTF peak, against a PSF peak of
 * 15 (2 arcsec seeing) to several 1000.  tools/parity_sweep.py holds the default against the oracle
 * over the BASELINE configs and the corners of the config-4 sweep (profiles/).  A value >= 745 (exp
 * underflows) disables the cut; psfr_psd_to_psf (full-grid parity mode) never applies it (nor the
 * grades below).
 *
 * Graded precision of the same pass (the OTF peak is exactly 1, so an entry's size bounds
 * what an error in it can do to the result):
 * PSFR_OPT_EXP_GRADE (default 20): a segment of 2 x 32 OTF entries whose every live entry is
 * below exp(-grade) = 2.1e-9 evaluates exp on the special-function unit in single precision
 * (relative error ~4e-6, i.e. < 1e-14 of the peak per entry); other blocks use the FP64 exp.
 * PSFR_OPT_F32_ROWS (default 25, dim 1280 only; dim 2560 runs every unit in FP64): a row pair whose every entry is below
 * exp(-thr) = 1.4e-11 is evaluated AND transformed in single precision; all other rows and
 * the whole column pass are FP64.
 * Measured against the all-FP64 evaluation (tools/diag_grade.py, seven seeing/L0 cases x five
 * wavelengths): largest difference 1.5e-14 of the PSF peak, 1.1e-11 pointwise on pixels above
 * 1e-6 of the peak - the all-FP64 kernel and numpy differ by 1e-10 there.  A threshold >= the
 * cut disables the respective grade (e.g. 1e30). */
enum { PSFR_OPT_EXP_CUT = 1, PSFR_OPT_EXP_GRADE = 2, PSFR_OPT_F32_ROWS = 3,
       /* row kernel of the pruned stage B: 2 (default) one 128-thread group per row transform, data in
        * shared memory, pruned third pass (csrc/psfr_hot2.cu); 1 one warp per transform, data in
        * registers (csrc/psfr_hot.cu) */
       PSFR_OPT_ROW_KERNEL = 4 };
PSFR_API int psfr_set_option(psfr_ctx* ctx, int key, double value);

/* Introspection used by tests and the bench --------------------------------------- */
/* numeric properties of a context: option values, capacities, and the number of sampled row-pass
 * frequencies kept per PSF in the hand-off buffer between the two passes of the pruned stage B */
enum { PSFR_INFO_Y_COLS = 1, PSFR_INFO_EXP_CUT = 2, PSFR_INFO_EXP_GRADE = 3, PSFR_INFO_F32_ROWS = 4,
       PSFR_INFO_ROW_KERNEL = 5, PSFR_INFO_MAX_PLANES = 6, PSFR_INFO_MAX_LAMBDA = 7, PSFR_INFO_DIM = 8 };
PSFR_API int psfr_get_info(const psfr_ctx* ctx, int key, double* out);
PSFR_API int psfr_get_otf(psfr_ctx* ctx, double* out);            /* [dim/2+2][dim] half-plane telescope OTF */
PSFR_API int psfr_get_structure_function(psfr_ctx* ctx, int plane, double* out); /* [dim/2+2][dim], transposed half-plane */
/* test hook: y[i] = the device exp() used for exp(-Dphi/2) (csrc/fast_exp.cuh), x[i] <= 0 */
PSFR_API int psfr_debug_exp(psfr_ctx* ctx, int n, const double* x, double* y);
PSFR_API long long psfr_kernel_launches(const psfr_ctx* ctx);     /* kernels launched by this context so far */
/* device-side duration [ms] of the stage-B row kernel in the last psfr_psf_cube /
 * psfr_compute_batch call (CUDA events on the launching stream), and its launch count */
PSFR_API int psfr_last_hot_timing(psfr_ctx* ctx, double* ms, int* launches, long long* psfs);

#ifdef __cplusplus
}
#endif
#endif /* PSFR_H */
