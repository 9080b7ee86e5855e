/*
 * fastb.h -- C ABI of libfastb.so: the B200 (sm_100a) implementation of the Monte-Carlo hot
 * path of FAST (ojdf/fast).  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference is pure Python and has no FFI of its own, so each entry point below names the
 * reference function(s) it replaces (paths relative to the reference repository); the
 * reference-side ctypes binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every `d_*` argument is a DEVICE pointer owned by the caller (e.g. a torch tensor); the
 *     library never frees or retains it.  `stream` is a cudaStream_t passed as void* (NULL =
 *     legacy default stream).  All work is enqueued asynchronously on that stream.
 *   - return value 0 = success; non-zero = FASTB_ERR_*; fastb_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread.
 *   - grids are row-major N x N: index r*N + c, frequency fx = (c - N/2) df, fy = (r - N/2) df
 *     (numpy.meshgrid layout of fast/fast.py:830-833,911).
 *   - there is no CPU fallback: every compute entry point fails with FASTB_ERR_CUDA when no
 *     CUDA device is usable.
 */
#ifndef FASTB_H
#define FASTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FASTB_VERSION 200          /* major*10000 + minor*100 + patch */
#define FASTB_MAX_LAYERS 32

enum {
    FASTB_OK = 0,
    FASTB_ERR_ARG = 1,             /* invalid argument (message says which) */
    FASTB_ERR_CUDA = 2,            /* CUDA runtime / launch failure, or no device */
    FASTB_ERR_UNSUPPORTED = 3      /* valid request this build cannot serve (e.g. grid size) */
};

/* AO modes of ao_power_spectra.G_AO_PAOLA (fast/ao_power_spectra.py:225-270) */
enum { FASTB_AO_NOAO = 0, FASTB_AO_AO = 1 /* 'AO' and 'TT' */, FASTB_AO_LGSAO = 2 };

/* ---------------------------------------------------------------------------------------
 * K1: residual phase PSD build.
 * Replaces Fast.compute_powerspec (fast/fast.py:445-492) and the functions it calls:
 *   funcs.turb_powerspectrum_vonKarman (fast/funcs.py:138-173),
 *   ao_power_spectra.mask_lf zonal branch (fast/ao_power_spectra.py:119-141),
 *   ao_power_spectra.G_AO_PAOLA (:225-270), Jol_alias_openloop (:163-223),
 *   Jol_noise_openloop (:148-161), logamp_powerspec (:272-301),
 *   and the frequency grids of SpatialFrequencies.make_main_freqs (fast/fast.py:830-833).
 * All arithmetic is float64, evaluated in the reference's operation order.
 * ------------------------------------------------------------------------------------- */
typedef struct FastbPsdParams {
    int32_t n;                      /* grid size N >= 2; the zero frequency sits at index N/2
                                       (integer division).  Main grids are even; the 3 x 3
                                       sub-harmonic levels (fast/fast.py:835-844) use N = 3 */
    int32_t n_layers;               /* L, 1..FASTB_MAX_LAYERS */
    int32_t ao_mode;                /* FASTB_AO_* */
    int32_t alias;                  /* 1: include WFS aliasing (ignored for NOAO) */
    int32_t lmax, kmax;             /* alias replicas, reference uses 5, 5 (fast/fast.py:462) */
    int32_t reserved0, reserved1;
    double df;                      /* 2 pi / (N dx)  [rad/m] */
    double k;                       /* 2 pi / wavelength */
    double wvl;                     /* wavelength [m] */
    double L0, l0;                  /* outer (may be +inf) / inner scale [m] */
    double dsubap;                  /* WFS sub-aperture pitch [m] */
    double tloop, texp;             /* loop delay, WFS exposure [s] */
    double noise_var;               /* NOISE (<= 0: no noise term) */
    double dtheta[2];               /* point-ahead (x, y) [arcsec] */
    double h[FASTB_MAX_LAYERS];     /* zenith-corrected layer heights (fast/fast.py:234) */
    double cn2[FASTB_MAX_LAYERS];   /* zenith-corrected Cn2 dh (fast/fast.py:235) */
    double vx[FASTB_MAX_LAYERS];    /* wind_vector[:,0] (fast/fast.py:253-257) */
    double vy[FASTB_MAX_LAYERS];    /* wind_vector[:,1] */
} FastbPsdParams;

/* Optional per-pixel inputs (NULL = not used). */
typedef struct FastbPsdInputs {
    const double* d_lf_mask;        /* N*N low-frequency (corrected-region) mask; NULL = zonal
                                       box |fx|,|fy| <= pi/dsubap computed on device.  Modal /
                                       TT masks (Bessel based) are supplied by the host. */
    const double* d_zfilter;        /* N*N Zernike(1..4) squared filter, required for LGSAO */
    const double* d_pupil_filter;   /* N*N |FT(P M)|^2 / (sum P M)^2, required for d_logamp */
} FastbPsdInputs;

/* Outputs; any pointer may be NULL except d_powerspec. */
typedef struct FastbPsdOutputs {
    double* d_powerspec;            /* N*N    sum_l powerspec_per_layer        (fast.py:481) */
    double* d_powerspec_per_layer;  /* L*N*N  2 pi k^2 (turb G + alias) + noise/L   (:478) */
    double* d_turb;                 /* L*N*N  von Karman per layer                   (:448) */
    double* d_g_ao;                 /* L*N*N  aniso-servo transfer function          (:451) */
    double* d_alias;                /* L*N*N  aliasing PSD per layer                 (:460) */
    double* d_noise;                /* N*N    WFS noise PSD                          (:471) */
    double* d_logamp;               /* N*N    aperture-filtered log-amplitude PSD    (:490) */
    double* d_integrands;           /* 3*N*N  [sum_l(G turb) M 2pi k^2, sum_l alias 2pi k^2,
                                               W (1-M)]: integrands of aniso_servo_error,
                                               alias_error, fitting_error (:456,464,483) */
    float*  d_weight;               /* N*N    (-1)^(r+c) sqrt(W) df : screen-synthesis weight,
                                               the input of fastb_screen_detect */
    float*  d_weight_per_layer;     /* L*N*N  same per layer (TEMPORAL mode, fast.py:611-612) */
} FastbPsdOutputs;

int fastb_psd_build(const FastbPsdParams* p, const FastbPsdInputs* in,
                    const FastbPsdOutputs* out, void* stream);

/* Modal corrected-region masks and Fourier-space Zernike filters, float64; replaces
 * ao_power_spectra.zernike_ft / zernike_squared_filter / mask_lf(modal=True)
 * (fast/ao_power_spectra.py:10-141; used by fast/fast.py:311-319 and G_AO_PAOLA's LGSAO branch
 * :262-267).  On the n x n frequency grid f = (index - n/2) df:
 *   noll_first <= noll_last : v = sum_j |Z_j(f)|^2 over the Noll range (disc of `diameter`),
 *                             DC pixel := 1 if noll_first == 1 else 0;  gtilt: v += J_1(|f| D/2)^2
 *   noll_last  <  noll_first: v = [ |f| <= modal_mult pi / d_wfs ]       (modal DM, no Zernike limit)
 *   clip_box                : out = [ |fx|,|fy| <= pi/d_wfs ] * min(v, 1)   (mask_lf) else out = v.
 * The outputs feed FastbPsdInputs.d_lf_mask / d_zfilter. */
typedef struct FastbZernikeParams {
    int32_t n, noll_first, noll_last, gtilt, clip_box, reserved;
    double df, diameter, d_wfs, modal_mult;
} FastbZernikeParams;
int fastb_zernike_filter(const FastbZernikeParams* p, double* d_out, void* stream);

/* d_weight[r*N+c] = (-1)^(r+c) * sqrt(d_W[r*N+c]) * df, for `batch` stacked N x N spectra.
 * Replaces `rand *= numpy.sqrt(self.powerspec)` + `rand * df` (fast/fast.py:594,
 * fast/funcs.py:218) for callers that bring their own PSD. */
int fastb_make_weight(const double* d_W, int32_t n, int32_t batch, double df, float* d_weight,
                      void* stream);

/* out[b] = w^T P_b w for b < batch: scipy.integrate.simpson twice (fast/funcs.py:100-115),
 * with the 1-D weight vector w (length n) supplied by the host so that the end correction
 * follows the installed scipy. */
int fastb_simpson2d(const double* d_P, int32_t n, int32_t batch, const double* d_w,
                    double* d_out, void* stream);

/* d_pf = |centred DFT2(d_pm)|^2 / (sum d_pm)^2 on the full N x N grid, float64, any N.
 * Replaces funcs.pupil_filter(spline=False) (fast/funcs.py:308-315 with aotools.ft2). */
int fastb_pupil_filter(const double* d_pm, int32_t n, double* d_pf, void* d_workspace,
                       int64_t workspace_bytes, void* stream);
int64_t fastb_pupil_filter_workspace_bytes(int32_t n);

/* ---------------------------------------------------------------------------------------
 * K2: phase-screen generation fused with the detector.
 * Replaces, per complex transform ("pair" = two realisations, fast/funcs.py:220-221):
 *   funcs.generate_random_coefficients (fast/funcs.py:352-356)      [device RNG mode]
 *   Fast.compute_phs (fast/fast.py:589-596) incl. funcs.make_phase_fft (fast/funcs.py:210-223)
 *   Fast.compute_logamp (fast/fast.py:639-645)                      [device RNG mode]
 *   Fast.compute_detector (fast/fast.py:647-668)
 * Per pair g: S = noise * |weight|, Phi = centred inverse DFT2 (N^2-scaled) of S restricted
 * to rows/cols [lo, lo+n_pup); phi_a = Re Phi, phi_b = Im Phi;
 *   z = exp(chi) * sum(U exp(i phi)) / sum(U);  result = |z|^2, or z when coherent.
 * Only the per-realisation scalars are written to HBM.
 *
 * Realisation numbering (Fast.run layout, fast/fast.py:117-134): pair g belongs to chunk
 * g / pairs_per_chunk; its Re realisation has global index
 *   i_a = (g / ppc) * 2 ppc + g % ppc   and the Im realisation  i_b = i_a + ppc.
 * chi is indexed by that global index.
 *
 * Device RNG (d_noise == NULL): Philox4x32-10, key = (seed lo, seed hi).  Noise block
 * b = r*S + t holds the 16 cells (r, t + S m), m < 16 (cells beyond the grid are not used).  The noise
 * stride S = fastb_noise_stride(N, n_pup) is the number of threads that serve a line in the kernel that
 * owns the grid, so that a thread's noise block is its transform input:
 *   N = 64..2048 power of two                          S = N / 16
 *   other even N with N + n_pup - 1 <= 2048 (chirp-z)   S = M / 16, M = 2^ceil(log2(N + n_pup - 1)) >= 64
 *   anything else (direct DFT)                         S = ceil(N / 16)
 * and does not depend on the FASTB_ALGO_* requested.  A block is fed by six calls
 * q < 6 with counter (b, g lo, g hi, 0x5CE7E000 + q): 24 words W[4q+j].  Word triple G < 8
 * (W[3G], W[3G+1], W[3G+2]) gives four 23-bit fields -- the top 23 bits of each word and one
 * field mixed from their low 9 bits -- so 24 words feed 16 Box-Muller pairs:
 *   cell m = 2G  : radius field W[3G] >> 9,   angle field W[3G+1] >> 9
 *   cell m = 2G+1: radius field W[3G+2] >> 9, angle field (W[3G]&511)<<14 | (W[3G+1]&511)<<5 | (W[3G+2]>>4)&31
 *   radius = sqrt(-2 ln(1 - field 2^-23)),  angle = 2 pi field 2^-23,
 *   Re = radius cos(angle), Im = radius sin(angle).
 * 'device-fast' stream (FASTB_RUN_RNG_FAST): same block / cell mapping, five Philox4x32-7 calls
 * q < 5 with counter (b, g lo, g hi, 0x5CE7F000 + q): 20 words W[4q+j].  Cell m < 16 owns word W[m]
 * and byte m % 4 of the extra word W[16 + m/4]:
 *   radius field = W[m] & 0x7FFFFF,   angle field = ((W[m] >> 9) & 0x7FC000) ^ (byte << 8)
 * (40 random bits per complex sample, angle on a 2^-15 turn lattice), same Box-Muller.
 * chi_i = sigma_chi * n_i, n_i = normal (i % 4) of call i / 4 with counter
 * (i/4 lo, i/4 hi, 0, 0x10CA3900), Box-Muller on the top 23 bits of word pairs (0,1), (2,3).  Results depend only on (seed, g), never on the launch
 * geometry or the number of GPUs.
 * ------------------------------------------------------------------------------------- */
enum {
    FASTB_ALGO_AUTO = 0,
    FASTB_ALGO_DIRECT = 1,          /* pruned direct DFT, any even N */
    FASTB_ALGO_RADIX = 2,           /* register radix-16 FFT, one line per thread group, N = 64..2048
                                       power of two; what AUTO selects for those sizes */
    FASTB_ALGO_RADIX_PAIR = 3,      /* same FFT on two adjacent lines per thread group, all arithmetic
                                       in packed FP32 (FADD2/FMUL2/FFMA2): fewer instructions, but
                                       measured slower than RADIX on B200; kept as a cross-check */
    FASTB_ALGO_BLUESTEIN = 4        /* any even N with N + n_pup - 1 <= 2048 that is not a RADIX size: chirp-z
                                       (Bluestein) on the radix line FFT of length
                                       M = 2^ceil(log2(N + n_pup - 1)); what AUTO selects for grids that are not
                                       a power of two (the reference's NPXLS 'auto' sizes,
                                       fast/fast.py:166-187, e.g. 164) */
};

/* FastbRunParams.flags */
#define FASTB_RUN_PREPARED 1        /* the workspace tables were filled by fastb_screen_detect_prepare for
                                       the same (params, weight, U); skip the per-call preparation kernels */
#define FASTB_RUN_RNG_FAST 2        /* opt-in 'device-fast' noise stream: Philox4x32-7, five calls per block
                                       of 16 cells (see below); statistically equivalent, not bit-compatible
                                       with the default stream */

typedef struct FastbRunParams {
    int32_t n;                      /* N */
    int32_t n_pup;                  /* Npxls_pup (fast/fast.py:211) */
    int32_t lo;                     /* (N - n_pup) / 2 (fast/fast.py:390) */
    int32_t coherent;               /* 0: |z|^2 -> 1 float; 1: z -> 2 floats (re, im) */
    int32_t algo;                   /* FASTB_ALGO_* */
    int32_t flags;                  /* FASTB_RUN_* bits, 0 by default */
    int64_t n_pairs;                /* pairs in this call */
    int64_t first_pair;             /* global index g of the first pair */
    int64_t pairs_per_chunk;        /* NITER/NCHUNKS/2; > 0 */
    uint64_t seed;
    double u_sum;                   /* sum(U) over the n_pup x n_pup crop */
    float sigma_chi;                /* sqrt(logamp_var); used when d_chi == NULL */
    float reserved_f;
} FastbRunParams;

/* Optional sub-harmonic correction (funcs.make_phase_subharm, fast/funcs.py:225-258, added to the
 * cropped screens at fast/fast.py:598-603).  27 plane waves: level i < 3, fy index q < 3, fx index
 * s < 3 (flat index m = (i*3 + q)*3 + s), frequencies (s-1, q-1) * df_i, df_i = 2 pi/(3^(i+1) N dx).
 * Per pair the complex amplitudes are noise_m * d_weight[m]; the complex screen
 *   sum_m amp_m exp(i (x fx_m + y fy_m)) - (its mean over the full N x N grid)
 * is added to Phi before the detector (Re -> realisation a, Im -> realisation b).
 * Device RNG: call j < 14 with counter (j, g lo, g hi, 0x5AB4A200) yields amplitudes m = 2j
 * (words 0,1) and m = 2j+1 (words 2,3) by the same Box-Muller as the main noise. */
typedef struct FastbSubharm {
    const float* d_weight;          /* 27 floats: sqrt(powerspec_subharm[i,q,s]) * df_i */
    const float* d_noise;           /* NULL (device RNG) or n_pairs*27 complex64: the reference's
                                       `rand_lo` (fast/fast.py:600) cast to complex64 */
    const float* d_ex;              /* 3*n_pup complex64: exp(i x_c df_i), x_c = pupil column coords */
    const float* d_ey;              /* 3*n_pup complex64: exp(i y_r df_i), y_r = pupil row coords */
    const float* d_mean;            /* 27 complex64: full-grid mean of each plane wave */
} FastbSubharm;

int64_t fastb_screen_detect_workspace_bytes(const FastbRunParams* p);

/* d_weight  N*N floats from fastb_psd_build / fastb_make_weight
 * d_U       n_pup*n_pup floats, pupil * pupil_mode crop (fast/fast.py:649)
 * d_chi     NULL (device RNG) or log-amplitudes indexed by GLOBAL realisation index
 * d_noise   NULL (device RNG) or n_pairs*N*N complex64 (re, im interleaved) unit normals,
 *           pair-major: the reference's `rand` array (fast/fast.py:593) cast to complex64
 * d_out_a   results of the Re realisations, element (g - first_pair) [x2 floats if coherent]
 * d_out_b   results of the Im realisations, same indexing
 * sh        NULL, or the sub-harmonic term (SUBHARM=True)
 */
int fastb_screen_detect(const FastbRunParams* p, const float* d_weight, const float* d_U,
                        const float* d_chi, const float* d_noise, const FastbSubharm* sh,
                        float* d_out_a, float* d_out_b, void* d_workspace,
                        int64_t workspace_bytes, void* stream);

/* ---- batched configurations and fused statistics ------------------------------------------------
 * fastb_screen_detect_batch runs ONE launch over `n_items` configurations that share the grid, the
 * crop and U but have their own weight table, sigma_chi and seed -- the elevation samples of a
 * satellite pass that the reference builds one Fast object at a time
 * (fast/complete_orbit_simulation.py:217-228).  Pairs are addressed by the flattened index
 *   q = item * pairs_per_item + g,       g = pair index inside its configuration (RNG counter),
 * FastbRunParams.first_pair / n_pairs select a range of q (so that (item x pair) ranges shard over
 * GPUs), outputs are indexed by q - first_pair, and every item's values are bit-identical to a
 * single-configuration run with that item's weight, sigma_chi and seed.  batch == NULL: a single
 * configuration (d_weight one table, seed / sigma_chi from FastbRunParams).
 *   d_weight   n_items * N*N floats
 * Fused K3 (stats != NULL): moments, extrema and the dB histogram of the results of this launch are
 * accumulated into per-item buffers while the scalars are written -- the same quantities, bin rule
 * and accumulate-into semantics as fastb_stats, without a second pass:
 *   d_sums [n_items][8], d_minmax [n_items][2], d_hist [n_items][nbins + 2].
 * fastb_screen_detect_prepare fills the derived tables of the workspace (transposed U, pre-scaled
 * interleaved weight copies or chirp tables); with FASTB_RUN_PREPARED set the run entry points skip
 * their own preparation kernels, so a step is a single launch.  The tables depend on
 * (n, n_pup, lo, algo, n_items, d_weight contents, d_U contents). */
typedef struct FastbRunBatch {
    int32_t n_items;                /* >= 1 */
    int32_t reserved;
    int64_t pairs_per_item;         /* NITER / 2 of each configuration */
    const float* d_sigma_chi;       /* n_items floats */
    const uint64_t* d_seeds;        /* n_items seeds */
} FastbRunBatch;

typedef struct FastbRunStats {
    double db_lo, db_hi;            /* histogram range [db_lo, db_hi) in dB */
    int32_t nbins, reserved;
    double* d_sums;                 /* [n_items][8]  += { n, sum r, sum r^2, sum dB, sum dB^2, n(r <= 0), 0, 0 } */
    double* d_minmax;               /* [n_items][2]   = { min(old, min r), max(old, max r) } */
    unsigned long long* d_hist;     /* [n_items][nbins + 2], under/overflow in the last two */
} FastbRunStats;

int64_t fastb_screen_detect_batch_workspace_bytes(const FastbRunParams* p, int32_t n_items);
int fastb_screen_detect_prepare(const FastbRunParams* p, int32_t n_items, const float* d_weight,
                                const float* d_U, void* d_workspace, int64_t workspace_bytes, void* stream);
int fastb_screen_detect_batch(const FastbRunParams* p, const FastbRunBatch* batch,
                              const FastbRunStats* stats, const float* d_weight, const float* d_U,
                              const float* d_chi, float* d_out_a, float* d_out_b, void* d_workspace,
                              int64_t workspace_bytes, void* stream);

/* Verification / inspection seam: the cropped phase screens themselves, as the reference keeps
 * them in `Fast.phs` (fast/fast.py:596, shape (J, n_pup, n_pup)).  Same inputs and numbering as
 * fastb_screen_detect; d_phs receives, for pair index p of this call, the Re screen at
 * [2p] and the Im screen at [2p+1], each n_pup*n_pup floats, row-major.  Slow path (direct
 * DFT), any even N <= 4096; the Monte-Carlo run never materialises screens. */
int fastb_screens_crop(const FastbRunParams* p, const float* d_weight, const float* d_noise,
                       const FastbSubharm* sh, float* d_phs, void* d_workspace,
                       int64_t workspace_bytes, void* stream);

/* Debug / verification aid: materialise the device-RNG noise tile of pair g (N*N complex64)
 * and the chi normals of realisations [first, first+count).  Either output may be NULL. */
int fastb_rng_dump(uint64_t seed, int64_t pair, int32_t n, float* d_noise_tile,
                   int64_t chi_first, int64_t chi_count, float* d_chi_normals, void* stream);
/* same with the noise stream selected: rng_fast = 0 default stream, 1 'device-fast'.  Both use the
 * stride ceil(n / 16), which is the K2 stride for powers of two and the stride of the K4 layer screens. */
int fastb_rng_dump_mode(uint64_t seed, int64_t pair, int32_t n, int32_t rng_fast, float* d_noise_tile,
                        int64_t chi_first, int64_t chi_count, float* d_chi_normals, void* stream);
/* noise stride S of the K2 device RNG for an (n, n_pup) problem (see the K2 contract above); -1 for
 * arguments fastb_screen_detect would reject */
int32_t fastb_noise_stride(int32_t n, int32_t n_pup);
/* the tile for an explicit stride (16 stride >= n) */
int fastb_rng_dump_stride(uint64_t seed, int64_t pair, int32_t n, int32_t stride, int32_t rng_fast,
                          float* d_noise_tile, int64_t chi_first, int64_t chi_count, float* d_chi_normals,
                          void* stream);

/* ---------------------------------------------------------------------------------------
 * K4: TEMPORAL (frozen-flow) mode -- Fast.compute_phs_temporal (fast/fast.py:607-637).
 *
 * K4a fastb_layer_screens: the chunk-0 branch (fast/fast.py:609-614): one full-size REAL screen
 * per turbulence layer, screen_l = Re centred-inverse-DFT2(noise_l * sqrt(powerspec_per_layer_l) df)
 * (funcs.make_phase_fft(double=False), fast/funcs.py:210-223).
 *   d_weight_per_layer  L*N*N signed weights from fastb_psd_build
 *   d_noise             NULL (device Philox; layer l uses pair index FASTB_LAYER_PAIR_BASE + l,
 *                       same cell mapping as fastb_screen_detect) or L*N*N complex64
 *   d_screens           L*N*N float
 * Any even N <= 4096: pruning does not apply (the whole screen is needed).  Two passes of N-point line
 * transforms: the register radix FFT for N = 64..2048 powers of two, the chirp-z convolution for any
 * other even N <= 1024 (e.g. the reference's auto-sized 164); a direct DFT serves the remaining sizes.
 *
 * K4b fastb_temporal_detect: the per-step body (fast/fast.py:619-633) fused with
 * Fast.compute_detector (fast/fast.py:647-668).  For step j and pupil pixel (a, b)
 *   phi = sum_l bilinear(screen_l; row = xi[l,j,a] + xf[l,j,a], col = yi[l,j,b] + yf[l,j,b])
 *   z_j = exp(chi_j) sum(U exp(i phi)) / sum(U);  out_j = |z_j|^2 or z_j (coherent)
 * The (integer, fraction) sample coordinates, layout [L][n_steps][n_pup], follow the reference
 * exactly (wrap, sort, roll, FITPACK clamp at N-1), with 0 <= xi, yi <= N-2 and 0 <= xf, yf <= 1;
 * the host may prepare them (fast_b200/temporal.py) or
 *
 * K4c fastb_temporal_coords: the coordinate bookkeeping of fast/fast.py:617-635 for EVERY step of a run
 * (n_chunks chunks of steps_per_chunk steps) on the device, operation for operation in float64 so that
 * the (integer, fraction) pairs equal numpy's: for layer l, axis (0: rows -> xi / xf, 1: columns ->
 * yi / yf), chunk c, step j, pupil pixel p
 *   x = (lo + p) + pixel_shifts[l][axis][j], then + pixel_shifts[l][axis][steps_per_chunk - 1] c times
 *       (the reference's `interp_coords +=` after every chunk, fast/fast.py:635);
 *   x mod N (numpy's float modulo), the n_pup values sorted, rolled back by the first argmax of the
 *   gaps (0 when all gaps are 1 within numpy.isclose), clamped at N - 1, split into floor and fraction.
 *   d_pixel_shifts  L*2*steps_per_chunk float64 (fast/fast.py:543-544)
 *   outputs         [L][n_chunks * steps_per_chunk][n_pup]
 * ------------------------------------------------------------------------------------- */
#define FASTB_LAYER_PAIR_BASE (1ULL << 62)

int64_t fastb_layer_screens_workspace_bytes(int32_t n, int32_t n_layers);
int fastb_layer_screens(int32_t n, int32_t n_layers, uint64_t seed, const float* d_weight_per_layer,
                        const float* d_noise, float* d_screens, void* d_workspace,
                        int64_t workspace_bytes, void* stream);

typedef struct FastbTemporalParams {
    int32_t n;                      /* N */
    int32_t n_pup;                  /* Npxls_pup */
    int32_t n_layers;               /* L */
    int32_t coherent;               /* 0: |z|^2; 1: z (re, im) */
    int64_t n_steps;                /* time steps in this call (one chunk) */
    double u_sum;                   /* sum(U) */
} FastbTemporalParams;

int fastb_temporal_coords(int32_t n, int32_t n_pup, int32_t lo, int32_t n_layers, int32_t steps_per_chunk,
                          int32_t n_chunks, const double* d_pixel_shifts, int32_t* d_xi, float* d_xf,
                          int32_t* d_yi, float* d_yf, void* stream);

int fastb_temporal_detect(const FastbTemporalParams* p, const float* d_screens, const int32_t* d_xi,
                          const float* d_xf, const int32_t* d_yi, const float* d_yf, const float* d_U,
                          const float* d_chi, float* d_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * K3: result statistics for the multi-GPU reduction (no reference equivalent; feeds the
 * FastResult summaries of fast/fast.py:949-994 without a full gather).
 * d_sums[8]   += { n, sum r, sum r^2, sum dB, sum dB^2, 0, 0, 0 }   (dB = 10 log10 r)
 * d_minmax[2]  = { min(old, min r), max(old, max r) }
 * d_hist[nbins+2] += histogram of dB on [db_lo, db_hi) with under/overflow in the last two.
 * Buffers are accumulated into (caller zeroes / initialises them), so that a sum all-reduce
 * of d_sums / d_hist and a min/max all-reduce of d_minmax combine ranks.
 * ------------------------------------------------------------------------------------- */
int fastb_stats(const float* d_r, int64_t n, double db_lo, double db_hi, int32_t nbins,
                double* d_sums, double* d_minmax, unsigned long long* d_hist, void* stream);

/* ---------------------------------------------------------------------------------------
 * K5: link metrics on the per-realisation results (consumers of FastResult.power,
 * fast/comms.py).  All arrays are device pointers; outputs are overwritten (not accumulated)
 * unless stated.  Arithmetic is float64 on the float32 samples, as the reference computes on
 * float64 copies of the same values.
 * ------------------------------------------------------------------------------------- */
enum { FASTB_CURVE_BER_OOK = 0, FASTB_CURVE_SEP_QAM = 1 };

/* Mean over the samples of a closed-form error probability, one value per SNR point; replaces
 * ber_ook / sep_qam (fast/comms.py:193-240; ber_qam :243-253 is sep_qam rescaled by the caller):
 *   s_i = x_i / mean(x)
 *   BER_OOK : Q(s_i sqrt(10^(snr_db/10)))
 *   SEP_QAM : 4 (a q - a^2 q^2), q = Q(sqrt(3/(M-1) 10^(snr_db/10) s_i^2)), a = (sqrt M - 1)/sqrt M
 * d_curve has k + 1 doubles: [0, k) the curve, [k] the sample mean that was used. */
int fastb_error_curve(const float* d_samples, int64_t n, int32_t kind, int32_t qam_order,
                      const double* d_snr_db, int32_t k, double* d_curve, void* stream);

/* Fade statistics of a time series for k thresholds; replaces fade_prob / fade_dur
 * (fast/comms.py:171-191).  d_out[4 j + {0,1,2,3}] = samples below threshold j, complete fades,
 * samples inside complete fades, index of the last sample not below the threshold (-1 if none).
 * A complete fade is a maximal below-threshold run that starts at index >= 1 and ends before the
 * last sample (the reference drops a fade in progress at either end of the window).
 * probability = out[0]/n; mean duration = dt out[2]/out[1]. */
int fastb_fade_stats(const float* d_series, int64_t n, const double* d_thresholds, int32_t k,
                     int64_t* d_out, void* stream);

/* Monte-Carlo modulator; replaces Modulator.run (fast/comms.py:13-146): for realisation i and
 * symbol slot s < symbols_per_iter draw a symbol, add AWGN of standard deviation
 *   OOK: Es / snr_i (real)      otherwise: sqrt(Es/2) / snr_i per component,
 *   snr_i = sqrt(10^(EsN0/10)) x_i / mean(x)          (= snr_scale * x_i)
 * and decide: OOK re > 0.5; BPSK re < 0 -> symbol 1; otherwise the nearest constellation point
 * (lowest index wins ties).  Es = mean |c|^2 over the constellation.
 * RNG (statistical parity; the reference uses numpy's unseeded global generator): Philox4x32-10,
 * key = seed, counter (i lo, i hi, s, 0x30D0A700) with i = first + local index -> words w0..w3:
 * symbol = (w0 * n_symbols) >> 32; noise = Box-Muller of (w1 >> 9, w2 >> 9) as in K2.
 * d_sums[3] += { symbol errors, sum |rx - tx|, sum |tx|^2 } over n * symbols_per_iter draws
 * (accumulated: the caller zeroes it; sum-reducible over ranks):
 *   SEP = sums[0] / (n S);   EVM = (sums[1] / (n S)) / sqrt(sums[2] / (n S)).
 * Optional outputs, [s * n + i] like the reference's (S, n) arrays, any may be NULL:
 * d_symbols (uint8), d_recv (float2), d_recv_symbols (uint8).
 * d_tx_symbols (optional, symbols_per_iter uint8): transmit this sequence in every realisation
 * instead of random symbols (the reference's `data` mode, fast/comms.py:53-56). */
enum { FASTB_MOD_OOK = 0, FASTB_MOD_BPSK = 1, FASTB_MOD_NEAREST = 2 };
typedef struct FastbModParams {
    int64_t n;                 /* realisations on this rank */
    int64_t first;             /* global index of the first one (RNG counter) */
    int32_t symbols_per_iter;
    int32_t n_symbols;         /* <= 1024 */
    int32_t scheme;            /* FASTB_MOD_* */
    int32_t has_awgn;          /* 0: EsN0 = None */
    double es;                 /* mean |c|^2 */
    double snr_scale;          /* sqrt(10^(EsN0/10)) / mean(x) */
    uint64_t seed;
} FastbModParams;
int fastb_modulator_mc(const FastbModParams* p, const float* d_power, const float* d_constellation /* 2 n_symbols: re, im */,
                       double* d_sums, uint8_t* d_symbols, float* d_recv /* 2 n S */, uint8_t* d_recv_symbols,
                       const uint8_t* d_tx_symbols, void* stream);

/* |z_i| of the samples in float64 (is_complex: d_samples holds n (re, im) pairs, else n reals)
 * and d_sums[2] = { sum |z|, sum |z|^2 } (overwritten); feeds the geometry of the I-Q histograms
 * (numpy.abs(samples), numpy.mean(numpy.abs(samples)), fast/comms.py:355-356,390). */
int fastb_amplitudes(const float* d_samples, int32_t is_complex, int64_t n, double* d_amp,
                     double* d_sums, void* stream);

/* I-Q plane histograms; replaces the binning of convolve_awgn_qam (fast/comms.py:380-393):
 * for every constellation point c the samples z = c * amp_i are binned on the edges
 * d_edges_x[c][0..npxls], d_edges_y[c][0..npxls] with numpy.histogram2d's rule (right-open bins,
 * last edge inclusive, outside values dropped).  d_counts[c][ix][iy] (uint32) is overwritten. */
int fastb_iq_histogram(const double* d_amp, int64_t n, const double* d_points /* 2 M: re, im */, int32_t m,
                       const double* d_edges_x, const double* d_edges_y, int32_t npxls,
                       uint32_t* d_counts, void* stream);

/* AWGN convolution of the histograms; replaces fast/comms.py:395-411.
 *   shot == 0: out[c] = G h G^T, h = counts / n, G[i][j] = taps[j - i + (npxls+1)/2] (zero outside:
 *              scipy.ndimage.correlate1d, mode 'constant'), taps has npxls + 1 entries.
 *   shot != 0: every occupied bin (i, j) spreads as exp(-((i-y)^2 + (j-x)^2) / (sigma2 mult)) h /
 *              (pi sigma2 mult), mult = mean_amp^2 / (edges_x[c][i]^2 + edges_y[c][j]^2).
 * d_out: m * npxls * npxls doubles; d_workspace: fastb_iq_convolve_workspace_bytes(m, npxls). */
int64_t fastb_iq_convolve_workspace_bytes(int32_t m, int32_t npxls);
int fastb_iq_convolve(const uint32_t* d_counts, int64_t n, int32_t m, int32_t npxls, const double* d_taps,
                      int32_t shot, double sigma2, double mean_amp, const double* d_edges_x,
                      const double* d_edges_y, double* d_out, void* d_workspace, int64_t workspace_bytes,
                      void* stream);

/* Mutual information measures of the convolved histograms f[c][pixel]; replaces the reductions
 * of mutual_information_qam / generalised_mutual_information_qam (fast/comms.py:262-303), with
 * x log2 x := 0 for x <= 0 (the reference's masked-array semantics):
 *   d_out[0] = (1/M) sum_c sum_pix f_c (log2 f_c - log2 fy),           fy = mean_c f_c
 *   d_out[1] = sum_{bit b < n_bits} 1/2 sum_pix sum_{v in {0,1}} fb_v (log2 fb_v - log2 fy),
 *              fb_v = mean of f_c over the symbols whose bit b (from the MSB) of d_gray[c] is v. */
int fastb_iq_information(const double* d_f, int32_t m, int32_t npxls, const uint32_t* d_gray, int32_t n_bits,
                         double* d_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * library
 * ------------------------------------------------------------------------------------- */
int fastb_version(void);
const char* fastb_last_error(void);
int fastb_device_count(void);       /* number of usable CUDA devices, 0 if none */
/* Kernels launched by this library on the calling thread since the last reset. */
int64_t fastb_launch_count(void);
void fastb_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FASTB_H */
