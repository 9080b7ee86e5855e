// K1: residual phase PSD build, float64, one thread per frequency pixel.
// Contract and reference citations: include/fastb.h.  Compiled with -fmad=false so that the
// arithmetic follows the reference's (numpy) operation order without FMA contraction.
#include "fastb_common.cuh"

#include <math.h>

namespace fastb {

namespace {

constexpr double kPi = 3.141592653589793;
constexpr int kLayerTile = 4;

struct PsdDev {
    FastbPsdParams p;
    FastbPsdInputs in;
    FastbPsdOutputs out;
};

// numpy.sinc: sin(pi x) / (pi x) with x == 0 replaced by 1e-20
__device__ __forceinline__ double np_sinc(double x) {
    const double y = kPi * (x == 0.0 ? 1.0e-20 : x);
    return sin(y) / y;
}

// 0.033 exp(-f^2/km^2) / (f^2 + k0^2)^(11/6), +-inf -> 0   (fast/funcs.py:151-170)
__device__ __forceinline__ double von_karman_base(double fabs, double km2, double k02) {
    const double f2 = fabs * fabs;
    double v = 0.033 * exp(-f2 / km2) / pow(f2 + k02, 11.0 / 6.0);
    if (isinf(v)) v = 0.0;
    return v;
}

__global__ void __launch_bounds__(128) psd_kernel(const __grid_constant__ PsdDev devv) {
    const PsdDev* dev = &devv;
    const FastbPsdParams& p = dev->p;
    const int n = p.n;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= n) return;
    const size_t pix = (size_t)r * n + c;
    const size_t nn = (size_t)n * n;
    const int L = p.n_layers;
    const int mid = n / 2;

    const double fx = (double)(c - mid) * p.df;
    const double fy = (double)(r - mid) * p.df;
    const double fabs_ = sqrt(fx * fx + fy * fy);
    const double f2 = fabs_ * fabs_;
    const double km = 5.92 / p.l0;
    const double km2 = km * km;
    const double k0 = 2.0 * kPi / p.L0;
    const double k02 = k0 * k0;
    const double two_pi = 2.0 * kPi;
    const double kk = p.k * p.k;
    const bool ao = p.ao_mode != FASTB_AO_NOAO;

    double M;
    if (dev->in.d_lf_mask) {
        M = dev->in.d_lf_mask[pix];
    } else {
        const double fmax = kPi / p.dsubap;
        M = (fabs(fx) <= fmax && fabs(fy) <= fmax) ? 1.0 : 0.0;
    }

    const double base = von_karman_base(fabs_, km2, k02);

    // WFS noise PSD (fast/ao_power_spectra.py:148-161)
    double noise = 0.0;
    const bool has_noise = ao && p.noise_var > 0.0;
    if (has_noise) {
        const double sx = np_sinc(p.dsubap * fx / two_pi);
        const double sy = np_sinc(p.dsubap * fy / two_pi);
        double ps = p.noise_var / (f2 * (sx * sx) * (sy * sy));
        if (r == mid && c == mid) ps = 0.0;
        noise = M * ps;
    }

    const bool has_alias = ao && p.alias != 0;
    const double t0 = has_alias ? (fx * fx) * (fy * fy) / pow(fabs_, 4.0) : 0.0;
    const double zf = (p.ao_mode == FASTB_AO_LGSAO) ? dev->in.d_zfilter[pix] : 0.0;

    double W = 0.0, as_sum = 0.0, alias_sum = 0.0, logamp = 0.0;
    const double noise_per_layer = noise / (double)L;
    const double la_c = 2.0 * kPi * ((two_pi / p.wvl) * (two_pi / p.wvl));
    const double pf = dev->in.d_pupil_filter ? dev->in.d_pupil_filter[pix] : 1.0;

    for (int l0 = 0; l0 < L; l0 += kLayerTile) {
        double acc[kLayerTile];
#pragma unroll
        for (int j = 0; j < kLayerTile; ++j) acc[j] = 0.0;

        // aliasing replicas (fast/ao_power_spectra.py:195-214); skipped where the mask is
        // exactly zero: the product with the mask is then 0 (or NaN -> 0) in the reference too
        if (has_alias && M != 0.0) {
            for (int l = -p.lmax; l <= p.lmax; ++l) {
                const double sy = fy - two_pi * (double)l / p.dsubap;
                for (int k = -p.kmax; k <= p.kmax; ++k) {
                    if (l == 0 && k == 0) continue;
                    const double sx = fx - two_pi * (double)k / p.dsubap;
                    const double fs = sqrt(sx * sx + sy * sy);
                    const double b2 = von_karman_base(fs, km2, k02);
                    const double q = fx / sy + fy / sx;
                    const double t1 = q * q;
                    const bool take_t2 = (l == 0 && r == mid) || (k == 0 && c == mid);
                    const bool dc = (r == mid && c == mid);
#pragma unroll
                    for (int j = 0; j < kLayerTile; ++j) {
                        if (l0 + j < L) {
                            const double t2 = b2 * p.cn2[l0 + j];
                            double m = (t1 * t2) * t0;
                            if (dc) m = 0.0;
                            if (take_t2) m = t2;
                            acc[j] += m;
                        }
                    }
                }
            }
        }

#pragma unroll
        for (int j = 0; j < kLayerTile; ++j) {
            const int li = l0 + j;
            if (li >= L) break;
            const double turb = base * p.cn2[li];
            double G = 1.0, alias = 0.0;
            if (ao) {
                const double drx = p.dtheta[0] / 206265.0 * p.h[li];
                const double dry = p.dtheta[1] / 206265.0 * p.h[li];
                const double a = fx * drx + fy * dry;
                const double b = fx * p.vx[li] + fy * p.vy[li];
                const double s = np_sinc(p.texp * b / two_pi);
                const double aniso = 1.0 - (2.0 * cos(a - p.tloop * b)) * s + s * s;
                if (p.ao_mode == FASTB_AO_LGSAO) {
                    const double aniso_lgs = 1.0 - (2.0 * cos(-p.tloop * b)) * s + s * s;
                    G = M * (zf * aniso + (1.0 - zf) * aniso_lgs) + (1.0 - M);
                } else {
                    G = aniso * M + (1.0 - M);
                }
                if (has_alias) {
                    alias = acc[j] * ((s * s) * M);
                    if (isnan(alias)) alias = 0.0;
                }
            }
            const double per_layer = ((2.0 * kPi) * kk) * (turb * G + alias) + noise_per_layer;
            W += per_layer;
            as_sum += G * turb;
            alias_sum += ((alias * 2.0) * kPi) * kk;
            if (dev->out.d_logamp) {
                const double sn = sin(((p.wvl * p.h[li]) * f2) / (4.0 * kPi));
                logamp += ((turb * la_c) * (sn * sn)) * pf;
            }
            const size_t o = (size_t)li * nn + pix;
            if (dev->out.d_powerspec_per_layer) dev->out.d_powerspec_per_layer[o] = per_layer;
            if (dev->out.d_turb) dev->out.d_turb[o] = turb;
            if (dev->out.d_g_ao) dev->out.d_g_ao[o] = G;
            if (dev->out.d_alias) dev->out.d_alias[o] = alias;
            if (dev->out.d_weight_per_layer) {
                const double wgt = sqrt(per_layer) * p.df;
                dev->out.d_weight_per_layer[o] = (float)(((r + c) & 1) ? -wgt : wgt);
            }
        }
    }

    dev->out.d_powerspec[pix] = W;
    if (dev->out.d_noise) dev->out.d_noise[pix] = noise;
    if (dev->out.d_logamp) dev->out.d_logamp[pix] = logamp;
    if (dev->out.d_integrands) {
        dev->out.d_integrands[pix] = (((as_sum * M) * 2.0) * kPi) * kk;
        dev->out.d_integrands[nn + pix] = alias_sum;
        dev->out.d_integrands[2 * nn + pix] = W * (1.0 - M);
    }
    if (dev->out.d_weight) {
        const double wgt = sqrt(W) * p.df;
        dev->out.d_weight[pix] = (float)(((r + c) & 1) ? -wgt : wgt);
    }
}

__global__ void make_weight_kernel(const double* __restrict__ W, int n, size_t total, double df,
                                   float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t pix = i % ((size_t)n * n);
    const int r = (int)(pix / n), c = (int)(pix % n);
    const double w = sqrt(W[i]) * df;
    out[i] = (float)(((r + c) & 1) ? -w : w);
}

// out[b] = sum_r w_r sum_c P[b, r, c] w_c ; one 1024-thread block per b, fixed-order tree
__global__ void __launch_bounds__(1024) simpson2d_kernel(const double* __restrict__ P, int n,
                                                         const double* __restrict__ w,
                                                         double* __restrict__ out) {
    __shared__ double red[1024];
    const double* Pb = P + (size_t)blockIdx.x * n * n;
    double acc = 0.0;
    for (int r = threadIdx.x / 32; r < n; r += 32) {          // one warp per row
        double row = 0.0;
        for (int c = threadIdx.x % 32; c < n; c += 32) row += Pb[(size_t)r * n + c] * w[c];
        acc += row * w[r];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// ---- pupil filter: centred DFT2 by two direct passes (float64, any N) -------------------
__global__ void twiddle_kernel(int n, double2* __restrict__ tw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)n, &s, &c);
    tw[j] = make_double2(c, s);
}

__device__ __forceinline__ int centred_index(int a, int b, int n) {
    const int mid = n / 2;
    long long x = (long long)(a - mid) * (long long)(b - mid);
    int m = (int)(x % n);
    return m < 0 ? m + n : m;
}

// T[r, kc] = sum_c pm[r, c] tw[(c-mid)(kc-mid) mod n]
__global__ void dft_rows_kernel(const double* __restrict__ pm, int n, const double2* __restrict__ tw,
                                double2* __restrict__ T) {
    const int kc = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (kc >= n) return;
    double re = 0.0, im = 0.0;
    for (int c = 0; c < n; ++c) {
        const double v = pm[(size_t)r * n + c];
        if (v != 0.0) {
            const double2 t = tw[centred_index(c, kc, n)];
            re += v * t.x;
            im += v * t.y;
        }
    }
    T[(size_t)r * n + kc] = make_double2(re, im);
}

// F[kr, kc] = sum_r T[r, kc] tw[(r-mid)(kr-mid) mod n] ; writes |F|^2
__global__ void dft_cols_kernel(const double2* __restrict__ T, int n, const double2* __restrict__ tw,
                                double* __restrict__ pf) {
    const int kc = blockIdx.x * blockDim.x + threadIdx.x;
    const int kr = blockIdx.y;
    if (kc >= n) return;
    double re = 0.0, im = 0.0;
    for (int r = 0; r < n; ++r) {
        const double2 v = T[(size_t)r * n + kc];
        const double2 t = tw[centred_index(r, kr, n)];
        re += v.x * t.x - v.y * t.y;
        im += v.x * t.y + v.y * t.x;
    }
    pf[(size_t)kr * n + kc] = re * re + im * im;
}

__global__ void pf_normalise_kernel(double* __restrict__ pf, int n) {
    // |F(0)|^2 = (sum pm)^2 sits at the centre pixel
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nn = (size_t)n * n;
    __shared__ double dc;
    if (threadIdx.x == 0) dc = pf[(size_t)(n / 2) * n + n / 2];
    __syncthreads();
    if (i < nn && i != (size_t)(n / 2) * n + n / 2) pf[i] = pf[i] / dc;
}

__global__ void pf_centre_kernel(double* __restrict__ pf, int n) {
    pf[(size_t)(n / 2) * n + n / 2] = 1.0;
}


// ---- modal masks: Fourier-space Zernike filters (fast/ao_power_spectra.py:10-141) -----------
// |Z_j(f)|^2 of Noll (1976): radial order n, azimuthal order m of Noll index j,
//   m == 0 : (n + 1) (2 J_{n+1}(x) / x)^2                     x = |f| D / 2
//   m != 0 : 2 (n + 1) (2 J_{n+1}(x) / x)^2 {cos, sin}^2(m phi)   (cos for even j), phi = atan2(fy, fx)
// summed over j = noll_first .. noll_last; DC pixel := 1 if the sum starts at j = 1, else 0.
__global__ void __launch_bounds__(128) zernike_filter_kernel(const __grid_constant__ FastbZernikeParams p,
                                                             double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= p.n) return;
    const int half = p.n / 2;
    const double fx = (double)(c - half) * p.df, fy = (double)(r - half) * p.df;
    const double fabs_ = sqrt(fx * fx + fy * fy);
    double val;
    if (p.noll_last < p.noll_first) {
        // modal DM without a Zernike limit: a disc of radius modal_mult * pi / d
        val = (fabs_ <= M_PI / p.d_wfs * p.modal_mult) ? 1.0 : 0.0;
    } else {
        const double x = fabs_ * p.diameter / 2.0;
        const double phi = atan2(fy, fx);
        val = 0.0;
        int n_prev = -1;
        double radial2 = 0.0;
        for (int j = p.noll_first; j <= p.noll_last; ++j) {
            // aotools zernIndex: Noll j -> (n, |m|)
            const int n = (int)((-1.0 + sqrt((double)(8 * (j - 1) + 1))) / 2.0);
            const int pj = j - n * (n + 1) / 2;
            const int k = n % 2;
            const int m = ((pj + k) / 2) * 2 - k;
            if (n != n_prev) {
                const double radial = 2.0 * jn(n + 1, x) / x;
                radial2 = radial * radial;
                n_prev = n;
            }
            if (m == 0) {
                val += (double)(n + 1) * radial2;
            } else {
                const double az = (j % 2 == 0) ? cos((double)m * phi) : sin((double)m * phi);
                val += 2.0 * (double)(n + 1) * radial2 * (az * az);
            }
        }
        if (r == half && c == half) val = (p.noll_first == 1) ? 1.0 : 0.0;
        if (p.gtilt) {
            const double j1 = jn(1, x);
            val += j1 * j1;
        }
    }
    if (p.clip_box) {
        // mask_lf: WFS box |fx|, |fy| <= pi/d times the DM term clipped to <= 1 (NaN -> 1)
        const double fmax = M_PI / p.d_wfs;
        const double dm = (val < 1.0) ? val : 1.0;
        val = (fabs(fx) <= fmax && fabs(fy) <= fmax) ? dm : 0.0;
    }
    out[(size_t)r * p.n + c] = val;
}

}  // namespace
}  // namespace fastb

using namespace fastb;

extern "C" int fastb_zernike_filter(const FastbZernikeParams* p, double* d_out, void* stream) {
    FASTB_REQUIRE(p && d_out, "fastb_zernike_filter: NULL pointer");
    FASTB_REQUIRE(p->n >= 2, "fastb_zernike_filter: n=%d must be >= 2", p->n);
    FASTB_REQUIRE(p->noll_first >= 1 && p->noll_last <= 100000, "fastb_zernike_filter: Noll range out of bounds");
    FASTB_REQUIRE(p->noll_last < p->noll_first || p->diameter > 0.0, "fastb_zernike_filter: diameter must be > 0");
    FASTB_REQUIRE(!(p->clip_box || p->noll_last < p->noll_first) || p->d_wfs > 0.0,
                  "fastb_zernike_filter: d_wfs must be > 0 for the box / disc");
    dim3 block(128), grid((p->n + 127) / 128, p->n);
    zernike_filter_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*p, d_out);
    return check_launch("zernike_filter_kernel");
}

extern "C" int fastb_psd_build(const FastbPsdParams* p, const FastbPsdInputs* in,
                               const FastbPsdOutputs* out, void* stream) {
    FASTB_REQUIRE(p && out, "fastb_psd_build: NULL params/outputs");
    FASTB_REQUIRE(p->n >= 2, "fastb_psd_build: n=%d must be >= 2", p->n);
    FASTB_REQUIRE(p->n_layers >= 1 && p->n_layers <= FASTB_MAX_LAYERS,
                  "fastb_psd_build: n_layers=%d outside 1..%d", p->n_layers, FASTB_MAX_LAYERS);
    FASTB_REQUIRE(p->ao_mode >= FASTB_AO_NOAO && p->ao_mode <= FASTB_AO_LGSAO,
                  "fastb_psd_build: unknown ao_mode %d", p->ao_mode);
    FASTB_REQUIRE(out->d_powerspec, "fastb_psd_build: d_powerspec is required");
    FASTB_REQUIRE(p->lmax >= 0 && p->kmax >= 0 && p->lmax <= 64 && p->kmax <= 64,
                  "fastb_psd_build: lmax/kmax out of range");
    FastbPsdInputs in0 = {nullptr, nullptr, nullptr};
    if (in) in0 = *in;
    FASTB_REQUIRE(p->ao_mode != FASTB_AO_LGSAO || in0.d_zfilter,
                  "fastb_psd_build: LGSAO needs d_zfilter");
    FASTB_REQUIRE(!out->d_logamp || in0.d_pupil_filter,
                  "fastb_psd_build: d_logamp needs d_pupil_filter");

    cudaStream_t st = (cudaStream_t)stream;
    PsdDev host;
    host.p = *p;
    host.in = in0;
    host.out = *out;
    dim3 block(128), grid((p->n + 127) / 128, p->n);
    psd_kernel<<<grid, block, 0, st>>>(host);
    return check_launch("psd_kernel");
}

extern "C" int fastb_make_weight(const double* d_W, int32_t n, int32_t batch, double df,
                                 float* d_weight, void* stream) {
    FASTB_REQUIRE(d_W && d_weight, "fastb_make_weight: NULL pointer");
    FASTB_REQUIRE(n >= 2 && batch >= 1, "fastb_make_weight: bad n/batch");
    const size_t total = (size_t)n * n * batch;
    make_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_W, n, total, df, d_weight);
    return check_launch("make_weight_kernel");
}

extern "C" int fastb_simpson2d(const double* d_P, int32_t n, int32_t batch, const double* d_w,
                               double* d_out, void* stream) {
    FASTB_REQUIRE(d_P && d_w && d_out, "fastb_simpson2d: NULL pointer");
    FASTB_REQUIRE(n >= 1 && batch >= 1, "fastb_simpson2d: bad n/batch");
    simpson2d_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(d_P, n, d_w, d_out);
    return check_launch("simpson2d_kernel");
}

extern "C" int64_t fastb_pupil_filter_workspace_bytes(int32_t n) {
    if (n < 2) return 0;
    return (int64_t)n * n * (int64_t)sizeof(double2) + (int64_t)n * (int64_t)sizeof(double2);
}

extern "C" int fastb_pupil_filter(const double* d_pm, int32_t n, double* d_pf, void* d_workspace,
                                  int64_t workspace_bytes, void* stream) {
    FASTB_REQUIRE(d_pm && d_pf && d_workspace, "fastb_pupil_filter: NULL pointer");
    FASTB_REQUIRE(n >= 2 && (n % 2) == 0, "fastb_pupil_filter: n=%d must be even", n);
    FASTB_REQUIRE(workspace_bytes >= fastb_pupil_filter_workspace_bytes(n),
                  "fastb_pupil_filter: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double2* T = (double2*)d_workspace;
    double2* tw = T + (size_t)n * n;
    twiddle_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, tw);
    int rc = check_launch("twiddle_kernel");
    if (rc) return rc;
    dim3 block(128), grid((n + 127) / 128, n);
    dft_rows_kernel<<<grid, block, 0, st>>>(d_pm, n, tw, T);
    if ((rc = check_launch("dft_rows_kernel"))) return rc;
    dft_cols_kernel<<<grid, block, 0, st>>>(T, n, tw, d_pf);
    if ((rc = check_launch("dft_cols_kernel"))) return rc;
    const size_t nn = (size_t)n * n;
    pf_normalise_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(d_pf, n);
    if ((rc = check_launch("pf_normalise_kernel"))) return rc;
    pf_centre_kernel<<<1, 1, 0, st>>>(d_pf, n);
    return check_launch("pf_centre_kernel");
}
