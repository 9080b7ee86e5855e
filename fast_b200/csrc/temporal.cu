// K4: TEMPORAL (frozen-flow) mode.  Contract and reference citations: include/fastb.h.
//   K4a  layer screens, once per simulation: two passes of line FFTs (layer_screens_fft.cu: radix for
//        powers of two, chirp-z for other even N <= 1024); the direct 2-D inverse DFT below serves the
//        remaining sizes (any even N <= 4096).
//   K4b  per time step: bilinear gather of the L layer screens at the wind-shifted pupil
//        coordinates, layer sum, detector -- one scalar per step reaches HBM.
#include "fastb_common.cuh"

namespace fastb {

bool layer_fft_ok(int n);
size_t layer_fft_workspace_bytes(int n, int L);
int layer_screens_fft(int n, int L, unsigned long long seed, const float* weight, const float* noise, float* screens,
                      void* workspace, cudaStream_t st);

namespace {

constexpr int kThreads = 256;

__global__ void twiddle_f32_kernel(int n, float2* __restrict__ tw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s, c;
    sincospi(2.0 * (double)j / (double)n, &s, &c);
    tw[j] = make_float2((float)c, (float)s);
}

// T[l][r'][c] = sum_c' S[l][r'][c'] w^(c' c),  S = noise * signed weight.  grid (N, L)
template <bool RNG>
__global__ void __launch_bounds__(kThreads) screens_rows_kernel(int n, unsigned long long seed,
                                                                const float* __restrict__ weight,
                                                                const float2* __restrict__ noise,
                                                                const float2* __restrict__ tw_g,
                                                                float2* __restrict__ T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float2* row = tw + n;
    const int r = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
    const size_t base = ((size_t)l * n + r) * n;
    for (int j = tid; j < n; j += kThreads) tw[j] = tw_g[j];
    if (RNG) {
        const unsigned long long g = FASTB_LAYER_PAIR_BASE + (unsigned long long)l;
        const int S = (n + 15) / 16;
        for (int t = tid; t < S; t += kThreads) {
            uint32_t mr[16], ma[16];
            noise_block_fields((uint32_t)(r * S + t), g, (uint32_t)seed, (uint32_t)(seed >> 32), mr, ma);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int j = t + S * m;
                if (j < n) row[j] = weighted_normal_m(mr[m], ma[m], weight[base + j]);
            }
        }
    } else {
        for (int j = tid; j < n; j += kThreads) {
            const float2 z = noise[base + j];
            const float w = weight[base + j];
            row[j] = make_float2(z.x * w, z.y * w);
        }
    }
    __syncthreads();
    for (int c = tid; c < n; c += kThreads) {
        int ti = 0;
        float sr = 0.f, si = 0.f;
        for (int cp = 0; cp < n; ++cp) {
            const float2 x = row[cp], t = tw[ti];
            sr = fmaf(x.x, t.x, fmaf(-x.y, t.y, sr));
            si = fmaf(x.x, t.y, fmaf(x.y, t.x, si));
            ti += c;
            if (ti >= n) ti -= n;
        }
        T[base + c] = make_float2(sr, si);
    }
}

// screen[l][r][c] = (-1)^(r+c) Re sum_r' T[l][r'][c] w^(r' r).  grid (N, L)
__global__ void __launch_bounds__(kThreads) screens_cols_kernel(int n, const float2* __restrict__ tw_g,
                                                                const float2* __restrict__ T,
                                                                float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);      // tw[(r' r) mod n] for this block's r
    const int r = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
    for (int rp = tid; rp < n; rp += kThreads) tw[rp] = tw_g[(int)(((long long)rp * r) % n)];
    __syncthreads();
    const float2* Tl = T + (size_t)l * n * n;
    for (int c = tid; c < n; c += kThreads) {
        float acc = 0.f;
        for (int rp = 0; rp < n; ++rp) {
            const float2 x = Tl[(size_t)rp * n + c], t = tw[rp];
            acc = fmaf(x.x, t.x, fmaf(-x.y, t.y, acc));
        }
        out[((size_t)l * n + r) * n + c] = ((r + c) & 1) ? -acc : acc;
    }
}

struct TemporalArgs {
    int n, n_pup, n_layers, coherent;
    long long n_steps;
    float inv_usum;
    const float* screens;
    const int* xi;
    const float* xf;
    const int* yi;
    const float* yf;
    const float* U;
    const float* chi;
    float* out;
};

__global__ void __launch_bounds__(kThreads) temporal_detect_kernel(const __grid_constant__ TemporalArgs a) {
    __shared__ float red[2 * (kThreads / 32)];
    const int P = a.n_pup, N = a.n, L = a.n_layers, tid = threadIdx.x;
    for (long long j = blockIdx.x; j < a.n_steps; j += gridDim.x) {
        float ac = 0.f, as = 0.f;
        for (int idx = tid; idx < P * P; idx += kThreads) {
            const int pa = idx / P, pb = idx % P;
            float phi = 0.f;
            for (int l = 0; l < L; ++l) {
                const size_t o = ((size_t)l * a.n_steps + j) * P;
                const int ix = a.xi[o + pa], iy = a.yi[o + pb];
                const float fx = a.xf[o + pa], fy = a.yf[o + pb];
                const float* s = a.screens + ((size_t)l * N + ix) * N + iy;
                const float top = s[0] * (1.f - fy) + s[1] * fy;
                const float bot = s[N] * (1.f - fy) + s[N + 1] * fy;
                phi += top * (1.f - fx) + bot * fx;
            }
            float sn, cs;
            const float k = rintf(phi * 0.15915494309189535f);       // explicit reduction: phases of
            float rr = fmaf(-k, 6.2831854820251465f, phi);           // un-corrected layers can be large
            rr = fmaf(-k, -1.7484556000744883e-07f, rr);
            __sincosf(rr, &sn, &cs);
            const float u = a.U[idx];
            ac = fmaf(u, cs, ac);
            as = fmaf(u, sn, as);
        }
        for (int o = 16; o > 0; o >>= 1) {
            ac += __shfl_xor_sync(0xffffffffu, ac, o);
            as += __shfl_xor_sync(0xffffffffu, as, o);
        }
        if ((tid & 31) == 0) {
            red[2 * (tid >> 5)] = ac;
            red[2 * (tid >> 5) + 1] = as;
        }
        __syncthreads();
        if (tid == 0) {
            float tc = 0.f, tsn = 0.f;
            for (int w = 0; w < kThreads / 32; ++w) {
                tc += red[2 * w];
                tsn += red[2 * w + 1];
            }
            const float e = expf(a.chi[j]) * a.inv_usum;
            const float zr = e * tc, zi = e * tsn;
            if (a.coherent) {
                a.out[2 * j] = zr;
                a.out[2 * j + 1] = zi;
            } else {
                a.out[j] = zr * zr + zi * zi;
            }
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace fastb

using namespace fastb;

extern "C" int64_t fastb_layer_screens_workspace_bytes(int32_t n, int32_t n_layers) {
    if (n < 2 || n_layers < 1) return 0;
    if (layer_fft_ok(n)) return (int64_t)layer_fft_workspace_bytes(n, n_layers);
    return (int64_t)sizeof(float2) * ((int64_t)n + (int64_t)n_layers * n * n);
}

extern "C" int fastb_layer_screens(int32_t n, int32_t n_layers, uint64_t seed, const float* d_weight_per_layer,
                                   const float* d_noise, float* d_screens, void* d_workspace,
                                   int64_t workspace_bytes, void* stream) {
    FASTB_REQUIRE(n >= 4 && (n % 2) == 0 && n <= 4096, "fastb_layer_screens: n=%d must be even, 4..4096", n);
    FASTB_REQUIRE(n_layers >= 1 && n_layers <= FASTB_MAX_LAYERS, "fastb_layer_screens: bad n_layers %d", n_layers);
    FASTB_REQUIRE(d_weight_per_layer && d_screens && d_workspace, "fastb_layer_screens: NULL pointer");
    FASTB_REQUIRE(workspace_bytes >= fastb_layer_screens_workspace_bytes(n, n_layers),
                  "fastb_layer_screens: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (layer_fft_ok(n))
        return layer_screens_fft(n, n_layers, seed, d_weight_per_layer, d_noise, d_screens, d_workspace, st);
    float2* tw = (float2*)d_workspace;
    float2* T = tw + n;
    twiddle_f32_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, tw);
    int rc = check_launch("twiddle_f32_kernel");
    if (rc) return rc;
    dim3 grid(n, n_layers);
    const size_t smem_rows = sizeof(float2) * 2 * (size_t)n, smem_cols = sizeof(float2) * (size_t)n;
    if (d_noise) {
        FASTB_CUDA(cudaFuncSetAttribute(screens_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_rows));
        screens_rows_kernel<false><<<grid, kThreads, smem_rows, st>>>(n, seed, d_weight_per_layer,
                                                                     (const float2*)d_noise, tw, T);
    } else {
        FASTB_CUDA(cudaFuncSetAttribute(screens_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_rows));
        screens_rows_kernel<true><<<grid, kThreads, smem_rows, st>>>(n, seed, d_weight_per_layer, nullptr, tw, T);
    }
    if ((rc = check_launch("screens_rows_kernel"))) return rc;
    screens_cols_kernel<<<grid, kThreads, smem_cols, st>>>(n, tw, T, d_screens);
    return check_launch("screens_cols_kernel");
}

namespace fastb {
namespace {

// ---- K4c: sample coordinates of every time step (fast/fast.py:617-635), one block per (layer, axis, step) row ----
// The reference's bookkeeping, operation for operation in float64 so that the integer / fraction pairs are the ones
// numpy produces: x_p = (lo + p) + shift[l][axis][j], moved on by the chunk's total shift once per earlier chunk;
// wrap with numpy's float modulo; sort; roll back by the first argmax of the gaps (0 when every gap is 1 within
// numpy.isclose); FITPACK's clamp at the last knot; floor / fraction.
__global__ void __launch_bounds__(kThreads) temporal_coords_kernel(int n, int n_pup, int lo, int n_layers, int J,
                                                                   int n_chunks, const double* __restrict__ shifts,
                                                                   int* __restrict__ xi, float* __restrict__ xf,
                                                                   int* __restrict__ yi, float* __restrict__ yf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* c = reinterpret_cast<double*>(smem_raw);             // np2 sorted coordinates (+inf padded)
    __shared__ double red_v[kThreads];
    __shared__ int red_i[kThreads];
    __shared__ int all_close, roll_s;
    const int tid = threadIdx.x, P = n_pup;
    int np2 = 1;
    while (np2 < P) np2 <<= 1;
    const long long n_steps = (long long)J * n_chunks;
    const long long row = blockIdx.x;                            // ((l * 2 + axis) * n_steps + step)
    const long long step = row % n_steps;
    const int axis = (int)((row / n_steps) % 2), l = (int)(row / (2 * n_steps));
    const int chunk = (int)(step / J), j = (int)(step % J);
    const double sh = shifts[((size_t)l * 2 + axis) * J + j], last = shifts[((size_t)l * 2 + axis) * J + (J - 1)];
    const double dn = (double)n;
    for (int p = tid; p < np2; p += kThreads) {
        double x = INFINITY;
        if (p < P) {
            x = __dadd_rn((double)(lo + p), sh);
            for (int k = 0; k < chunk; ++k) x = __dadd_rn(x, last);
            double r = fmod(x, dn);                              // numpy's % on floats (npy_divmod)
            if (r != 0.0) {
                if (r < 0.0) r = __dadd_rn(r, dn);
            } else {
                r = 0.0;
            }
            x = r;
        }
        c[p] = x;
    }
    if (tid == 0) all_close = 1;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)                           // bitonic sort, ascending
        for (int s = k >> 1; s > 0; s >>= 1) {
            for (int i = tid; i < np2; i += kThreads) {
                const int q = i ^ s;
                if (q > i) {
                    const double a = c[i], b = c[q];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        c[i] = b;
                        c[q] = a;
                    }
                }
            }
            __syncthreads();
        }
    // first argmax of the gaps, and whether all of them are 1 within numpy.isclose (rtol 1e-5, atol 1e-8)
    double best = -1.0;
    int best_i = 0x7fffffff;
    bool close = true;
    for (int i = tid; i < P - 1; i += kThreads) {
        const double g = fabs(__dadd_rn(c[i + 1], -c[i]));
        if (g > best) {
            best = g;
            best_i = i;
        }
        if (!(fabs(__dadd_rn(g, -1.0)) <= 1e-8 + 1e-5)) close = false;
    }
    if (!close) all_close = 0;
    red_v[tid] = best;
    red_i[tid] = best_i;
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) {
        if (tid < s) {
            const double v = red_v[tid + s];
            const int i = red_i[tid + s];
            if (v > red_v[tid] || (v == red_v[tid] && i < red_i[tid])) {
                red_v[tid] = v;
                red_i[tid] = i;
            }
        }
        __syncthreads();
    }
    if (tid == 0) roll_s = (P < 2 || all_close) ? 0 : red_i[0];
    __syncthreads();
    const int roll = roll_s;
    int* oi = axis == 0 ? xi : yi;
    float* of = axis == 0 ? xf : yf;
    const size_t o = ((size_t)l * n_steps + step) * P;
    for (int p = tid; p < P; p += kThreads) {
        double at = c[(p + roll) % P];
        at = fmin(at, dn - 1.0);
        int i0 = (int)floor(at);
        if (i0 > n - 2) i0 = n - 2;
        oi[o + p] = i0;
        of[o + p] = (float)__dadd_rn(at, -(double)i0);
    }
}

}  // namespace
}  // namespace fastb

extern "C" int fastb_temporal_coords(int32_t n, int32_t n_pup, int32_t lo, int32_t n_layers, int32_t steps_per_chunk,
                                     int32_t n_chunks, const double* d_pixel_shifts, int32_t* d_xi, float* d_xf,
                                     int32_t* d_yi, float* d_yf, void* stream) {
    FASTB_REQUIRE(n >= 2 && n_pup >= 1 && n_pup <= n && lo >= 0 && lo + n_pup <= n, "fastb_temporal_coords: bad n / n_pup / lo");
    FASTB_REQUIRE(n_layers >= 1 && n_layers <= FASTB_MAX_LAYERS, "fastb_temporal_coords: bad n_layers");
    FASTB_REQUIRE(steps_per_chunk >= 1 && n_chunks >= 0, "fastb_temporal_coords: bad step counts");
    FASTB_REQUIRE(d_pixel_shifts && d_xi && d_xf && d_yi && d_yf, "fastb_temporal_coords: NULL pointer");
    const long long rows = 2LL * n_layers * steps_per_chunk * n_chunks;
    if (rows == 0) return FASTB_OK;
    FASTB_REQUIRE(rows < (1LL << 31), "fastb_temporal_coords: too many rows");
    int np2 = 1;
    while (np2 < n_pup) np2 <<= 1;
    const size_t smem = sizeof(double) * (size_t)np2;
    FASTB_CUDA(cudaFuncSetAttribute(temporal_coords_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    temporal_coords_kernel<<<(unsigned)rows, kThreads, smem, (cudaStream_t)stream>>>(
        n, n_pup, lo, n_layers, steps_per_chunk, n_chunks, d_pixel_shifts, d_xi, d_xf, d_yi, d_yf);
    return check_launch("temporal_coords_kernel");
}

extern "C" int fastb_temporal_detect(const FastbTemporalParams* p, const float* d_screens, const int32_t* d_xi,
                                     const float* d_xf, const int32_t* d_yi, const float* d_yf,
                                     const float* d_U, const float* d_chi, float* d_out, void* stream) {
    FASTB_REQUIRE(p, "fastb_temporal_detect: NULL params");
    FASTB_REQUIRE(p->n >= 2 && p->n_pup >= 1 && p->n_pup <= p->n, "fastb_temporal_detect: bad n / n_pup");
    FASTB_REQUIRE(p->n_layers >= 1 && p->n_layers <= FASTB_MAX_LAYERS, "fastb_temporal_detect: bad n_layers");
    FASTB_REQUIRE(p->n_steps >= 0 && p->u_sum != 0.0, "fastb_temporal_detect: bad n_steps / u_sum");
    FASTB_REQUIRE(d_screens && d_xi && d_xf && d_yi && d_yf && d_U && d_chi && d_out,
                  "fastb_temporal_detect: NULL pointer");
    if (p->n_steps == 0) return FASTB_OK;
    TemporalArgs a;
    a.n = p->n;
    a.n_pup = p->n_pup;
    a.n_layers = p->n_layers;
    a.coherent = p->coherent;
    a.n_steps = p->n_steps;
    a.inv_usum = (float)(1.0 / p->u_sum);
    a.screens = d_screens;
    a.xi = d_xi;
    a.xf = d_xf;
    a.yi = d_yi;
    a.yf = d_yf;
    a.U = d_U;
    a.chi = d_chi;
    a.out = d_out;
    long long grid = p->n_steps < 148LL * 8 ? p->n_steps : 148LL * 8;
    temporal_detect_kernel<<<(unsigned)grid, kThreads, 0, (cudaStream_t)stream>>>(a);
    return check_launch("temporal_detect_kernel");
}
