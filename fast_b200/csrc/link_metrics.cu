// K5: link metrics on the per-realisation results -- error-probability curves, fade statistics,
// the Monte-Carlo modulator, AWGN-convolved I-Q histograms and mutual information.
// Contract: include/fastb.h (K5).  Reference: fast/comms.py.
//
// Everything here is a streaming pass over n float32 samples (HBM/L2-bound, n = 1e5..1e7) or a
// small dense float64 problem (npxls^3 per symbol); float64 keeps parity with the reference's
// numpy arithmetic.  Sums are reduced warp -> block -> one float64 atomic per block.
#include "fastb_common.cuh"

namespace fastb {
namespace {

constexpr int kBlock = 256;
constexpr uint32_t kStreamMod = 0x30D0A700u;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of `v`; the result is valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* red /* kBlock/32 */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                         // red may still be read from a previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kBlock / 32; ++w) t += red[w];
    return t;
}

int grid_for(long long n, int cap = 1184) {
    long long b = (n + kBlock - 1) / kBlock;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__global__ void scale_kernel(double* v, double f, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) v[i] *= f;
}

// ---- error curves ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) sum_kernel(const float* __restrict__ x, long long n, double* out) {
    __shared__ double red[kBlock / 32];
    double s = 0;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
        s += (double)x[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

__device__ __forceinline__ double q_func(double x) { return 0.5 * erfc(x / sqrt(2.0)); }

__global__ void __launch_bounds__(kBlock) error_curve_kernel(const float* __restrict__ x, long long n, int kind,
                                                             int M, const double* __restrict__ snr_db,
                                                             const double* mean, double* curve) {
    __shared__ double red[kBlock / 32];
    const int j = blockIdx.y;
    const double frac = pow(10.0, snr_db[j] / 10.0);
    const double snr = sqrt(frac);
    const double mu = *mean;
    const double a = (sqrt((double)M) - 1.0) / sqrt((double)M), c3 = 3.0 / ((double)M - 1.0);
    double acc = 0;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const double s = (double)x[i] / mu;
        if (kind == FASTB_CURVE_BER_OOK) {
            acc += q_func(s * snr);
        } else {
            const double q = q_func(sqrt(c3 * (frac * (s * s))));
            acc += 4.0 * (a * q - a * a * q * q);
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(&curve[j], acc);
}

// ---- fade statistics ------------------------------------------------------------------------
// per threshold: o[0] += below, o[1] += 0->1 transitions at i >= 1, o[2] = min index not below,
// o[3] = max index not below
__global__ void __launch_bounds__(kBlock) fade_kernel(const float* __restrict__ x, long long n,
                                                      const double* __restrict__ thr, long long* out) {
    __shared__ long long red[4][kBlock / 32];
    const double t = thr[blockIdx.y];
    long long below = 0, starts = 0, first = n, last = -1;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const bool m = (double)x[i] < t;
        if (m) {
            ++below;
            if (i >= 1 && !((double)x[i - 1] < t)) ++starts;
        } else {
            first = i < first ? i : first;
            last = i > last ? i : last;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        below += __shfl_xor_sync(0xffffffffu, below, o);
        starts += __shfl_xor_sync(0xffffffffu, starts, o);
        const long long f2 = __shfl_xor_sync(0xffffffffu, first, o), l2 = __shfl_xor_sync(0xffffffffu, last, o);
        first = f2 < first ? f2 : first;
        last = l2 > last ? l2 : last;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = below;
        red[1][warp] = starts;
        red[2][warp] = first;
        red[3][warp] = last;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kBlock / 32; ++w) {
            below += red[0][w];
            starts += red[1][w];
            first = red[2][w] < first ? red[2][w] : first;
            last = red[3][w] > last ? red[3][w] : last;
        }
        long long* o = out + 4 * blockIdx.y;
        atomicAdd((unsigned long long*)&o[0], (unsigned long long)below);
        atomicAdd((unsigned long long*)&o[1], (unsigned long long)starts);
        atomicMin(&o[2], first);
        atomicMax(&o[3], last);
    }
}

__global__ void fade_init_kernel(long long* out, int k, long long n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) {
        out[4 * j] = 0;
        out[4 * j + 1] = 0;
        out[4 * j + 2] = n;
        out[4 * j + 3] = -1;
    }
}

__global__ void fade_finish_kernel(long long* out, int k, long long n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    long long* o = out + 4 * j;
    const long long below = o[0], starts = o[1], first = o[2], last = o[3];
    long long fades = 0, inside = 0;
    if (last >= 0) {                                  // some sample is not fading
        const long long lead = first, trail = n - 1 - last;
        fades = starts - (trail > 0 ? 1 : 0);         // the run still in progress at the end
        inside = below - lead - trail;
    }
    o[1] = fades;
    o[2] = inside;
}

// ---- Monte-Carlo modulator ------------------------------------------------------------------
struct ModArgs {
    FastbModParams p;
    const float* power;
    const float2* pts;
    double* sums;
    uint8_t* symbols;
    float2* recv;
    uint8_t* recv_symbols;
    const uint8_t* tx_symbols;
    int chunk;                 // symbol slots per blockIdx.y
};

__global__ void __launch_bounds__(kBlock) modulator_kernel(const __grid_constant__ ModArgs a) {
    extern __shared__ float2 pts[];
    __shared__ double red[kBlock / 32];
    const int ns = a.p.n_symbols;
    for (int j = threadIdx.x; j < ns; j += kBlock) pts[j] = a.pts[j];
    __syncthreads();
    const uint32_t k0 = (uint32_t)a.p.seed, k1 = (uint32_t)(a.p.seed >> 32);
    const int s0 = blockIdx.y * a.chunk;
    const int s1 = min(s0 + a.chunk, a.p.symbols_per_iter);
    const float es = (float)a.p.es;
    double errs = 0, dev = 0, tx2 = 0;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < a.p.n; i += (long long)gridDim.x * kBlock) {
        const unsigned long long gi = (unsigned long long)(a.p.first + i);
        float sd = 0.f;
        if (a.p.has_awgn) {
            const float snr = (float)(a.p.snr_scale * (double)a.power[i]);
            sd = (a.p.scheme == FASTB_MOD_OOK ? es : sqrtf(0.5f * es)) / snr;
        }
        unsigned e = 0;
        float d_acc = 0.f, t_acc = 0.f;
        for (int s = s0; s < s1; ++s) {
            const uint4 w = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)s, kStreamMod, k0, k1);
            const int sym = a.tx_symbols ? (int)a.tx_symbols[s] : (int)__umulhi(w.x, (uint32_t)ns);
            const float2 tx = pts[sym];
            float2 nz = box_muller(w.y, w.z);
            nz.x *= sd;
            nz.y = (a.p.scheme == FASTB_MOD_OOK) ? 0.f : nz.y * sd;
            const float2 rx = make_float2(tx.x + nz.x, tx.y + nz.y);
            int dec;
            if (a.p.scheme == FASTB_MOD_OOK) dec = rx.x > 0.5f;
            else if (a.p.scheme == FASTB_MOD_BPSK) dec = rx.x < 0.f;
            else {
                float best = INFINITY;
                dec = 0;
                for (int c = 0; c < ns; ++c) {
                    const float dx = rx.x - pts[c].x, dy = rx.y - pts[c].y;
                    const float d2 = dx * dx + dy * dy;
                    if (d2 < best) {
                        best = d2;
                        dec = c;
                    }
                }
            }
            e += dec != sym;
            d_acc += sqrtf(nz.x * nz.x + nz.y * nz.y);
            t_acc += tx.x * tx.x + tx.y * tx.y;
            const long long o = (long long)s * a.p.n + i;
            if (a.symbols) a.symbols[o] = (uint8_t)sym;
            if (a.recv) a.recv[o] = rx;
            if (a.recv_symbols) a.recv_symbols[o] = (uint8_t)dec;
        }
        errs += e;
        dev += d_acc;
        tx2 += t_acc;
    }
    errs = block_sum(errs, red);
    dev = block_sum(dev, red);
    tx2 = block_sum(tx2, red);
    if (threadIdx.x == 0) {
        atomicAdd(&a.sums[0], errs);
        atomicAdd(&a.sums[1], dev);
        atomicAdd(&a.sums[2], tx2);
    }
}

// ---- I-Q histograms ---------------------------------------------------------------------------
// numpy.histogramdd's rule on explicit edges: bin = #(edges <= v) - 1, the last edge inclusive
__device__ __forceinline__ int find_bin(const double* __restrict__ e, int npxls, double v) {
    if (!(v >= e[0]) || v > e[npxls]) return -1;
    if (v == e[npxls]) return npxls - 1;
    int lo = 0, hi = npxls;                       // invariant: e[lo] <= v < e[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (e[mid] <= v) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kBlock) amplitude_kernel(const float* __restrict__ x, int is_complex, long long n,
                                                           double* amp, double* sums) {
    __shared__ double red[kBlock / 32];
    double s = 0, s2 = 0;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const double v = is_complex ? hypot((double)x[2 * i], (double)x[2 * i + 1]) : fabs((double)x[i]);
        amp[i] = v;
        s += v;
        s2 += v * v;
    }
    s = block_sum(s, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[0], s);
        atomicAdd(&sums[1], s2);
    }
}

__global__ void __launch_bounds__(kBlock) iq_histogram_kernel(const double* __restrict__ amp, long long n,
                                                              const double* __restrict__ pts, int npxls,
                                                              const double* __restrict__ ex,
                                                              const double* __restrict__ ey, uint32_t* counts) {
    const int c = blockIdx.y;
    const double pr = pts[2 * c], pi = pts[2 * c + 1];
    const double* exc = ex + (size_t)c * (npxls + 1);
    const double* eyc = ey + (size_t)c * (npxls + 1);
    uint32_t* cc = counts + (size_t)c * npxls * npxls;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const double a = amp[i];
        const int bx = find_bin(exc, npxls, pr * a), by = find_bin(eyc, npxls, pi * a);
        if (bx >= 0 && by >= 0) atomicAdd(&cc[bx * npxls + by], 1u);
    }
}

// t[c][j][l] = sum_k h[c][j][k] taps[k - l + K/2]      (correlation along the second axis)
__global__ void __launch_bounds__(kBlock) iq_corr_axis1_kernel(const uint32_t* __restrict__ counts, double inv_n,
                                                               int npxls, const double* __restrict__ taps,
                                                               double* __restrict__ t, long long total) {
    const long long o = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (o >= total) return;
    const int l = (int)(o % npxls);
    const long long row = o / npxls;                // c * npxls + j
    const uint32_t* h = counts + row * npxls;
    const int half = (npxls + 1) / 2;
    double acc = 0;
    for (int k = 0; k < npxls; ++k) {
        const int q = k - l + half;
        const uint32_t cnt = h[k];
        if (cnt && q >= 0 && q <= npxls) acc += ((double)cnt * inv_n) * taps[q];
    }
    t[o] = acc;
}

// out[c][i][l] = sum_j taps[j - i + K/2] t[c][j][l]    (correlation along the first axis)
__global__ void __launch_bounds__(kBlock) iq_corr_axis0_kernel(const double* __restrict__ t, int npxls,
                                                               const double* __restrict__ taps,
                                                               double* __restrict__ out, long long total) {
    const long long o = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (o >= total) return;
    const int l = (int)(o % npxls);
    const int i = (int)((o / npxls) % npxls);
    const long long c = o / ((long long)npxls * npxls);
    const double* tc = t + c * npxls * npxls + l;
    const int half = (npxls + 1) / 2;
    double acc = 0;
    for (int j = 0; j < npxls; ++j) {
        const int q = j - i + half;
        if (q >= 0 && q <= npxls) acc += taps[q] * tc[(long long)j * npxls];
    }
    out[o] = acc;
}

// shot-noise variant, step 1: list the occupied bins of every symbol
struct ShotBin {
    int i, j;
    double h;
};
__global__ void __launch_bounds__(kBlock) iq_compact_kernel(const uint32_t* __restrict__ counts, double inv_n,
                                                            int npxls, long long total, ShotBin* list,
                                                            unsigned* nnz) {
    const long long o = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (o >= total) return;
    const uint32_t cnt = counts[o];
    if (!cnt) return;
    const long long per = (long long)npxls * npxls;
    const long long c = o / per;
    const int rem = (int)(o % per);
    const unsigned slot = atomicAdd(&nnz[c], 1u);
    list[c * per + slot] = ShotBin{rem / npxls, rem % npxls, (double)cnt * inv_n};
}

// step 2: every output pixel sums the Gaussians of the occupied bins
__global__ void __launch_bounds__(kBlock) iq_shot_kernel(const ShotBin* __restrict__ list,
                                                         const unsigned* __restrict__ nnz, int npxls,
                                                         double sigma2, double mean_amp,
                                                         const double* __restrict__ ex, const double* __restrict__ ey,
                                                         double* __restrict__ out, long long total) {
    const long long o = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (o >= total) return;
    const long long per = (long long)npxls * npxls;
    const long long c = o / per;
    const int rem = (int)(o % per);
    const int y = rem / npxls, x = rem % npxls;
    const double* exc = ex + c * (npxls + 1);
    const double* eyc = ey + c * (npxls + 1);
    const ShotBin* lc = list + c * per;
    const unsigned cnt = nnz[c];
    double acc = 0;
    for (unsigned q = 0; q < cnt; ++q) {
        const ShotBin b = lc[q];
        const double mult = mean_amp * mean_amp / (exc[b.i] * exc[b.i] + eyc[b.j] * eyc[b.j]);
        const double s2m = sigma2 * mult;
        const double dx = (double)(b.j - x), dy = (double)(b.i - y);
        acc += b.h * exp(-(dx * dx + dy * dy) / s2m) / (M_PI * s2m);
    }
    out[o] = acc;
}

// ---- mutual information -------------------------------------------------------------------
__device__ __forceinline__ double xlog_ratio(double f, double log2_fy, bool fy_ok) {
    return (f > 0.0 && fy_ok) ? f * (log2(f) - log2_fy) : 0.0;
}

constexpr int kMaxBits = 16;

__global__ void __launch_bounds__(kBlock) iq_information_kernel(const double* __restrict__ f, int M, long long npix,
                                                                const uint32_t* __restrict__ gray, int n_bits,
                                                                double* out) {
    __shared__ double red[kBlock / 32];
    double mi = 0, gmi = 0;
    for (long long p = (long long)blockIdx.x * kBlock + threadIdx.x; p < npix; p += (long long)gridDim.x * kBlock) {
        double tot = 0;
        double s0[kMaxBits], s1[kMaxBits];
        int c0[kMaxBits];
#pragma unroll
        for (int b = 0; b < kMaxBits; ++b) {
            s0[b] = 0;
            s1[b] = 0;
            c0[b] = 0;
        }
        for (int c = 0; c < M; ++c) {
            const double v = f[(long long)c * npix + p];
            const uint32_t g = gray[c];
            tot += v;
#pragma unroll
            for (int b = 0; b < kMaxBits; ++b) {
                if (b < n_bits) {
                    if ((g >> (n_bits - 1 - b)) & 1u) s1[b] += v;
                    else {
                        s0[b] += v;
                        ++c0[b];
                    }
                }
            }
        }
        const double fy = tot / (double)M;
        const bool fy_ok = fy > 0.0;
        const double l2fy = fy_ok ? log2(fy) : 0.0;
        double m = 0;
        for (int c = 0; c < M; ++c) m += xlog_ratio(f[(long long)c * npix + p], l2fy, fy_ok);
        mi += m;
#pragma unroll
        for (int b = 0; b < kMaxBits; ++b) {
            if (b < n_bits) {
                const double f0 = c0[b] ? s0[b] / (double)c0[b] : 0.0;
                const double f1 = (M - c0[b]) ? s1[b] / (double)(M - c0[b]) : 0.0;
                gmi += 0.5 * (xlog_ratio(f0, l2fy, fy_ok) + xlog_ratio(f1, l2fy, fy_ok));
            }
        }
    }
    mi = block_sum(mi, red);
    gmi = block_sum(gmi, red);
    if (threadIdx.x == 0) {
        atomicAdd(&out[0], mi / (double)M);
        atomicAdd(&out[1], gmi);
    }
}

}  // namespace
}  // namespace fastb

using namespace fastb;

extern "C" int fastb_error_curve(const float* d_samples, int64_t n, int32_t kind, int32_t qam_order,
                                 const double* d_snr_db, int32_t k, double* d_curve, void* stream) {
    FASTB_REQUIRE(d_samples && d_snr_db && d_curve, "fastb_error_curve: NULL pointer");
    FASTB_REQUIRE(n >= 1 && k >= 1 && k <= 65535, "fastb_error_curve: need n >= 1 and 1 <= k <= 65535");
    FASTB_REQUIRE(kind == FASTB_CURVE_BER_OOK || (kind == FASTB_CURVE_SEP_QAM && qam_order >= 4),
                  "fastb_error_curve: kind must be BER_OOK or SEP_QAM with qam_order >= 4");
    cudaStream_t st = (cudaStream_t)stream;
    FASTB_CUDA(cudaMemsetAsync(d_curve, 0, sizeof(double) * (size_t)(k + 1), st));
    sum_kernel<<<grid_for(n), kBlock, 0, st>>>(d_samples, n, d_curve + k);
    int rc = check_launch("sum_kernel");
    if (rc) return rc;
    scale_kernel<<<1, 32, 0, st>>>(d_curve + k, 1.0 / (double)n, 1);
    if ((rc = check_launch("scale_kernel"))) return rc;
    const dim3 grid((unsigned)grid_for(n, 592), (unsigned)k);
    error_curve_kernel<<<grid, kBlock, 0, st>>>(d_samples, n, kind, qam_order, d_snr_db, d_curve + k, d_curve);
    if ((rc = check_launch("error_curve_kernel"))) return rc;
    scale_kernel<<<(k + 255) / 256, 256, 0, st>>>(d_curve, 1.0 / (double)n, k);
    return check_launch("scale_kernel");
}

extern "C" int fastb_fade_stats(const float* d_series, int64_t n, const double* d_thresholds, int32_t k,
                                int64_t* d_out, void* stream) {
    FASTB_REQUIRE(d_series && d_thresholds && d_out, "fastb_fade_stats: NULL pointer");
    FASTB_REQUIRE(n >= 1 && k >= 1 && k <= 65535, "fastb_fade_stats: need n >= 1 and 1 <= k <= 65535");
    cudaStream_t st = (cudaStream_t)stream;
    long long* out = reinterpret_cast<long long*>(d_out);
    fade_init_kernel<<<(k + 255) / 256, 256, 0, st>>>(out, k, n);
    int rc = check_launch("fade_init_kernel");
    if (rc) return rc;
    const dim3 grid((unsigned)grid_for(n, 592), (unsigned)k);
    fade_kernel<<<grid, kBlock, 0, st>>>(d_series, n, d_thresholds, out);
    if ((rc = check_launch("fade_kernel"))) return rc;
    fade_finish_kernel<<<(k + 255) / 256, 256, 0, st>>>(out, k, n);
    return check_launch("fade_finish_kernel");
}

extern "C" int fastb_modulator_mc(const FastbModParams* p, const float* d_power, const float* d_constellation,
                                  double* d_sums, uint8_t* d_symbols, float* d_recv, uint8_t* d_recv_symbols,
                                  const uint8_t* d_tx_symbols, void* stream) {
    FASTB_REQUIRE(p && d_power && d_constellation && d_sums, "fastb_modulator_mc: NULL pointer");
    FASTB_REQUIRE(p->n >= 1 && p->symbols_per_iter >= 1, "fastb_modulator_mc: empty problem");
    FASTB_REQUIRE(p->n_symbols >= 2 && p->n_symbols <= 1024, "fastb_modulator_mc: n_symbols must be 2..1024");
    FASTB_REQUIRE(p->scheme >= FASTB_MOD_OOK && p->scheme <= FASTB_MOD_NEAREST, "fastb_modulator_mc: bad scheme");
    FASTB_REQUIRE((p->scheme == FASTB_MOD_NEAREST) || p->n_symbols == 2,
                  "fastb_modulator_mc: OOK / BPSK have two symbols");
    FASTB_REQUIRE(!(d_symbols || d_recv_symbols || d_tx_symbols) || p->n_symbols <= 256,
                  "fastb_modulator_mc: symbol outputs are uint8 (n_symbols <= 256)");
    FASTB_REQUIRE(!p->has_awgn || (p->snr_scale > 0.0 && p->es >= 0.0), "fastb_modulator_mc: bad SNR scale");
    ModArgs a;
    a.p = *p;
    a.power = d_power;
    a.pts = reinterpret_cast<const float2*>(d_constellation);
    a.sums = d_sums;
    a.symbols = d_symbols;
    a.recv = reinterpret_cast<float2*>(d_recv);
    a.recv_symbols = d_recv_symbols;
    a.tx_symbols = d_tx_symbols;
    // enough CTAs to fill the GPU: split the symbol slots when there are few realisations
    const int gx = grid_for(p->n, 1184);
    int gy = (2368 + gx - 1) / gx;
    if (gy > p->symbols_per_iter) gy = p->symbols_per_iter;
    if (gy < 1) gy = 1;
    a.chunk = (p->symbols_per_iter + gy - 1) / gy;
    gy = (p->symbols_per_iter + a.chunk - 1) / a.chunk;
    modulator_kernel<<<dim3((unsigned)gx, (unsigned)gy), kBlock, sizeof(float2) * (size_t)p->n_symbols,
                       (cudaStream_t)stream>>>(a);
    return check_launch("modulator_kernel");
}

extern "C" int fastb_amplitudes(const float* d_samples, int32_t is_complex, int64_t n, double* d_amp,
                                double* d_sums, void* stream) {
    FASTB_REQUIRE(d_samples && d_amp && d_sums, "fastb_amplitudes: NULL pointer");
    FASTB_REQUIRE(n >= 1, "fastb_amplitudes: need n >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    FASTB_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * 2, st));
    amplitude_kernel<<<grid_for(n), kBlock, 0, st>>>(d_samples, is_complex, n, d_amp, d_sums);
    return check_launch("amplitude_kernel");
}

extern "C" int fastb_iq_histogram(const double* d_amp, int64_t n, const double* d_points, int32_t m,
                                  const double* d_edges_x, const double* d_edges_y, int32_t npxls,
                                  uint32_t* d_counts, void* stream) {
    FASTB_REQUIRE(d_amp && d_points && d_edges_x && d_edges_y && d_counts, "fastb_iq_histogram: NULL pointer");
    FASTB_REQUIRE(n >= 1 && m >= 1 && m <= 65535 && npxls >= 1 && npxls <= 4096,
                  "fastb_iq_histogram: need n >= 1, 1 <= m <= 65535, 1 <= npxls <= 4096");
    cudaStream_t st = (cudaStream_t)stream;
    FASTB_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * (size_t)m * npxls * npxls, st));
    const dim3 grid((unsigned)grid_for(n, 592), (unsigned)m);
    iq_histogram_kernel<<<grid, kBlock, 0, st>>>(d_amp, n, d_points, npxls, d_edges_x, d_edges_y, d_counts);
    return check_launch("iq_histogram_kernel");
}

extern "C" int64_t fastb_iq_convolve_workspace_bytes(int32_t m, int32_t npxls) {
    if (m < 1 || npxls < 1) return 0;
    const int64_t per = (int64_t)npxls * npxls;
    // shot variant: bin list (16 B per bin) + one counter per symbol; plain: one float64 plane set
    return (int64_t)m * per * (int64_t)sizeof(ShotBin) + 256 + (int64_t)sizeof(unsigned) * m;
}

extern "C" int fastb_iq_convolve(const uint32_t* d_counts, int64_t n, int32_t m, int32_t npxls,
                                 const double* d_taps, int32_t shot, double sigma2, double mean_amp,
                                 const double* d_edges_x, const double* d_edges_y, double* d_out,
                                 void* d_workspace, int64_t workspace_bytes, void* stream) {
    FASTB_REQUIRE(d_counts && d_out && d_workspace, "fastb_iq_convolve: NULL pointer");
    FASTB_REQUIRE(n >= 1 && m >= 1 && npxls >= 1 && npxls <= 4096, "fastb_iq_convolve: bad sizes");
    FASTB_REQUIRE(workspace_bytes >= fastb_iq_convolve_workspace_bytes(m, npxls),
                  "fastb_iq_convolve: workspace too small (%lld B)", (long long)workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const long long per = (long long)npxls * npxls, total = per * m;
    const unsigned blocks = (unsigned)((total + kBlock - 1) / kBlock);
    const double inv_n = 1.0 / (double)n;
    int rc;
    if (!shot) {
        FASTB_REQUIRE(d_taps, "fastb_iq_convolve: d_taps is required when shot == 0");
        double* t = reinterpret_cast<double*>(d_workspace);
        iq_corr_axis1_kernel<<<blocks, kBlock, 0, st>>>(d_counts, inv_n, npxls, d_taps, t, total);
        if ((rc = check_launch("iq_corr_axis1_kernel"))) return rc;
        iq_corr_axis0_kernel<<<blocks, kBlock, 0, st>>>(t, npxls, d_taps, d_out, total);
        return check_launch("iq_corr_axis0_kernel");
    }
    FASTB_REQUIRE(d_edges_x && d_edges_y && sigma2 > 0.0, "fastb_iq_convolve: shot variant needs edges and sigma2");
    ShotBin* list = reinterpret_cast<ShotBin*>(d_workspace);
    unsigned* nnz = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(d_workspace) +
                                                (((size_t)total * sizeof(ShotBin) + 255) / 256) * 256);
    FASTB_CUDA(cudaMemsetAsync(nnz, 0, sizeof(unsigned) * (size_t)m, st));
    iq_compact_kernel<<<blocks, kBlock, 0, st>>>(d_counts, inv_n, npxls, total, list, nnz);
    if ((rc = check_launch("iq_compact_kernel"))) return rc;
    iq_shot_kernel<<<blocks, kBlock, 0, st>>>(list, nnz, npxls, sigma2, mean_amp, d_edges_x, d_edges_y, d_out,
                                              total);
    return check_launch("iq_shot_kernel");
}

extern "C" int fastb_iq_information(const double* d_f, int32_t m, int32_t npxls, const uint32_t* d_gray,
                                    int32_t n_bits, double* d_out, void* stream) {
    FASTB_REQUIRE(d_f && d_gray && d_out, "fastb_iq_information: NULL pointer");
    FASTB_REQUIRE(m >= 1 && npxls >= 1 && n_bits >= 0 && n_bits <= kMaxBits,
                  "fastb_iq_information: need m, npxls >= 1 and 0 <= n_bits <= %d", kMaxBits);
    cudaStream_t st = (cudaStream_t)stream;
    FASTB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 2, st));
    const long long npix = (long long)npxls * npxls;
    iq_information_kernel<<<grid_for(npix, 592), kBlock, 0, st>>>(d_f, m, npix, d_gray, n_bits, d_out);
    return check_launch("iq_information_kernel");
}
