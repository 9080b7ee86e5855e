// K2 for grids that are not a power of two (any even N with N + n_pup - 1 <= 2048): the reference's
// NPXLS 'auto' rule (fast/fast.py:166-187) produces sizes such as 164.  Each pruned N-point line
// transform of the two passes is evaluated as a chirp-z (Bluestein) convolution on the register
// radix FFT of length M = 2^ceil(log2(N + n_pup - 1)):
//   X[k] = sum_n x[n] e^{+2 pi i n k / N} = c[k] sum_n (x[n] c[n]) conj(c[k - n]),   c[m] = e^{i pi m^2 / N}
// Only the n_pup outputs k in [lo, lo + n_pup) are wanted, so the kernel b[d] = conj(c[d]) is needed on
// d = k - n in (lo - N, lo + n_pup): N + n_pup - 1 consecutive values, alias-free on a circle of length M.
// With G the (inverse-sign, unnormalised) line FFT:  y = conj(G(conj(G(a) G(b)))) / M,  a = x c.
// G(b) / M and c are tabulated once per call in float64 (bluestein_tables_kernel).
// Same pass structure, scratch layout, RNG contract and epilogue as the radix kernel
// (screen_detect_kernel.cuh), including the fused sub-harmonic term; replaces the O(N^2)-per-line direct
// kernel on the default path.
#include "screen_detect_kernel.cuh"
#include "bluestein.cuh"

namespace fastb {
namespace {

int blue_log2m(int n, int n_pup) {
    int l = 6;
    while ((1 << l) < n + n_pup - 1) ++l;
    return l;
}

// tables: chirp[N] then bhat[M] (float2 each)
__global__ void __launch_bounds__(256) bluestein_tables_kernel(int N, int M, int lo, int P, float2* chirp,
                                                               float2* bhat) {
    __shared__ double red[2][8];
    const int q = blockIdx.x;
    if (q >= M) {                                  // the last blocks write the chirp
        for (int n = (q - M) * 256 + threadIdx.x; n < N; n += 256 * (gridDim.x - M)) {
            double s, c;
            sincospi((double)(((long long)n * n) % (2LL * N)) / (double)N, &s, &c);
            chirp[n] = make_float2((float)c, (float)s);
        }
        return;
    }
    // bhat[q] = (1/M) sum_j b[j] e^{+2 pi i j q / M},  b[j] = conj(c[d]) for the d in (lo - N, lo + P)
    // congruent to j mod M, 0 when there is none
    double sr = 0.0, si = 0.0;
    const int d_lo = lo - N + 1, d_hi = lo + P - 1;
    for (int d = d_lo + threadIdx.x; d <= d_hi; d += 256) {
        const int j = ((d % M) + M) % M;
        const long long d2 = ((long long)d * d) % (2LL * N);
        // phase / pi = -d^2 / N + 2 j q / M
        const double ph = -(double)d2 / (double)N + 2.0 * (double)(((long long)j * q) % M) / (double)M;
        double s, c;
        sincospi(ph, &s, &c);
        sr += c;
        si += s;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sr;
        red[1][threadIdx.x >> 5] = si;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, ti = 0.0;
        for (int w = 0; w < 8; ++w) {
            tr += red[0][w];
            ti += red[1][w];
        }
        bhat[q] = make_float2((float)(tr / M), (float)(ti / M));
    }
}


template <int LOG2M, int RNG, bool SH, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS <= 128 ? 4 : 2) screen_detect_bluestein(const __grid_constant__ RunArgs a,
                                                                      const float2* __restrict__ tables) {
    using F = LineFFT<LOG2M>;
    using Tw = typename F::Tw;
    constexpr int M = F::N, S1 = F::S1, LPB = THREADS / S1;
    static_assert(THREADS % S1 == 0 && LPB >= 1 && (S1 <= 32 || LPB <= 15), "line/barrier layout");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.n, P = a.n_pup, lo = a.lo;
    Tw* twa = reinterpret_cast<Tw*>(smem_raw);
    Tw* twb = twa + F::kTwA;
    float2* bufs = reinterpret_cast<float2*>(twb + F::kTwB);
    float2* bhat = bufs + LPB * F::kBuf;            // M
    float2* chirp = bhat + M;                       // N
    float2* rows = chirp + N;                       // LPB x N staged inputs x[n] c[n]
    double* st = reinterpret_cast<double*>(rows + (size_t)LPB * N);
    float* red = reinterpret_cast<float*>(st + kStatWords);
    float2* sh_amp = reinterpret_cast<float2*>(red + 4 * (THREADS / 32));     // SH only
    float2* sh_tab = sh_amp + 28;

    const int tid = threadIdx.x;
    const int ln = tid / S1, u = tid % S1;
    float2* buf = bufs + ln * F::kBuf;
    const LineSync<S1> sync{ln};

    for (int j = tid; j < F::kTwA + F::kTwB; j += THREADS) {
        const int ex = j < F::kTwA ? F::twa_exponent(j) : F::twb_exponent(j - F::kTwA);
        double s, c;
        sincospi(2.0 * (double)ex / (double)M, &s, &c);
        twa[j] = make_tw((float)c, (float)s, (Tw*)nullptr);
    }
    for (int j = tid; j < N; j += THREADS) chirp[j] = tables[j];
    for (int j = tid; j < M; j += THREADS) bhat[j] = tables[N + j];
    if (tid == 0) stats_reset(st, 0);
    __syncthreads();

    float2* T = a.scratch + (size_t)blockIdx.x * N * P;
    const int S = (N + 15) / 16;                    // noise blocks per row (include/fastb.h)

    // wanted outputs of this thread: k = k_base + k_off(e) inside [lo, lo + P)
    const int kb = F::k_base(u) - lo;
    unsigned need = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e)
        if ((unsigned)(kb + F::k_off(e)) < (unsigned)P) need |= 1u << e;

    // chirp-z of the line held as v[m] = a[u + S1 m] (a = x c, zero beyond N): on return v[e] holds
    // sum_n a[n] conj(c[k - n]) for k = k_out(u, e), valid where `need` says so
    auto convolve = [&](float2 (&v)[16]) { chirp_convolve<F>(u, v, twa, twb, buf, bhat, sync); };

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const PairId id = pair_id(a, pair);
        const unsigned long long g = id.g;
        const uint32_t k0 = (uint32_t)id.seed, k1 = (uint32_t)(id.seed >> 32);
        const float* weight = a.weight + (size_t)id.item * N * N;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (SH) sh_prepare(a, pair, id, sh_amp, sh_tab);       // table complete after the barriers of pass 1

        // ---- pass 1: LPB rows at a time
        for (int row0 = 0; row0 < N; row0 += LPB) {
            const int nr = min(LPB, N - row0);
            __syncthreads();                                   // the previous group's lines are loaded
            if (RNG != kRngHost) {
                for (int idx = tid; idx < nr * S; idx += THREADS) {
                    const int rl = idx / S, t = idx % S, r = row0 + rl;
                    uint32_t mr[16], ma[16];
                    if (RNG == kRngFast) noise_block_fields_fast((uint32_t)(r * S + t), g, k0, k1, mr, ma);
                    else noise_block_fields((uint32_t)(r * S + t), g, k0, k1, mr, ma);
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int j = t + S * m;
                        if (j < N)
                            rows[rl * N + j] = cmulf(weighted_normal_m(mr[m], ma[m], __ldg(weight + (size_t)r * N + j)), chirp[j]);
                    }
                }
            } else {
                for (int idx = tid; idx < nr * N; idx += THREADS) {
                    const int rl = idx / N, j = idx % N, r = row0 + rl;
                    const float2 nz = a.noise[((size_t)pair * N + r) * N + j];
                    const float w0 = weight[(size_t)r * N + j];
                    rows[rl * N + j] = cmulf(make_float2(nz.x * w0, nz.y * w0), chirp[j]);
                }
            }
            __syncthreads();
            // every line runs (idle ones on zeros) so that warps stay converged
            const int r = row0 + ln;
            float2 v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int n = u + S1 * m;
                v[m] = (n < N && r < N) ? rows[ln * N + n] : make_float2(0.f, 0.f);
            }
            convolve(v);
            if (r < N) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (need & (1u << e)) {
                        const int k = F::k_base(u) + F::k_off(e);
                        __stcg(T + ((long long)(k - lo) * N + r), cmulf(v[e], chirp[k]));
                    }
            }
        }
        __syncthreads();                                       // every row of T is stored

        // ---- pass 2: LPB kept columns at a time
        for (int col0 = 0; col0 < P; col0 += LPB) {
            const int c = col0 + ln;
            float2 v[16];
            const float2* tcol = T + (size_t)(c < P ? c : 0) * N;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int n = u + S1 * m;
                v[m] = (n < N && c < P) ? cmulf(__ldcg(tcol + n), chirp[n]) : make_float2(0.f, 0.f);
            }
            convolve(v);
            if (c < P) {
                const float* ub = a.u_t + ((long long)c * P + kb);
                // output sign (-1)^(row + column) = (-1)^(k + c + lo); k_off is even
                const float sgn = ((F::k_base(u) + c + lo) & 1) ? -1.f : 1.f;
                float2 ex[3];
                if (SH) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) ex[i] = __ldg(a.sh_ex + i * P + c);
                }
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (need & (1u << e)) {
                        const int k = F::k_base(u) + F::k_off(e);
                        const float uu = __ldg(ub + F::k_off(e));
                        const float2 phi = cmulf(v[e], chirp[k]);
                        if (SH) {
                            const float2 sp = sh_phase(sh_tab + (k - lo) * kShTab, ex);
                            accumulate(make_float2(fmaf(sgn, phi.x, sp.x), fmaf(sgn, phi.y, sp.y)), uu, uu, acc);
                        } else {
                            accumulate(phi, uu, uu * sgn, acc);
                        }
                    }
            }
        }
        finish_pair<THREADS>(a, pair, id, acc, red, st);
    }
    if (a.st_sums && tid == 0) stats_flush(a, st);
}

template <int LOG2M>
struct BlueCfg {
    static constexpr int kThreadsPerCta = LOG2M <= 8 ? 128 : 256;
};

template <int LOG2M>
size_t blue_smem_bytes(int n, bool sh, int n_pup) {
    using F = LineFFT<LOG2M>;
    constexpr int T = BlueCfg<LOG2M>::kThreadsPerCta, LPB = T / F::S1;
    return sizeof(float2) * ((size_t)F::kTwA + F::kTwB + (size_t)LPB * F::kBuf + F::N + (size_t)n + (size_t)LPB * n) +
           sizeof(double) * kStatWords + sizeof(float) * 4 * (T / 32) + sh_smem_bytes(sh, n_pup);
}

template <int LOG2M>
int launch_blue(const RunArgs& a, const RadixRequest& rq, const float2* tables, cudaStream_t st) {
    constexpr int T = BlueCfg<LOG2M>::kThreadsPerCta;
    const bool sh = a.sh_weight != nullptr;
    void (*kern)(RunArgs, const float2*) = nullptr;
    if (sh) kern = rq.rng == kRngHost   ? screen_detect_bluestein<LOG2M, kRngHost, true, T>
                   : rq.rng == kRngFast ? screen_detect_bluestein<LOG2M, kRngFast, true, T>
                                        : screen_detect_bluestein<LOG2M, kRngPhilox, true, T>;
    else kern = rq.rng == kRngHost   ? screen_detect_bluestein<LOG2M, kRngHost, false, T>
                : rq.rng == kRngFast ? screen_detect_bluestein<LOG2M, kRngFast, false, T>
                                     : screen_detect_bluestein<LOG2M, kRngPhilox, false, T>;
    const size_t smem = blue_smem_bytes<LOG2M>(a.n, sh, a.n_pup);
    if (smem > 227 * 1024) {
        set_error("screen_detect_bluestein: N=%d needs %zu B of shared memory", a.n, smem);
        return FASTB_ERR_UNSUPPORTED;
    }
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0, sms = 0;
    FASTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
    if (per_sm < 1) {
        set_error("screen_detect_bluestein: kernel does not fit (smem %zu B)", smem);
        return FASTB_ERR_UNSUPPORTED;
    }
    if (per_sm > kMaxCtasPerSm) per_sm = kMaxCtasPerSm;
    const int rc = sm_count(&sms);
    if (rc) return rc;
    long long grid = (long long)per_sm * sms;
    if (grid > a.n_pairs) grid = a.n_pairs;
    if (grid > rq.max_grid) grid = rq.max_grid;
    kern<<<(unsigned)grid, T, smem, st>>>(a, tables);
    return check_launch("screen_detect_bluestein");
}

}  // namespace

int bluestein_log2m(int n, int n_pup) { return blue_log2m(n, n_pup); }

bool bluestein_ok(int n, int n_pup) { return n >= 4 && (n % 2) == 0 && n + n_pup - 1 <= 2048; }

int bluestein_ctas_per_sm(int n, int n_pup) { return blue_log2m(n, n_pup) <= 8 ? 4 : 2; }

size_t bluestein_table_bytes(int n, int n_pup) {
    return sizeof(float2) * ((size_t)n + ((size_t)1 << blue_log2m(n, n_pup)));
}

int bluestein_prepare(int n, int n_pup, int lo, void* tables, cudaStream_t st) {
    const int M = 1 << blue_log2m(n, n_pup);
    float2* t = (float2*)tables;
    bluestein_tables_kernel<<<M + 4, 256, 0, st>>>(n, M, lo, n_pup, t, t + n);
    return check_launch("bluestein_tables_kernel");
}

int launch_bluestein(const RunArgs& a, const RadixRequest& rq, const void* tables, cudaStream_t st) {
    const float2* t = (const float2*)tables;
    switch (blue_log2m(a.n, a.n_pup)) {
        case 6: return launch_blue<6>(a, rq, t, st);
        case 7: return launch_blue<7>(a, rq, t, st);
        case 8: return launch_blue<8>(a, rq, t, st);
        case 9: return launch_blue<9>(a, rq, t, st);
        case 10: return launch_blue<10>(a, rq, t, st);
        case 11: return launch_blue<11>(a, rq, t, st);
        default: break;
    }
    set_error("screen_detect_bluestein: N=%d, n_pup=%d out of range", a.n, a.n_pup);
    return FASTB_ERR_UNSUPPORTED;
}

}  // namespace fastb
