// K2 for grids that are not a power of two (any even N with N + n_pup - 1 <= 2048): the reference's
// NPXLS 'auto' rule (fast/fast.py:166-187) produces sizes such as 164.  Each pruned N-point line
// transform of the two passes is evaluated as a chirp-z (Bluestein) convolution on the register
// radix FFT of length M = 2^ceil(log2(N + n_pup - 1)):
//   X[k] = sum_n x[n] e^{+2 pi i n k / N} = c[k] sum_n (x[n] c[n]) conj(c[k - n]),   c[m] = e^{i pi m^2 / N}
// Only the n_pup outputs k in [lo, lo + n_pup) are wanted, so the kernel b[d] = conj(c[d]) is needed on
// d = k - n in (lo - N, lo + n_pup): N + n_pup - 1 consecutive values, alias-free on a circle of length M.
// With G the (inverse-sign, unnormalised) line FFT:  y = conj(G(conj(G(a) G(b)))) / M,  a = x c.
// G(b) / M and c are tabulated once per call in float64 (bluestein_tables_kernel).
// The input chirps of both passes are folded into a complex copy of the weight table (chirp_weight_kernel), every
// line generates and colours its own row (no CTA-wide barrier inside a pass).  Same scratch layout, RNG contract
// and epilogue as the radix kernel (screen_detect_kernel.cuh), including the fused sub-harmonic term; replaces
// the O(N^2)-per-line direct kernel on the default path.
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


// wc[item][r][j] = weight[item][r][j] * c[j] * c[r]: the input chirp of pass 1 (c[j]) and, because both passes are
// linear, the input chirp of pass 2 (c[r'], constant along a row) folded into one complex table, so that colouring a
// noise sample is a single complex multiply and pass 2 reads its inputs ready-made.  Phase in float64.
__global__ void chirp_weight_kernel(const float* __restrict__ w, float2* __restrict__ wc, int N, long long total) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const long long j = o % N, r = (o / N) % N;
    double s, c;
    sincospi((double)((j * j + r * r) % (2LL * N)) / (double)N, &s, &c);
    const float x = w[o];
    wc[o] = make_float2(x * (float)c, x * (float)s);
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
    double* st = reinterpret_cast<double*>(chirp + N + (N & 1));
    float* red = reinterpret_cast<float*>(st + kStatWords);
    float2* sh_amp = reinterpret_cast<float2*>(red + 4 * (THREADS / 32));     // SH only
    float2* sh_tab = sh_amp + 28;

    const int tid = threadIdx.x;
    const int ln = tid / S1, u = tid % S1;
    float2* buf = bufs + ln * F::kBuf;              // the line's exchange buffer; also stages its input row
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
    const int S = (N + 15) / 16;                    // noise blocks per row (include/fastb.h); S <= S1 because M >= N
    const int n1 = (N + LPB - 1) / LPB, n2 = (P + LPB - 1) / LPB;

    // wanted outputs of this thread: k = k_base + k_off(e) inside [lo, lo + P)
    const int kb = F::k_base(u) - lo;
    unsigned need = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e)
        if ((unsigned)(kb + F::k_off(e)) < (unsigned)P) need |= 1u << e;

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const PairId id = pair_id(a, pair);
        const unsigned long long g = id.g;
        const uint32_t k0 = (uint32_t)id.seed, k1 = (uint32_t)(id.seed >> 32);
        const float2* wc = reinterpret_cast<const float2*>(a.weight_s) + (size_t)id.item * N * N;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (SH) sh_prepare(a, pair, id, sh_amp, sh_tab);       // complete after the barrier between the passes

        // One loop body serves both passes (instruction-cache footprint): iterations [0, n1) are rows, [n1, n1 + n2)
        // kept columns.  Pass 1: every line makes its own row (no CTA-wide barrier inside a pass): thread u < S owns
        // noise block u of the row, colours its 16 cells with one complex multiply each and drops them into the line
        // buffer in natural order; the line then picks them up in the FFT's input layout.  Idle lines run on zeros
        // (warps stay converged).  Pass 2: T already carries the input chirp c[r'].
        for (int it = 0; it < n1 + n2; ++it) {
            const bool rows = it < n1;
            if (it == n1) __syncthreads();                     // every row of T is stored before a column is read
            const int line = (rows ? it : it - n1) * LPB + ln; // r' or c
            const bool live = line < (rows ? N : P);
            float2 v[16];
            if (rows) {
                if (live && u < S) {
                    // the 16 complex weights first: their L2 latency hides behind the Philox rounds
                    const float2* wrow = wc + (size_t)line * N + u;
                    float2 wv[16];
#pragma unroll
                    for (int m = 0; m < 16; ++m) wv[m] = (u + S * m < N) ? __ldg(wrow + S * m) : make_float2(0.f, 0.f);
                    if (RNG != kRngHost) {
                        uint32_t mr[16], ma[16];
                        if (RNG == kRngFast) noise_block_fields_fast((uint32_t)(line * S + u), g, k0, k1, mr, ma);
                        else noise_block_fields((uint32_t)(line * S + u), g, k0, k1, mr, ma);
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            const int j = u + S * m;
                            if (j < N) buf[j] = cmul(weighted_normal_m(mr[m], ma[m], 1.0f), wv[m]);
                        }
                    } else {
                        const float2* nrow = a.noise + ((size_t)pair * N + line) * N;
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            const int j = u + S * m;
                            if (j < N) buf[j] = cmul(__ldg(nrow + j), wv[m]);
                        }
                    }
                }
                sync();
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int n = u + S1 * m;
                    v[m] = (live && n < N) ? buf[n] : make_float2(0.f, 0.f);
                }
                sync();                                        // phase A of the transform rewrites the buffer
            } else {
                const float2* tcol = T + (size_t)(live ? line : 0) * N;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int n = u + S1 * m;
                    v[m] = (live && n < N) ? __ldcg(tcol + n) : make_float2(0.f, 0.f);
                }
            }

            chirp_convolve<F>(u, v, twa, twb, buf, bhat, sync);

            if (!live) continue;
            if (rows) {
                float2* tb = T + ((long long)kb * N + line);
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (need & (1u << e)) __stcg(tb + (long long)F::k_off(e) * N, cmul(v[e], chirp[kb + lo + F::k_off(e)]));
            } else {
                const float* ub = a.u_t + ((long long)line * P + kb);
                // output sign (-1)^(row + column) = (-1)^(k + c + lo); k_off is even
                const float sgn = ((F::k_base(u) + line + lo) & 1) ? -1.f : 1.f;
                float2 ex[3];
                if (SH) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) ex[i] = __ldg(a.sh_ex + i * P + line);
                }
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (need & (1u << e)) {
                        const float uu = __ldg(ub + F::k_off(e));
                        const float2 phi = cmul(v[e], chirp[kb + lo + F::k_off(e)]);
                        if (SH) {
                            const float2 sp = sh_phase(sh_tab + (kb + F::k_off(e)) * kShTab, ex);
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
    return sizeof(float2) * ((size_t)F::kTwA + F::kTwB + (size_t)LPB * F::kBuf + F::N + (size_t)n + (n & 1)) +
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

int bluestein_prepare_weights(int n, int n_items, const float* weight, void* wc, cudaStream_t st) {
    const long long total = (long long)n * n * n_items;
    chirp_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(weight, (float2*)wc, n, total);
    return check_launch("chirp_weight_kernel");
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
