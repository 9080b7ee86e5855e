// K2 for grids that are not a power of two: the chirp-z kernels of ONE transform length M = 2^LOG2M,
// compiled once per LOG2M = 6..11 (-DFASTB_LOG2M=k, see build_fastb.py).  Geometry, dispatch and the
// contract: screen_detect_bluestein.cu, bluestein.cuh.
//
// Each pruned N-point line transform of the two passes is a chirp-z convolution on the register line FFT G
// of length M >= N + n_pup - 1 (fft_core.cuh):
//   X[lo + k'] = c[lo + k'] y[k'],   y = conj(G(conj(G(a) Bhat))),   a[n] = x[n] c[n],   c[m] = e^{i pi m^2 / N},
//   Bhat = G(b') / M,  b'[d] = conj(c[d + lo]) on d in (-N, n_pup)  (the kernel shifted so that the wanted
//   outputs are the FIRST n_pup of the circle).
// What makes it cheap (round 3 of the kernel; profiles/experiments_r02.txt, 13):
//   * the noise block of thread u of a line IS the thread's FFT input (noise stride S = S1 = M / 16,
//     include/fastb.h): Philox + Box-Muller + one complex multiply by the chirped weight land in the
//     registers the transform starts from -- no staging through shared memory, no idle lanes;
//   * a thread owns at most 2C of its 16 inputs (class C, bluestein.cuh): only those cells are generated
//     (5 instead of 6 Philox calls and 12 instead of 16 Box-Muller samples at N = 164), the zeros prune the
//     first butterflies, and the crop bound that comes with the class prunes the last stage of the second
//     transform, all at compile time;
//   * Bhat is tabulated in the register order of the transform (8 x 128-bit shared loads per thread), and for
//     M = 256 the output order of G equals its input order, so the two transforms chain in registers;
//   * both chirps of the input side are folded into the complex weight table.
#include "screen_detect_kernel.cuh"
#include "bluestein.cuh"

#ifndef FASTB_LOG2M
#error "compile with -DFASTB_LOG2M=6..11"
#endif
#define FASTB_CAT_(a, b) a##b
#define FASTB_CAT(a, b) FASTB_CAT_(a, b)

namespace fastb {
namespace {

constexpr int kLog2M = FASTB_LOG2M;
using F = LineFFT<kLog2M>;
constexpr int kM = F::N, kS1 = F::S1;
constexpr int kBlueThreads = kLog2M <= 8 ? 128 : 256;
constexpr int kRowB = 18;                     // float2 per Bhat row (16 used): 144-byte rows, conflict-free 128-bit loads

// ---- tables -------------------------------------------------------------------------------------------
// block b < 16 S1: Bhat entry (u, e) = (b / 16, b % 16); the blocks after that write the output chirp
__global__ void __launch_bounds__(256) blue_tables_kernel(int N, int lo, int P, float2* __restrict__ bhatp,
                                                          float2* __restrict__ chirp_out) {
    __shared__ double red[2][8];
    const int b = blockIdx.x;
    if (b >= 16 * kS1) {
        for (int k = (b - 16 * kS1) * 256 + threadIdx.x; k < P; k += 256 * (gridDim.x - 16 * kS1)) {
            const long long m = k + lo;
            double s, c;
            sincospi((double)((m * m) % (2LL * N)) / (double)N, &s, &c);
            chirp_out[k] = make_float2((float)c, (float)s);
        }
        return;
    }
    const int u = b >> 4, e = b & 15, q = F::k_out(u, e);
    // Bhat[q] = (1/M) sum_d b'[d] e^{+2 pi i (d mod M) q / M},  b'[d] = conj(c[d + lo]),  -N < d < P
    double sr = 0.0, si = 0.0;
    for (int d = 1 - N + threadIdx.x; d <= P - 1; d += 256) {
        const int j = ((d % kM) + kM) % kM;
        const long long dl = d + lo;
        const long long d2 = (dl * dl) % (2LL * N);
        const double ph = -(double)d2 / (double)N + 2.0 * (double)(((long long)j * q) % kM) / (double)kM;
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
        bhatp[u * kRowB + e] = make_float2((float)(tr / kM), (float)(ti / kM));
        if (e < kRowB - 16) bhatp[u * kRowB + 16 + e] = make_float2(0.f, 0.f);
    }
}

// wcq[((item N + r) C + j) S1 + u] = (wc[r][n0], wc[r][n0 + S1]),  n0 = u + S1 2j,
// wc[r][n] = weight[r][n] c[n] c[r] (0 for n >= N): the input chirp of pass 1 (c[n]) and, because both passes
// are linear, the input chirp of pass 2 (c[r'], constant along a row) in one complex table.  Phase in float64.
__global__ void blue_weight_kernel(const float* __restrict__ w, float4* __restrict__ wcq, int N, int C,
                                   long long total) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int u = (int)(o % kS1), j = (int)((o / kS1) % C);
    const long long row = o / ((long long)kS1 * C);        // item N + r
    const long long r = row % N;
    float v[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long n = u + (long long)kS1 * (2 * j + h);
        float re = 0.f, im = 0.f;
        if (n < N) {
            double s, c;
            sincospi((double)((n * n + r * r) % (2LL * N)) / (double)N, &s, &c);
            const float x = w[row * N + n];
            re = x * (float)c;
            im = x * (float)s;
        }
        v[2 * h] = re;
        v[2 * h + 1] = im;
    }
    wcq[o] = make_float4(v[0], v[1], v[2], v[3]);
}

// conj(a) w
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 w) { return cmul(make_float2(a.x, -a.y), w); }

// C: cell pairs per thread (bluestein.cuh); TWO: the two transforms of a convolution as two inlined copies
// (the first pruned by the zero inputs, the second by the crop) instead of one copy executed twice
template <int RNG, bool SH, int C, bool TWO>
__global__ void __launch_bounds__(kBlueThreads, kBlueThreads <= 128 ? 4 : 2)
    screen_detect_bluestein(const __grid_constant__ RunArgs a, const float2* __restrict__ tables) {
    using Tw = typename F::Tw;
    constexpr int THREADS = kBlueThreads, M = kM, S1 = kS1, LPB = THREADS / S1;
    constexpr int NC = 2 * C;                                   // cells per thread
    static_assert(THREADS % S1 == 0 && LPB >= 1 && (S1 <= 32 || LPB <= 15), "line/barrier layout");
    static_assert(C >= 5 && C <= 8, "cell-pair class");
    constexpr bool kIdent = blue_identity_layout<F>();
    constexpr unsigned kKeep = blue_keep_mask_low<F>(blue_output_bound(C, S1));
    constexpr int kLinesPerWarp = S1 <= 32 ? 32 / S1 : 1;
    constexpr bool kShuffle = F::kShflC && S1 == 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.n, P = a.n_pup;
    Tw* twa = reinterpret_cast<Tw*>(smem_raw);
    Tw* twb = twa + F::kTwA;
    float2* bufs = reinterpret_cast<float2*>(twb + F::kTwB);
    float2* bhatp = bufs + LPB * F::kBuf;           // S1 rows of kRowB
    float2* chirp_out = bhatp + kRowB * S1;         // P (+ pad to even)
    double* st = reinterpret_cast<double*>(chirp_out + P + (P & 1));
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
    for (int j = tid; j < kRowB * S1; j += THREADS) bhatp[j] = tables[j];
    for (int j = tid; j < P; j += THREADS) chirp_out[j] = tables[kRowB * S1 + j];
    if (tid == 0) stats_reset(st, 0);
    __syncthreads();

    float2* T = a.scratch + (size_t)blockIdx.x * N * P;
    const int n1 = (N + LPB - 1) / LPB, n2 = (P + LPB - 1) / LPB;

    // wanted outputs of this thread: crop index k' = kb + k_off(e) < P
    const int kb = F::k_base(u);
    unsigned need = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e)
        if (((kKeep >> e) & 1u) && kb + F::k_off(e) < P) need |= 1u << e;
    // which of this thread's cells lie inside the grid (n = u + S1 m < N): bit m
    unsigned inside = 0;
#pragma unroll
    for (int m = 0; m < NC; ++m)
        if (u + S1 * m < N) inside |= 1u << m;
    const float4* brow = reinterpret_cast<const float4*>(bhatp + u * kRowB);

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const PairId id = pair_id(a, pair);
        const unsigned long long g = id.g;
        const uint32_t k0 = (uint32_t)id.seed, k1 = (uint32_t)(id.seed >> 32);
        const float4* wcq = reinterpret_cast<const float4*>(a.weight_s) + (size_t)id.item * N * C * S1 + u;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (SH) sh_prepare(a, pair, id, sh_amp, sh_tab);       // complete after the barrier between the passes

        // One loop body serves both passes (instruction-cache footprint): iterations [0, n1) are rows, [n1, n1 + n2)
        // kept columns.  No CTA-wide barrier inside a pass.  Lines beyond the grid / crop inside a live warp run on
        // zeros (the warp stays converged for the line syncs); warps without a live line skip the iteration.
        for (int it = 0; it < n1 + n2; ++it) {
            const bool rows = it < n1;
            if (it == n1) __syncthreads();                     // every row of T is stored before a column is read
            const int line = (rows ? it : it - n1) * LPB + ln; // r' or c
            const int nlines = rows ? N : P;
            if (line - (ln % kLinesPerWarp) >= nlines) continue;
            const bool live = line < nlines;
            float2 v[16];
#pragma unroll
            for (int m = NC; m < 16; ++m) v[m] = make_float2(0.f, 0.f);
            if (rows) {
                // the chirped weights first: their L2 latency hides behind the Philox rounds
                float4 wv[C];
                const float4* wq = wcq + (size_t)(live ? line : 0) * (C * S1);
#pragma unroll
                for (int j = 0; j < C; ++j) wv[j] = __ldg(wq + j * S1);
                if (RNG != kRngHost) {
                    uint32_t mr[16], ma[16];
                    if (RNG == kRngFast) noise_block_fields_fast((uint32_t)(line * S1 + u), g, k0, k1, mr, ma);
                    else noise_block_fields((uint32_t)(line * S1 + u), g, k0, k1, mr, ma);
#pragma unroll
                    for (int m = 0; m < NC; ++m) {
                        const float4 w4 = wv[m >> 1];
                        const float2 w = (m & 1) ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
                        v[m] = cmul(weighted_normal_m(mr[m], ma[m], 1.0f), w);
                    }
                } else {
                    const float2* nrow = a.noise + ((size_t)pair * N + (live ? line : 0)) * N + u;
#pragma unroll
                    for (int m = 0; m < NC; ++m) {
                        const float4 w4 = wv[m >> 1];
                        const float2 w = (m & 1) ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
                        const float2 nz = (inside >> m) & 1u ? __ldg(nrow + S1 * m) : make_float2(0.f, 0.f);
                        v[m] = cmul(nz, w);
                    }
                }
            } else {
                const float2* tcol = T + (size_t)(live ? line : 0) * N + u;
#pragma unroll
                for (int m = 0; m < NC; ++m) v[m] = (inside >> m) & 1u ? __ldcg(tcol + S1 * m) : make_float2(0.f, 0.f);
            }

            // ---- the convolution: v <- G(conj(G(v) Bhat)) --------------------------------------------------
            auto pointwise = [&]() {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w = brow[j];
                    v[2 * j] = cmul(v[2 * j], make_float2(w.x, w.y));
                    v[2 * j + 1] = cmul(v[2 * j + 1], make_float2(w.z, w.w));
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e].y = -v[e].y;
                if (!kIdent) {
                    // back to the input order of the transform through the line buffer (free after run())
#pragma unroll
                    for (int e = 0; e < 16; ++e) buf[F::k_base(u) + F::k_off(e)] = v[e];
                    sync();
#pragma unroll
                    for (int m = 0; m < 16; ++m) v[m] = buf[u + S1 * m];
                    sync();
                }
            };
            if constexpr (TWO && kShuffle) {
                // M = 512: the last radix-2 stage on warp shuffles, as in the radix kernel of that size
                F::template run_shfl<0xffffu>(u, v, twa, twb, buf, sync);
                pointwise();
                F::template run_shfl<kKeep>(u, v, twa, twb, buf, sync);
            } else if (TWO) {
                F::run(u, v, twa, twb, buf, sync);
                pointwise();
                F::run(u, v, twa, twb, buf, sync);
            } else {
#pragma unroll 1
                for (int rep = 0; rep < 2; ++rep) {
                    F::run(u, v, twa, twb, buf, sync);
                    if (rep == 0) pointwise();
                }
            }

            if (!live) continue;
            if (rows) {
                float2* tb = T + ((long long)kb * N + line);
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (((kKeep >> e) & 1u) && (need & (1u << e)))
                        __stcg(tb + (long long)F::k_off(e) * N, cmul_conj(v[e], chirp_out[kb + F::k_off(e)]));
            } else {
                const float* ub = a.u_t + ((long long)line * P + kb);
                // output sign (-1)^(row + column) = (-1)^(k' + lo + c + lo) = (-1)^(k' + c); k_off is even
                const float sgn = ((kb + line) & 1) ? -1.f : 1.f;
                float2 ex[3];
                if (SH) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) ex[i] = __ldg(a.sh_ex + i * P + line);
                }
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (((kKeep >> e) & 1u) && (need & (1u << e))) {
                        const float uu = __ldg(ub + F::k_off(e));
                        const float2 phi = cmul_conj(v[e], chirp_out[kb + F::k_off(e)]);
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

size_t blue_smem_bytes(bool sh, int n_pup) {
    constexpr int LPB = kBlueThreads / kS1;
    return sizeof(float2) * ((size_t)F::kTwA + F::kTwB + (size_t)LPB * F::kBuf + (size_t)kRowB * kS1 + n_pup + (n_pup & 1)) +
           sizeof(double) * kStatWords + sizeof(float) * 4 * (kBlueThreads / 32) + sh_smem_bytes(sh, n_pup);
}

#ifndef FASTB_BLUE_TWO
#define FASTB_BLUE_TWO 1
#endif
constexpr bool kTwo = FASTB_BLUE_TWO != 0;

template <int RNG>
void (*pick(bool sh, int C))(RunArgs, const float2*) {
    // the sub-harmonic instances are not specialised beyond the cell count
    if (sh) {
        switch (C) {
            case 5: return screen_detect_bluestein<RNG, true, 5, kTwo>;
            case 6: return screen_detect_bluestein<RNG, true, 6, kTwo>;
            case 7: return screen_detect_bluestein<RNG, true, 7, kTwo>;
            default: return screen_detect_bluestein<RNG, true, 8, kTwo>;
        }
    }
    switch (C) {
        case 5: return screen_detect_bluestein<RNG, false, 5, kTwo>;
        case 6: return screen_detect_bluestein<RNG, false, 6, kTwo>;
        case 7: return screen_detect_bluestein<RNG, false, 7, kTwo>;
        default: return screen_detect_bluestein<RNG, false, 8, kTwo>;
    }
}

}  // namespace

int FASTB_CAT(prepare_blue_, FASTB_LOG2M)(int n, int n_pup, int lo, int C, int n_items, const float* weight, void* wcq,
                                          void* tables, cudaStream_t st) {
    float2* t = (float2*)tables;
    blue_tables_kernel<<<16 * kS1 + 4, 256, 0, st>>>(n, lo, n_pup, t, t + kRowB * kS1);
    int rc = check_launch("blue_tables_kernel");
    if (rc || !weight) return rc;
    const long long total = (long long)n_items * n * C * kS1;
    blue_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(weight, (float4*)wcq, n, C, total);
    return check_launch("blue_weight_kernel");
}

int FASTB_CAT(launch_blue_, FASTB_LOG2M)(const RunArgs& a, const RadixRequest& rq, const void* tables, int C,
                                         cudaStream_t st) {
    const bool sh = a.sh_weight != nullptr;
    void (*kern)(RunArgs, const float2*) = rq.rng == kRngHost   ? pick<kRngHost>(sh, C)
                                           : rq.rng == kRngFast ? pick<kRngFast>(sh, C)
                                                                : pick<kRngPhilox>(sh, C);
    const size_t smem = blue_smem_bytes(sh, a.n_pup);
    if (smem > 227 * 1024) {
        set_error("screen_detect_bluestein: N=%d needs %zu B of shared memory", a.n, smem);
        return FASTB_ERR_UNSUPPORTED;
    }
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0, sms = 0;
    FASTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlueThreads, smem));
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
    kern<<<(unsigned)grid, kBlueThreads, smem, st>>>(a, (const float2*)tables);
    return check_launch("screen_detect_bluestein");
}

}  // namespace fastb
