// K2: phase-screen synthesis fused with the fibre-overlap detector.  Contract: include/fastb.h.
//
// One persistent CTA owns one complex transform ("pair" = two realisations) at a time:
//   pass 1  for every frequency row r': white noise (Philox + Box-Muller in registers, or the
//           caller's noise) x weight -> N-point line FFT in registers -> keep the n_pup output
//           columns of the pupil crop -> CTA-private scratch T[c][r'] (L2 resident)
//   pass 2  for every kept column c: N-point line FFT over r' -> keep the n_pup rows of the crop
//           -> U (cos phi, sin phi) accumulated in registers for Re and Im screens
//   final   fixed-order block reduction, exp(chi), normalisation -> 1 scalar per realisation.
// Signs: the weight carries (-1)^(r'+c') and the output (-1)^(r+c), which turns the reference's
// centred (fftshift-ed) inverse DFT (fast/funcs.py:218 via aotools.ift2) into a plain one.
#include "fastb_common.cuh"
#include "fft_core.cuh"

#include <stdlib.h>

namespace fastb {
namespace {

constexpr int kThreads = 256;

struct RunArgs {
    int n, n_pup, lo, coherent;
    long long n_pairs, first_pair, ppc;
    unsigned long long seed;
    float inv_usum, sigma_chi;
    const float* weight;      // N*N signed weight
    float* weight_s;          // radix kernel, device RNG: weight * sqrt(2 ln 2), interleaved per thread
                              // (workspace; written by scale_weight_kernel): element (row r, thread u,
                              // register m) at r*N + (m/4)*(4*S1) + 4*u + m%4
    const float* u_t;         // n_pup*n_pup, transposed: u_t[c*n_pup + r]
    const float2* u_p;        // line-pair kernel: u_p[cp*n_pup + r] = (U[r][2cp], U[r][2cp+1] or 0)
    const float* chi;         // global-index log-amplitudes or NULL
    const float2* noise;      // n_pairs*N*N or NULL
    float* out_a;
    float* out_b;
    float2* scratch;          // gridDim.x slots of n*n_pup float2
    int rows_per_block;       // direct kernel only
    int stage_shift;          // radix kernel: 1 = stage two rows per line slot before storing, 0 = store directly
    // sub-harmonics (NULL weight = off)
    const float* sh_weight;   // 27
    const float2* sh_noise;   // n_pairs*27 or NULL
    const float2* sh_ex;      // 3*n_pup
    const float2* sh_ey;      // 3*n_pup
    const float2* sh_mean;    // 27
    float* phs;               // direct kernel only: write the cropped screens instead of detecting
#ifdef FASTB_TUNE_DBG
    int dbg;                  // timing experiments (results wrong): 1 no scratch stores, 2 no weight
                              // loads, 4 no scratch loads, 8 no detector loads, 16 no noise, 32 no
                              // MUFU in Box-Muller, 64 cheap hash instead of Philox, 128 no detector sincos
#endif
};
#ifdef FASTB_TUNE_DBG
#define FASTB_DBG(a, bit) ((a).dbg & (bit))
#else
#define FASTB_DBG(a, bit) 0
#endif

// ---- sub-harmonic term (include/fastb.h FastbSubharm) ------------------------------------
// Per pair: 27 amplitudes -> per pupil row a table of 7 complex numbers
//   tab[r] = { B, A_0[-], A_0[+], A_1[-], A_1[+], A_2[-], A_2[+] },
//   A_i[s](r) = sum_q amp_i[q][s] Ey_i[q](r),  B = sum_i A_i[0](r) - grid mean,
// so that a pixel costs 6 complex MACs: phi_sh = B + sum_i (A_i[-] conj(Ex_i) + A_i[+] Ex_i).
constexpr int kShTab = 7;

__device__ __forceinline__ float2 cmac(float2 acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x));
    acc.y = fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y));
    return acc;
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// amp: 28 float2 in shared memory (27 amplitudes + the mean), tab: n_pup*7 float2.
// Ends with the table complete only after the caller's next __syncthreads().
__device__ void sh_prepare(const RunArgs& a, long long pair, float2* amp, float2* tab) {
    const int tid = threadIdx.x;
    const unsigned long long g = (unsigned long long)(a.first_pair + pair);
    if (tid < 14) {
        float2 n0, n1;
        if (a.sh_noise) {
            n0 = a.sh_noise[pair * 27 + 2 * tid];
            n1 = (2 * tid + 1 < 27) ? a.sh_noise[pair * 27 + 2 * tid + 1] : make_float2(0.f, 0.f);
        } else {
            const uint4 w = philox4x32_10((uint32_t)tid, (uint32_t)g, (uint32_t)(g >> 32), kStreamSubharm,
                                          (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            n0 = box_muller(w.x, w.y);
            n1 = box_muller(w.z, w.w);
        }
        const float w0 = a.sh_weight[2 * tid];
        amp[2 * tid] = make_float2(n0.x * w0, n0.y * w0);
        if (2 * tid + 1 < 27) {
            const float w1 = a.sh_weight[2 * tid + 1];
            amp[2 * tid + 1] = make_float2(n1.x * w1, n1.y * w1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        float2 m = make_float2(0.f, 0.f);
        for (int k = 0; k < 27; ++k) m = cmac(m, amp[k], a.sh_mean[k]);
        amp[27] = m;
    }
    __syncthreads();
    const int P = a.n_pup;
    for (int r = tid; r < P; r += blockDim.x) {
        float2 B = make_float2(-amp[27].x, -amp[27].y);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float2 ey = a.sh_ey[i * P + r], eyc = cconj(ey);
            const float2* ai = amp + i * 9;           // [q][s]
#pragma unroll
            for (int sx = 0; sx < 3; ++sx) {
                float2 acc = ai[3 + sx];              // q = 1: fy = 0
                acc = cmac(acc, ai[sx], eyc);         // q = 0: fy = -df
                acc = cmac(acc, ai[6 + sx], ey);      // q = 2: fy = +df
                if (sx == 1) {
                    B.x += acc.x;
                    B.y += acc.y;
                } else {
                    tab[r * kShTab + 1 + 2 * i + (sx == 2)] = acc;
                }
            }
        }
        tab[r * kShTab] = B;
    }
}

__device__ __forceinline__ float2 sh_phase(const float2* tabrow, const float2 (&ex)[3]) {
    float2 p = tabrow[0];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        p = cmac(p, tabrow[1 + 2 * i], cconj(ex[i]));
        p = cmac(p, tabrow[2 + 2 * i], ex[i]);
    }
    return p;
}

// accumulate U exp(i s phi) for the two screens carried by one complex sample; us = s * u with
// s = +-1 the output sign of the centred transform (cos is even, so only the sine terms see it).
// sin.approx / cos.approx reduce the argument internally (x / 2pi in fp32): for |phi| < ~30 rad
// the phase error stays ~1e-6 rad, far below the 1e-4 parity budget on the power.
__device__ __forceinline__ void accumulate(float2 phi, float u, float us, float (&acc)[4]) {
    float s, c;
    __sincosf(phi.x, &s, &c);
    acc[0] = fmaf(u, c, acc[0]);
    acc[1] = fmaf(us, s, acc[1]);
    __sincosf(phi.y, &s, &c);
    acc[2] = fmaf(u, c, acc[2]);
    acc[3] = fmaf(us, s, acc[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fixed-order reduction of the 4 accumulators over the CTA, then the per-pair epilogue
template <int THREADS = kThreads>
__device__ void finish_pair(const RunArgs& a, long long pair, float (&acc)[4], float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) red[warp * 4 + i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < THREADS / 32; ++w)
            for (int i = 0; i < 4; ++i) t[i] += red[w * 4 + i];
        const long long g = a.first_pair + pair;
        const long long chunk = g / a.ppc, pp = g % a.ppc;
        const long long ia = chunk * 2 * a.ppc + pp, ib = ia + a.ppc;
        float chia, chib;
        if (a.chi) {
            chia = a.chi[ia];
            chib = a.chi[ib];
        } else {
            chia = a.sigma_chi * chi_normal(a.seed, (uint64_t)ia);
            chib = a.sigma_chi * chi_normal(a.seed, (uint64_t)ib);
        }
        const float ea = expf(chia) * a.inv_usum, eb = expf(chib) * a.inv_usum;
        const float zar = ea * t[0], zai = ea * t[1], zbr = eb * t[2], zbi = eb * t[3];
        if (a.coherent) {
            a.out_a[2 * pair] = zar;
            a.out_a[2 * pair + 1] = zai;
            a.out_b[2 * pair] = zbr;
            a.out_b[2 * pair + 1] = zbi;
        } else {
            a.out_a[pair] = zar * zar + zai * zai;
            a.out_b[pair] = zbr * zbr + zbi * zbi;
        }
    }
    __syncthreads();
}

// ---- TMA bulk copy (cp.async.bulk, 1-D) + mbarrier helpers -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> this CTA's shared memory; completion (byte count) is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Synchronise the S1 threads that share one line.  S1 <= 32: the line lives inside a warp.
// S1 = 64 / 128: a named barrier per line (ids 1..LPB; 0 is __syncthreads), so lines do not
// wait for each other.
template <int S1>
struct LineSync {
    int ln;
    __device__ __forceinline__ void operator()() const {
        if (S1 > 32) asm volatile("bar.sync %0, %1;" ::"r"(ln + 1), "n"(S1) : "memory");
        else __syncwarp();
    }
};

// Line-FFT flavour and CTA shape per grid size, measured on B200 (profiles/)
//   N <= 256 : 128 threads x 4 CTAs/SM (128 registers) -- small CTAs keep the three per-pair
//              barriers cheap and balance the pupil columns over 8 lines per iteration
//   N  = 512 : 256 threads x 2 CTAs/SM (125 registers); N >= 1024: 256 x 3 (80 registers)
// All use the 16-elements-per-thread FFT.  A 32-elements-per-thread flavour (512 = 32 x 16,
// 1024 = 32 x 32: one exchange per line, 168 registers) is kept for tuning builds
// (FASTB_TUNE, FASTB_E=32): it measured the same throughput in every CTA shape -- the kernel is
// bound by register-file-limited ILP x TLP, not by the number of shared-memory exchanges.
template <int LOG2N, int E>
struct RadixCfg;
template <int LOG2N>
struct RadixCfg<LOG2N, 16> {
    using F = LineFFT<LOG2N>;
    static constexpr int kThreadsPerCta = LOG2N <= 8 ? 128 : 256;
    static constexpr int kMinBlocks = LOG2N <= 8 ? 4 : (LOG2N == 9 ? 2 : 3);
};
template <int LOG2N>
struct RadixCfg<LOG2N, 32> {
    using F = LineFFT32<LOG2N>;
    static constexpr int kThreadsPerCta = 128;
    static constexpr int kMinBlocks = 3;
};
template <int LOG2N>
constexpr int radix_default_e() { return 16; }

// One loop body serves both passes (keeps the kernel inside the instruction cache): iterations
// [0, n1) are frequency rows (noise -> FFT -> pruned store to T[c][r']), iterations [n1, n1+n2)
// are kept columns (load T[c][:] -> FFT -> detector accumulation).  No CTA-wide barrier inside
// a pass.  TMA != 0 (tuning builds, E = 16, lines inside a warp): the warp's contiguous input
// block of the next iteration is fetched by one cp.async.bulk into a per-warp stage
// (1: weights and scratch, 2: scratch only, 3: weights only) -- measured slower, off by default.
// WIN (compile time) promises that the crop lies inside the centred window of half-width
// window_half<N>(WIN) (0: no promise): only the registers keep_mask<F>() names can then hold a kept
// output, and the compiler drops the last-stage butterflies (and shared loads) that feed the others.
template <int N>
constexpr int window_half(int win) { return win == 1 ? N / 8 : win == 2 ? 3 * N / 16 : win == 3 ? N / 4 : N; }

template <class F, bool RNG, bool SH, int THREADS, int MINB, int TMA = 0, int WIN = 0>
__global__ void __launch_bounds__(THREADS, MINB) screen_detect_radix(const __grid_constant__ RunArgs a) {
    constexpr int N = F::N, S1 = F::S1, E = F::E, LPB = THREADS / S1;
    static_assert(THREADS % S1 == 0 && LPB >= 1 && (S1 <= 32 || LPB <= 15), "line/barrier layout");
    static_assert(E == 16 || E == 32, "elements per thread");
    constexpr unsigned kKeep = WIN == 0 ? 0xffffffffu : keep_mask<F>(window_half<N>(WIN));
    constexpr bool kTma = (S1 <= 32) && TMA != 0 && E == 16;
    constexpr bool kTmaW = kTma && (TMA == 1 || TMA == 3);      // weight rows through the stage
    constexpr bool kTmaT = kTma && (TMA == 1 || TMA == 2);      // scratch columns through the stage
    constexpr int kLinesPerWarp = S1 <= 32 ? 32 / S1 : 1;
    constexpr int kWarps = THREADS / 32;
    constexpr int kStageBytes = 32 * E * 8;

    using Tw = typename F::Tw;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tw* twa = reinterpret_cast<Tw*>(smem_raw);
    Tw* twb = twa + F::kTwA;
    float2* bufs = reinterpret_cast<float2*>(twb + F::kTwB);
    // pass-1 output staging: per line slot R = 2^stage_shift (1 or 2) planes of n_pup kept outputs
    const int rs = kTma ? 0 : a.stage_shift, R = 1 << rs;
    float2* tiles = bufs + LPB * F::kBuf;
    unsigned char* stage_all = reinterpret_cast<unsigned char*>(tiles + (rs ? LPB * R * a.n_pup : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + (kTma ? kWarps * kStageBytes : 0));
    float* red = reinterpret_cast<float*>(bars + (kTma ? kWarps : 0));
    float2* sh_amp = reinterpret_cast<float2*>(red + 4 * kWarps);     // SH only
    float2* sh_tab = sh_amp + 28;

    const int tid = threadIdx.x;
    const int ln = tid / S1, u = tid % S1;
    const int warp = tid >> 5, lane = tid & 31;
    float2* buf = bufs + ln * F::kBuf;
    const int P = a.n_pup, lo = a.lo;
    float2* tile = tiles + ln * R * P;
    unsigned char* stage = stage_all + warp * kStageBytes;
    uint64_t* bar = bars + warp;
    uint32_t parity = 0;
    if (kTma && lane == 0) mbar_init(bar, 1);
    const LineSync<S1> sync{ln};

    for (int j = tid; j < F::kTwA + F::kTwB; j += THREADS) {
        const int ex = j < F::kTwA ? F::twa_exponent(j) : F::twb_exponent(j - F::kTwA);
        double s, c;
        sincospi(2.0 * (double)ex / (double)N, &s, &c);
        twa[j] = make_tw((float)c, (float)s, (Tw*)nullptr);
    }
    if (kTma) fence_proxy_async();            // mbarrier init visible to the async proxy
    __syncthreads();

    float2* T = a.scratch + (size_t)blockIdx.x * N * P;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const int n1 = N / LPB, n2 = (P + LPB - 1) / LPB;

    static_assert(F::k_off_all_even(), "the output sign is taken per thread: k_off must be even");
    // which of this thread's E outputs fall inside the crop [lo, lo+P): the same for every line
    // of both passes, so it is computed once and tested bit by bit
    const int kb = F::k_base(u) - lo;
    unsigned need = 0;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if ((unsigned)(kb + F::k_off(e)) < (unsigned)P) need |= 1u << e;

    // Enqueue this warp's input block of iteration `itn` (all lanes call it after a __syncwarp;
    // lane 0 issues).  Nothing is issued -- and nothing will be waited for -- when the warp has
    // no line inside the crop in that iteration.
    auto prefetch = [&](int itn) {
        const bool rows_n = itn < n1;
        if ((rows_n && !kTmaW) || (!rows_n && !kTmaT)) return;
        const int line0 = (rows_n ? itn : itn - n1) * LPB + warp * kLinesPerWarp;
        const int nlines = rows_n ? kLinesPerWarp : min(kLinesPerWarp, P - line0);
        if (nlines <= 0 || lane != 0) return;
        const uint32_t bytes = (uint32_t)nlines * N * (rows_n ? 4u : 8u);
        const void* src = rows_n ? (const void*)(a.weight + (size_t)line0 * N) : (const void*)(T + (size_t)line0 * N);
        fence_proxy_async();                  // earlier generic reads of the stage precede the async write
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stage, src, bytes, bar);
    };

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const unsigned long long g = (unsigned long long)(a.first_pair + pair);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (SH) sh_prepare(a, pair, sh_amp, sh_tab);   // table visible after the barrier at it == n1
        if (kTma) prefetch(0);
        for (int it = 0; it < n1 + n2; ++it) {
            const bool rows = it < n1;
            if (it == n1) {
                if (kTmaT) fence_proxy_async();       // T was written through the generic proxy
                __syncthreads();                      // every row of T is stored before a column is read
                if (kTma) prefetch(n1);
            }
            // pass 1: a line slot takes R (1 or 2) consecutive rows in R consecutive iterations, so
            // that its kept outputs leave as 16-byte stores of two adjacent rows per column
            const int sub = it & (R - 1);
            const int line = rows ? (((it >> rs) * LPB + ln) << rs) + sub       // r'
                                  : (it - n1) * LPB + ln;                        // c
            // last column iteration: warps whose lines all lie beyond the crop have nothing to do
            // (line barriers involve only the threads of that line)
            if (!rows && line - (ln % kLinesPerWarp) >= P) continue;

            float2 v[E];
            const bool staged = (rows && kTmaW) || (!rows && kTmaT);
            if (staged) {
                mbar_wait(bar, parity);
                parity ^= 1;
            }
            if (rows) {
                float w[E];
                if (kTmaW) {
                    const float* ws = reinterpret_cast<const float*>(stage) + (ln % kLinesPerWarp) * N;
#pragma unroll
                    for (int m = 0; m < E; ++m) w[m] = ws[u + S1 * m];
                    __syncwarp();
                    if (it + 1 < n1) prefetch(it + 1);
                } else if (RNG) {
                    const float4* wq = reinterpret_cast<const float4*>(a.weight_s + (size_t)line * N) + u;
#pragma unroll
                    for (int j = 0; j < E / 4; ++j) {
                        const float4 t = __ldg(wq + j * S1);
                        w[4 * j] = t.x;
                        w[4 * j + 1] = t.y;
                        w[4 * j + 2] = t.z;
                        w[4 * j + 3] = t.w;
                    }
                } else {
                    const float* wrow = a.weight + (size_t)line * N;
#pragma unroll
                    for (int m = 0; m < E; ++m) w[m] = FASTB_DBG(a, 2) ? 1.f + m : __ldg(wrow + u + S1 * m);
                }
                if (RNG && FASTB_DBG(a, 16)) {
#pragma unroll
                    for (int m = 0; m < E; ++m) v[m] = make_float2(w[m], w[m] * u);
                } else if (RNG) {
                    // thread (line, u) owns noise blocks t' = u + S1 h of its row: cell j of block h
                    // is element m = (E/16) j + h (include/fastb.h)
#pragma unroll
                    for (int h = 0; h < E / 16; ++h) {
                        uint32_t mr[16], ma[16];
                        if (FASTB_DBG(a, 64)) {                 // timing only: a cheap hash instead of Philox
                            uint32_t x = (uint32_t)(line * (N / 16) + u + S1 * h) * 0x9E3779B9u + (uint32_t)g;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                x = x * 1664525u + 1013904223u;
                                mr[j] = x >> 9;
                                ma[j] = (x * 0x85EBCA6Bu) >> 9;
                            }
                        } else {
                            noise_block_fields((uint32_t)(line * (N / 16) + u + S1 * h), g, k0, k1, mr, ma);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int m = (E / 16) * j + h;
                            if (FASTB_DBG(a, 32)) {             // timing only: no MUFU in Box-Muller
                                const float fr = __uint_as_float(0x3f800000u | mr[j]) - 1.5f;
                                const float fa = __uint_as_float(0x3f800000u | ma[j]) - 1.5f;
                                v[m] = make_float2(fr * w[m], fa * w[m]);
                            } else if (kTmaW) {
                                v[m] = weighted_normal_m(mr[j], ma[j], w[m]);     // staged from the caller's table
                            } else {
                                v[m] = weighted_normal_s(mr[j], ma[j], w[m]);     // pre-scaled copy
                            }
                        }
                    }
                } else {
                    const float2* nrow = a.noise + ((size_t)pair * N + line) * N;
#pragma unroll
                    for (int m = 0; m < E; ++m) {
                        const float2 nz = __ldg(nrow + u + S1 * m);
                        v[m] = make_float2(nz.x * w[m], nz.y * w[m]);
                    }
                }
            } else if (kTmaT) {
                const float2* ts = reinterpret_cast<const float2*>(stage) + (ln % kLinesPerWarp) * N;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = ts[u + S1 * m];
                __syncwarp();
                if (it + 1 < n1 + n2) prefetch(it + 1);
            } else {
                const float2* tcol = T + (size_t)(line < P ? line : 0) * N;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = FASTB_DBG(a, 4) ? make_float2(m, u) : __ldcg(tcol + u + S1 * m);
            }

            F::run(u, v, twa, twb, buf, sync);

            if (rows && rs == 0) {
                float2* tb = T + ((long long)kb * N + line);  // &T[(k - lo) * N + r'] at k_off = 0
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (((kKeep >> e) & 1u) && (need & (1u << e)) && !FASTB_DBG(a, 1))
                        __stcg(tb + (long long)F::k_off(e) * N, v[e]);
            } else if (rows) {
                float2* tl = tile + (sub * P + kb);           // &tile[sub][k - lo] at k_off = 0
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (((kKeep >> e) & 1u) && (need & (1u << e))) tl[F::k_off(e)] = v[e];
                if (sub == R - 1) {
                    // flush the slot: column c gets rows line-1, line as one 16-byte store.  The
                    // tile is next written after the line barriers of the following iteration's
                    // FFT, so no barrier is needed after the reads.
                    sync();
                    float2* tr = T + (line - (R - 1));
#pragma unroll 2
                    for (int c = u; c < P; c += S1) {
                        const float2 x0 = tile[c], x1 = tile[P + c];
                        __stcg(reinterpret_cast<float4*>(tr + (long long)c * N), make_float4(x0.x, x0.y, x1.x, x1.y));
                    }
                }
            } else if (line < P) {
                const float* ub = a.u_t + ((long long)line * P + kb);
                // output sign (-1)^(r + c): k_off is even, so it is one value per thread and line
                const float sgn = ((F::k_base(u) + line + lo) & 1) ? -1.f : 1.f;
                float2 ex[3];
                if (SH) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) ex[i] = __ldg(a.sh_ex + i * P + line);
                }
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if (((kKeep >> e) & 1u) && (need & (1u << e))) {
                        const float uu = FASTB_DBG(a, 8) ? 1.f : __ldg(ub + F::k_off(e));
                        if (SH) {
                            const float2 sp = sh_phase(sh_tab + (kb + F::k_off(e)) * kShTab, ex);
                            accumulate(make_float2(fmaf(sgn, v[e].x, sp.x), fmaf(sgn, v[e].y, sp.y)), uu, uu, acc);
                        } else {
                            accumulate(v[e], uu, uu * sgn, acc);
                        }
                    }
                }
            }
        }
        finish_pair<THREADS>(a, pair, acc, red);
    }
}

// ---- line-PAIR kernel: two adjacent lines per thread group, planar packed FP32 -----------------
// Same algorithm and results as screen_detect_radix, but every thread carries the same position
// of TWO adjacent lines (rows 2p, 2p+1 in pass 1; pupil columns 2q, 2q+1 in pass 2) as planar
// pairs (fft_core.cuh, value type pc), so that all FFT arithmetic and most of Box-Muller are
// packed FP32 (FADD2 / FMUL2 / FFMA2).  Scratch layout: T4[q][r'] = float4 (re(2q), re(2q+1),
// im(2q), im(2q+1)), which pass 2 reads with one 128-bit load per element pair.
__device__ __forceinline__ pc weighted_normal_pair(uint32_t mrA, uint32_t maA, uint32_t mrB, uint32_t maB, float2 w) {
    const float2 u1 = sub2(bc2(2.0f), make_float2(__uint_as_float(0x3f800000u | mrA), __uint_as_float(0x3f800000u | mrB)));
    const float2 r2 = mul2(make_float2(lg2_ftz(u1.x), lg2_ftz(u1.y)), bc2(-1.3862943611198906f));
    float2 rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.x) : "f"(r2.x));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.y) : "f"(r2.y));
    rad = mul2(rad, w);
    const float2 ang = mul2(bc2(6.283185307179586f),
                            make_float2(__uint_as_float(0x3f800000u | maA), __uint_as_float(0x3f800000u | maB)));
    float2 sn, cs;
    __sincosf(ang.x, &sn.x, &cs.x);
    __sincosf(ang.y, &sn.y, &cs.y);
    return pc{mul2(rad, cs), mul2(rad, sn)};
}

// U exp(i s phi) for the two columns (A, B) of a pair and both screens; us = (s_A u_A, s_B u_B)
__device__ __forceinline__ void accumulate_pair(pc phi, float2 u, float2 us, float (&acc)[4]) {
    float s, c;
    __sincosf(phi.re.x, &s, &c);
    acc[0] = fmaf(u.x, c, acc[0]);
    acc[1] = fmaf(us.x, s, acc[1]);
    __sincosf(phi.re.y, &s, &c);
    acc[0] = fmaf(u.y, c, acc[0]);
    acc[1] = fmaf(us.y, s, acc[1]);
    __sincosf(phi.im.x, &s, &c);
    acc[2] = fmaf(u.x, c, acc[2]);
    acc[3] = fmaf(us.x, s, acc[3]);
    __sincosf(phi.im.y, &s, &c);
    acc[2] = fmaf(u.y, c, acc[2]);
    acc[3] = fmaf(us.y, s, acc[3]);
}

template <int LOG2N, bool RNG, bool SH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) screen_detect_pair(const __grid_constant__ RunArgs a) {
    using F = LineFFT<LOG2N, pc>;
    using Tw = typename F::Tw;
    constexpr int N = F::N, S1 = F::S1, LPB = THREADS / S1;       // LPB line PAIRS per iteration
    static_assert(THREADS % S1 == 0 && LPB >= 1 && (S1 <= 32 || LPB <= 15), "line/barrier layout");
    static_assert((N / 2) % LPB == 0, "row pairs per iteration");
    constexpr int kPairsPerWarp = S1 <= 32 ? 32 / S1 : 1;
    constexpr int kWarps = THREADS / 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tw* twa = reinterpret_cast<Tw*>(smem_raw);
    Tw* twb = twa + F::kTwA;
    float2* bufs = reinterpret_cast<float2*>(twb + F::kTwB);
    float* red = reinterpret_cast<float*>(bufs + LPB * F::kBuf);
    float2* sh_amp = reinterpret_cast<float2*>(red + 4 * kWarps);     // SH only
    float2* sh_tab = sh_amp + 28;

    const int tid = threadIdx.x;
    const int lp = tid / S1, u = tid % S1;
    float2* buf = bufs + lp * F::kBuf;
    const int P = a.n_pup, lo = a.lo, PP = (P + 1) >> 1;             // PP pupil-column pairs
    const LineSync<S1> sync{lp};

    for (int j = tid; j < F::kTwA + F::kTwB; j += THREADS) {
        const int ex = j < F::kTwA ? F::twa_exponent(j) : F::twb_exponent(j - F::kTwA);
        double s, c;
        sincospi(2.0 * (double)ex / (double)N, &s, &c);
        twa[j] = make_tw((float)c, (float)s, (Tw*)nullptr);
    }
    __syncthreads();

    float4* T4 = reinterpret_cast<float4*>(a.scratch) + (size_t)blockIdx.x * N * PP;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const int n1 = (N / 2) / LPB, n2 = (PP + LPB - 1) / LPB;

    static_assert(F::k_off_all_even(), "column parity / output sign are taken per thread: k_off must be even");
    const int kb = F::k_base(u) - lo;          // crop index of this thread's output at k_off = 0
    unsigned need = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e)
        if ((unsigned)(kb + F::k_off(e)) < (unsigned)P) need |= 1u << e;
    // output sign (-1)^(r + c) for the even column of a pair; the odd column has the opposite one
    const float sgn_a = ((F::k_base(u) + lo) & 1) ? -1.f : 1.f;
    const float2 sgn = make_float2(sgn_a, -sgn_a);

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const unsigned long long g = (unsigned long long)(a.first_pair + pair);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (SH) sh_prepare(a, pair, sh_amp, sh_tab);   // table visible after the barrier at it == n1
        for (int it = 0; it < n1 + n2; ++it) {
            const bool rows = it < n1;
            if (it == n1) __syncthreads();            // every row of T is stored before a column is read
            const int pl = (rows ? it : it - n1) * LPB + lp;        // row-pair (pass 1) / column-pair (pass 2)
            if (!rows && pl - (lp % kPairsPerWarp) >= PP) continue; // warp has no column pair inside the crop

            pc v[16];
            if (rows) {
                const int ra = 2 * pl;
                const float* wa = a.weight + (size_t)ra * N;
                if (RNG) {
                    uint32_t mra[16], maa[16], mrb[16], mab[16];
                    noise_block_fields((uint32_t)(ra * S1 + u), g, k0, k1, mra, maa);
                    noise_block_fields((uint32_t)((ra + 1) * S1 + u), g, k0, k1, mrb, mab);
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int j = u + S1 * m;
                        v[m] = weighted_normal_pair(mra[m], maa[m], mrb[m], mab[m],
                                                    make_float2(__ldg(wa + j), __ldg(wa + N + j)));
                    }
                } else {
                    const float2* na = a.noise + ((size_t)pair * N + ra) * N;
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int j = u + S1 * m;
                        const float2 za = __ldg(na + j), zb = __ldg(na + N + j);
                        const float2 w = make_float2(__ldg(wa + j), __ldg(wa + N + j));
                        v[m] = pc{mul2(make_float2(za.x, zb.x), w), mul2(make_float2(za.y, zb.y), w)};
                    }
                }
            } else {
                const float4* tcol = T4 + (size_t)(pl < PP ? pl : 0) * N;
                const bool valid_b = 2 * pl + 1 < P;      // odd P: the last pair has no second column
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const float4 q = __ldcg(tcol + u + S1 * m);
                    v[m] = pc{make_float2(q.x, valid_b ? q.y : 0.f), make_float2(q.z, valid_b ? q.w : 0.f)};
                }
            }

            F::run(u, v, twa, twb, buf, sync);

            if (rows) {
                // output k -> crop column c = kb + k_off(e); its pair is c >> 1 and its slot c & 1
                // (= kb & 1: k_off is even).  Rows 2 pl and 2 pl + 1 are consecutive float4 of T4.
                float* tb = reinterpret_cast<float*>(T4 + ((long long)(kb >> 1) * N + 2 * pl)) + (kb & 1);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    if (need & (1u << e)) {
                        float* q = tb + (long long)(F::k_off(e) / 2) * N * 4;
                        __stcg(q, v[e].re.x);
                        __stcg(q + 2, v[e].im.x);
                        __stcg(q + 4, v[e].re.y);
                        __stcg(q + 6, v[e].im.y);
                    }
                }
            } else if (pl < PP) {
                const float2* ub = a.u_p + ((long long)pl * P + kb);
                float2 exa[3], exb[3];
                if (SH) {
                    const int cb = min(2 * pl + 1, P - 1);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        exa[i] = __ldg(a.sh_ex + i * P + 2 * pl);
                        exb[i] = __ldg(a.sh_ex + i * P + cb);
                    }
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    if (need & (1u << e)) {
                        const float2 uu = __ldg(ub + F::k_off(e));
                        if (SH) {
                            const float2* tabrow = sh_tab + (kb + F::k_off(e)) * kShTab;
                            const float2 spa = sh_phase(tabrow, exa), spb = sh_phase(tabrow, exb);
                            pc ph;
                            ph.re = fma2(sgn, v[e].re, make_float2(spa.x, spb.x));
                            ph.im = fma2(sgn, v[e].im, make_float2(spa.y, spb.y));
                            accumulate_pair(ph, uu, uu, acc);
                        } else {
                            accumulate_pair(v[e], uu, mul2(uu, sgn), acc);
                        }
                    }
                }
            }
        }
        finish_pair<THREADS>(a, pair, acc, red);
    }
}

// ---- general even N: pruned direct DFT (slow path; also used for N not a power of two) ----
template <bool RNG, bool SH>
__global__ void __launch_bounds__(kThreads) screen_detect_direct(const __grid_constant__ RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.n, P = a.n_pup, lo = a.lo, R = a.rows_per_block;
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float2* rows = tw + N;                       // R x N coloured noise
    float* red = reinterpret_cast<float*>(rows + (size_t)R * N);
    float2* sh_amp = reinterpret_cast<float2*>(red + 4 * (kThreads / 32));     // SH only
    float2* sh_tab = sh_amp + 28;
    const int tid = threadIdx.x;

    for (int j = tid; j < N; j += kThreads) {
        double s, c;
        sincospi(2.0 * (double)j / (double)N, &s, &c);
        tw[j] = make_float2((float)c, (float)s);
    }
    __syncthreads();

    float2* T = a.scratch + (size_t)blockIdx.x * N * P;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const unsigned long long g = (unsigned long long)(a.first_pair + pair);
        if (SH) sh_prepare(a, pair, sh_amp, sh_tab);    // followed by barriers in the row loop
        for (int row0 = 0; row0 < N; row0 += R) {
            const int nr = min(R, N - row0);
            if (RNG) {
                const int S = (N + 15) / 16;
                for (int idx = tid; idx < nr * S; idx += kThreads) {
                    const int rl = idx / S, t = idx % S, r = row0 + rl;
                    uint32_t mr[16], ma[16];
                    noise_block_fields((uint32_t)(r * S + t), g, k0, k1, mr, ma);
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int j = t + S * m;
                        if (j < N) rows[rl * N + j] = weighted_normal_m(mr[m], ma[m], a.weight[(size_t)r * N + j]);
                    }
                }
            } else {
                for (int idx = tid; idx < nr * N; idx += kThreads) {
                    const int rl = idx / N, j = idx % N, r = row0 + rl;
                    const float2 nz = a.noise[((size_t)pair * N + r) * N + j];
                    const float w0 = a.weight[(size_t)r * N + j];
                    rows[rl * N + j] = make_float2(nz.x * w0, nz.y * w0);
                }
            }
            __syncthreads();
            for (int idx = tid; idx < nr * P; idx += kThreads) {
                const int rl = idx / P, c = idx % P;
                const int kk = c + lo;                 // output column, 0 <= kk < N
                int ti = 0;
                float sr = 0.f, si = 0.f;
                const float2* row = rows + rl * N;
                for (int cp = 0; cp < N; ++cp) {
                    const float2 x = row[cp], t = tw[ti];
                    sr = fmaf(x.x, t.x, fmaf(-x.y, t.y, sr));
                    si = fmaf(x.x, t.y, fmaf(x.y, t.x, si));
                    ti += kk;
                    if (ti >= N) ti -= N;
                }
                T[(size_t)c * N + row0 + rl] = make_float2(sr, si);
            }
            __syncthreads();
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int idx = tid; idx < P * P; idx += kThreads) {
            const int c = idx / P, rr = idx % P;
            const int kk = rr + lo;
            const float2* col = T + (size_t)c * N;
            int ti = 0;
            float sr = 0.f, si = 0.f;
            for (int rp = 0; rp < N; ++rp) {
                const float2 x = __ldcg(col + rp), t = tw[ti];
                sr = fmaf(x.x, t.x, fmaf(-x.y, t.y, sr));
                si = fmaf(x.x, t.y, fmaf(x.y, t.x, si));
                ti += kk;
                if (ti >= N) ti -= N;
            }
            if (a.phs) {     // inspection seam (fastb_screens_crop): store phi, skip the detector
                const float sgn = ((kk + c + lo) & 1) ? -1.f : 1.f;
                float2 ph = make_float2(sgn * sr, sgn * si);
                if (SH) {
                    const float2 ex[3] = {a.sh_ex[c], a.sh_ex[P + c], a.sh_ex[2 * P + c]};
                    const float2 sp = sh_phase(sh_tab + rr * kShTab, ex);
                    ph.x += sp.x;
                    ph.y += sp.y;
                }
                float* dst = a.phs + (size_t)pair * 2 * P * P + (size_t)rr * P + c;
                dst[0] = ph.x;
                dst[(size_t)P * P] = ph.y;
                continue;
            }
            const float uu = a.u_t[(size_t)c * P + rr];
            if (SH) {
                const float sgn = ((kk + c + lo) & 1) ? -1.f : 1.f;
                const float2 ex[3] = {a.sh_ex[c], a.sh_ex[P + c], a.sh_ex[2 * P + c]};
                const float2 sp = sh_phase(sh_tab + rr * kShTab, ex);
                accumulate(make_float2(fmaf(sgn, sr, sp.x), fmaf(sgn, si, sp.y)), uu, uu, acc);
            } else {
                accumulate(make_float2(sr, si), uu, ((kk + c + lo) & 1) ? -uu : uu, acc);
            }
        }
        if (a.phs) {
            __syncthreads();          // scratch and tables are reused by the next pair
            continue;
        }
        finish_pair(a, pair, acc, red);
    }
}

// weight * sqrt(2 ln 2) for the radix kernel's device-RNG path (weighted_normal_s), interleaved so
// that thread u of a line fetches its registers m = 4j .. 4j+3 (cells u + S1 m) with one 128-bit load
// per j and the threads of a line read consecutive 16-byte words
__global__ void scale_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int n, int s1) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (long long)n * n) return;
    const int r = (int)(o / n), rem = (int)(o % n);
    const int j = rem / (4 * s1), u = (rem % (4 * s1)) / 4, q = rem % 4;
    out[o] = w[(size_t)r * n + u + s1 * (4 * j + q)] * kBoxMullerScale;
}

__global__ void transpose_u_kernel(const float* __restrict__ U, int P, float* __restrict__ u_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * P) return;
    const int r = i / P, c = i % P;
    u_t[(size_t)c * P + r] = U[i];
}

// u_p[(q * P + r) * 2 + j] = U[r][2q + j] (0 beyond the last column): U for the line-pair kernel
__global__ void pair_u_kernel(const float* __restrict__ U, int P, float* __restrict__ u_p) {
    const int PP = (P + 1) >> 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PP * P * 2) return;
    const int j = i & 1, r = (i >> 1) % P, q = (i >> 1) / P, c = 2 * q + j;
    u_p[i] = c < P ? U[(size_t)r * P + c] : 0.f;
}

__global__ void rng_dump_kernel(unsigned long long seed, unsigned long long g, int N, float2* tile,
                                long long chi_first, long long chi_count, float* chi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int S = (N + 15) / 16;
    if (tile && i < (long long)N * S) {
        const int r = (int)(i / S), t = (int)(i % S);
        uint32_t mr[16], ma[16];
        noise_block_fields((uint32_t)(r * S + t), g, (uint32_t)seed, (uint32_t)(seed >> 32), mr, ma);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int j = t + S * m;
            if (j < N) tile[(size_t)r * N + j] = weighted_normal_m(mr[m], ma[m], 1.0f);
        }
    }
    if (chi && i < chi_count) chi[i] = chi_normal(seed, (uint64_t)(chi_first + i));
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t sh_smem_bytes(bool sh, int n_pup) { return sh ? sizeof(float2) * (28 + (size_t)kShTab * n_pup) : 0; }

template <class F>
size_t radix_smem_bytes(bool sh, int n_pup, int threads, bool use_tma, int stage_shift = 0) {
    const int LPB = threads / F::S1;
    const size_t tma = (F::S1 <= 32 && use_tma) ? (size_t)(threads / 32) * (32 * F::E * 8 + sizeof(uint64_t)) : 0;
    const size_t tile = stage_shift ? sizeof(float2) * (size_t)LPB * ((size_t)n_pup << stage_shift) : 0;
    return sizeof(typename F::Tw) * ((size_t)F::kTwA + F::kTwB) + sizeof(float2) * (size_t)LPB * F::kBuf + tma + tile +
           sizeof(float) * 4 * (threads / 32) + sh_smem_bytes(sh, n_pup);
}

int direct_rows(int n) {
    int r = (int)((96 * 1024) / ((size_t)n * sizeof(float2)));
    return r < 1 ? 1 : (r > 8 ? 8 : r);
}
size_t direct_smem_bytes(int n, bool sh, int n_pup) {
    return sizeof(float2) * ((size_t)n + (size_t)direct_rows(n) * n) + sizeof(float) * 4 * (kThreads / 32) +
           sh_smem_bytes(sh, n_pup);
}

bool radix_ok(int n) { return n >= 64 && n <= 2048 && (n & (n - 1)) == 0; }

int sm_count(int* out) {
    int dev = 0;
    FASTB_CUDA(cudaGetDevice(&dev));
    FASTB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return FASTB_OK;
}

constexpr int kMaxCtasPerSm = 12;

int launch_kernel(void (*kern)(RunArgs), const RunArgs& args, int threads, size_t smem, int max_grid,
                  cudaStream_t st) {
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
    int per_sm = 0, sms = 0;
    FASTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) {
        set_error("screen_detect_radix: kernel does not fit (smem %zu B)", smem);
        return FASTB_ERR_UNSUPPORTED;
    }
    if (per_sm > kMaxCtasPerSm) per_sm = kMaxCtasPerSm;
    int rc = sm_count(&sms);
    if (rc) return rc;
    long long grid = (long long)per_sm * sms;
    if (grid > args.n_pairs) grid = args.n_pairs;
    if (grid > max_grid) grid = max_grid;
    kern<<<(unsigned)grid, threads, smem, st>>>(args);
    return check_launch("screen_detect_radix");
}

template <int LOG2N, int E>
int launch_radix_e(const RunArgs& args, bool rng, int max_grid, cudaStream_t st) {
    using C = RadixCfg<LOG2N, E>;
    using F = typename C::F;
    constexpr int T = C::kThreadsPerCta, M = C::kMinBlocks;
    const bool sh = args.sh_weight != nullptr;
    void (*kern)(RunArgs) = nullptr;
    if (sh) kern = rng ? screen_detect_radix<F, true, true, T, M> : screen_detect_radix<F, false, true, T, M>;
    else kern = rng ? screen_detect_radix<F, true, false, T, M> : screen_detect_radix<F, false, false, T, M>;
    bool use_tma = false;
    int threads = T;
    // Window-specialised instances (device RNG, no sub-harmonics): the smallest centred window
    // class that contains the crop.  FAST's pupil crop is centred and 1/6 .. 1/3 of the grid wide.
    if constexpr (E == 16 && LOG2N >= 8) {
        if (rng && !sh) {
            const int lo = args.lo, hi = args.lo + args.n_pup, c = F::N / 2;
            const int half = (c - lo) > (hi - c) ? (c - lo) : (hi - c);
            bool spec = lo <= c && hi >= c;
#ifdef FASTB_TUNE
            if (const char* e = getenv("FASTB_KEEP")) spec = spec && atoi(e) != 0;
#endif
#ifdef FASTB_TUNE
            // tuning builds only: FASTB_WSHAPE=<threads><min blocks> for the window-class-2 instance
            if constexpr (LOG2N == 8 || LOG2N == 9) {
                const char* e = getenv("FASTB_WSHAPE");
                const int v = e ? atoi(e) : 0;
                if (spec && v && half <= window_half<F::N>(2)) {
                    spec = false;
                    if (v == 1284) { kern = screen_detect_radix<F, true, false, 128, 4, 0, 2>; threads = 128; }
                    if (v == 1285) { kern = screen_detect_radix<F, true, false, 128, 5, 0, 2>; threads = 128; }
                    if (v == 1286) { kern = screen_detect_radix<F, true, false, 128, 6, 0, 2>; threads = 128; }
                    if (v == 2562) { kern = screen_detect_radix<F, true, false, 256, 2, 0, 2>; threads = 256; }
                    if (v == 2563) { kern = screen_detect_radix<F, true, false, 256, 3, 0, 2>; threads = 256; }
                    if (v == 648) { kern = screen_detect_radix<F, true, false, 64, 8, 0, 2>; threads = 64; }
                    if (v == 6410) { kern = screen_detect_radix<F, true, false, 64, 10, 0, 2>; threads = 64; }
                }
            }
#endif
            if (spec && half <= window_half<F::N>(1)) kern = screen_detect_radix<F, true, false, T, M, 0, 1>;
            else if (spec && half <= window_half<F::N>(2)) kern = screen_detect_radix<F, true, false, T, M, 0, 2>;
            else if (spec && half <= window_half<F::N>(3)) kern = screen_detect_radix<F, true, false, T, M, 0, 3>;
        }
    }
#ifdef FASTB_TUNE
    // tuning builds only: FASTB_SHAPE=<threads><minblocks> for the 32-element flavour
    if constexpr (E == 32) {
        if (rng && !sh) {
            const char* e = getenv("FASTB_SHAPE");
            const int v = e ? atoi(e) : 0;
            if (v == 1284) { kern = screen_detect_radix<F, true, false, 128, 4>; threads = 128; }
            if (v == 1282) { kern = screen_detect_radix<F, true, false, 128, 2>; threads = 128; }
            if (v == 2562) { kern = screen_detect_radix<F, true, false, 256, 2>; threads = 256; }
            if (v == 2561) { kern = screen_detect_radix<F, true, false, 256, 1>; threads = 256; }
            if (v == 646) { kern = screen_detect_radix<F, true, false, 64, 6>; threads = 64; }
        }
    }
    // tuning builds only: FASTB_SHAPE=<threads><min blocks> for the 16-element flavour
    if constexpr (E == 16) {
        if (rng && !sh) {
            const char* e = getenv("FASTB_SHAPE");
            const int v = e ? atoi(e) : 0;
            if constexpr (LOG2N <= 9) {
                if (v == 648) { kern = screen_detect_radix<F, true, false, 64, 8>; threads = 64; }
                if (v == 6410) { kern = screen_detect_radix<F, true, false, 64, 10>; threads = 64; }
                if (v == 6412) { kern = screen_detect_radix<F, true, false, 64, 12>; threads = 64; }
            }
            if constexpr (LOG2N <= 10) {
                if (v == 1284) { kern = screen_detect_radix<F, true, false, 128, 4>; threads = 128; }
                if (v == 1285) { kern = screen_detect_radix<F, true, false, 128, 5>; threads = 128; }
                if (v == 1286) { kern = screen_detect_radix<F, true, false, 128, 6>; threads = 128; }
                if (v == 1288) { kern = screen_detect_radix<F, true, false, 128, 8>; threads = 128; }
            }
            if (v == 2562) { kern = screen_detect_radix<F, true, false, 256, 2>; threads = 256; }
            if (v == 2563) { kern = screen_detect_radix<F, true, false, 256, 3>; threads = 256; }
            if (v == 2564) { kern = screen_detect_radix<F, true, false, 256, 4>; threads = 256; }
        }
    }
    // tuning builds only: FASTB_TMA=1|2|3 routes weights+scratch / scratch only / weights only
    // through per-warp TMA staging (cp.async.bulk + mbarrier) on the bench path
    if constexpr (E == 16) {
        if (rng && !sh) {
            const char* e = getenv("FASTB_TMA");
            const int v = e ? atoi(e) : 0;
            if (v == 1) { kern = screen_detect_radix<F, true, false, T, M, 1>; use_tma = true; }
            if (v == 2) { kern = screen_detect_radix<F, true, false, T, M, 2>; use_tma = true; }
            if (v == 3) { kern = screen_detect_radix<F, true, false, T, M, 3>; use_tma = true; }
        }
    }
#endif
    if (rng) {
        const long long n2 = (long long)F::N * F::N;
        scale_weight_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(args.weight, args.weight_s, F::N, F::S1);
        const int rc = check_launch("scale_weight_kernel");
        if (rc) return rc;
    }
    // N <= 512: stage two rows per line slot in shared memory and store them as one 16-byte word
    // per column (a scattered 8-byte store costs the L1 data pipe one wavefront per lane: 49 % of
    // all wavefronts at N = 256).  Same-box A/B: +1 % at N = 256, +4 % at N = 512, -2 % at N = 1024.
    // Skipped when the extra shared memory would cost a resident CTA.
    RunArgs a2 = args;
    a2.stage_shift = 0;
    size_t smem = radix_smem_bytes<F>(sh, args.n_pup, threads, use_tma);
    int want = LOG2N <= 9 ? 1 : 0;
#ifdef FASTB_TUNE
    if (const char* e = getenv("FASTB_STAGE")) want = atoi(e) ? 1 : 0;
#endif
    const size_t smem2 = radix_smem_bytes<F>(sh, args.n_pup, threads, false, 1);
    if (!use_tma && want && (F::N / (threads / F::S1)) % 2 == 0 && smem2 <= 227 * 1024) {
        int occ0 = 0, occ = 0;
        FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
        FASTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, kern, threads, smem));
        FASTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem2));
        if (occ >= occ0 && occ >= 1) {
            a2.stage_shift = 1;
            smem = smem2;
        }
    }
    return launch_kernel(kern, a2, threads, smem, max_grid, st);
}

// line-pair kernel: N <= 256: 128 threads x 5 CTAs/SM (96 registers); above: 256 x 2 (128 registers)
template <int LOG2N>
int launch_pair(const RunArgs& args, bool rng, int max_grid, cudaStream_t st) {
    using F = LineFFT<LOG2N, pc>;
    constexpr int T = LOG2N <= 8 ? 128 : 256, M = LOG2N <= 8 ? 5 : 2;
    const bool sh = args.sh_weight != nullptr;
    void (*kern)(RunArgs) = nullptr;
    if (sh) kern = rng ? screen_detect_pair<LOG2N, true, true, T, M> : screen_detect_pair<LOG2N, false, true, T, M>;
    else kern = rng ? screen_detect_pair<LOG2N, true, false, T, M> : screen_detect_pair<LOG2N, false, false, T, M>;
    int threads = T;
#ifdef FASTB_TUNE
    if (rng && !sh) {      // tuning builds only: FASTB_PAIR_SHAPE=<threads><min blocks>
        const char* e = getenv("FASTB_PAIR_SHAPE");
        const int v = e ? atoi(e) : 0;
        if (v == 1284) { kern = screen_detect_pair<LOG2N, true, false, 128, 4>; threads = 128; }
        if (v == 1285) { kern = screen_detect_pair<LOG2N, true, false, 128, 5>; threads = 128; }
        if constexpr (LOG2N >= 7) {
            if (v == 2562) { kern = screen_detect_pair<LOG2N, true, false, 256, 2>; threads = 256; }
        }
        if constexpr (LOG2N <= 9) {
            if (v == 648) { kern = screen_detect_pair<LOG2N, true, false, 64, 8>; threads = 64; }
            if (v == 646) { kern = screen_detect_pair<LOG2N, true, false, 64, 6>; threads = 64; }
        }
    }
#endif
    return launch_kernel(kern, args, threads, radix_smem_bytes<F>(sh, args.n_pup, threads, false), max_grid, st);
}

template <int LOG2N>
int launch_radix(const RunArgs& args, bool rng, int max_grid, cudaStream_t st) {
#ifdef FASTB_TUNE
    // tuning builds only: FASTB_E=32 selects the 32-element flavour for N = 512 / 1024
    if constexpr (LOG2N == 9 || LOG2N == 10) {
        const char* e = getenv("FASTB_E");
        if (e && atoi(e) == 32) return launch_radix_e<LOG2N, 32>(args, rng, max_grid, st);
    }
#endif
    return launch_radix_e<LOG2N, radix_default_e<LOG2N>()>(args, rng, max_grid, st);
}

}  // namespace
}  // namespace fastb

using namespace fastb;

static int validate_run(const FastbRunParams* p) {
    FASTB_REQUIRE(p, "fastb_screen_detect: NULL params");
    FASTB_REQUIRE(p->n >= 4 && (p->n % 2) == 0, "fastb_screen_detect: n=%d must be even and >= 4", p->n);
    FASTB_REQUIRE(p->n_pup >= 1 && p->n_pup <= p->n, "fastb_screen_detect: n_pup=%d outside 1..n", p->n_pup);
    FASTB_REQUIRE(p->lo >= 0 && p->lo + p->n_pup <= p->n, "fastb_screen_detect: crop [%d,%d) outside grid",
                  p->lo, p->lo + p->n_pup);
    FASTB_REQUIRE(p->n_pairs >= 0 && p->first_pair >= 0, "fastb_screen_detect: negative pair range");
    FASTB_REQUIRE(p->pairs_per_chunk > 0, "fastb_screen_detect: pairs_per_chunk must be > 0");
    FASTB_REQUIRE(p->algo >= FASTB_ALGO_AUTO && p->algo <= FASTB_ALGO_RADIX_PAIR, "fastb_screen_detect: bad algo");
    FASTB_REQUIRE(p->u_sum != 0.0, "fastb_screen_detect: u_sum is zero");
    if ((p->algo == FASTB_ALGO_RADIX || p->algo == FASTB_ALGO_RADIX_PAIR) && !radix_ok(p->n)) {
        set_error("fastb_screen_detect: radix path needs N = 64..2048 power of two, got %d", p->n);
        return FASTB_ERR_UNSUPPORTED;
    }
    if (p->n > 4096) {
        set_error("fastb_screen_detect: N=%d > 4096 not supported", p->n);
        return FASTB_ERR_UNSUPPORTED;
    }
    return FASTB_OK;
}

extern "C" int64_t fastb_screen_detect_workspace_bytes(const FastbRunParams* p) {
    if (validate_run(p)) return -1;
    int sms = 0;
    if (sm_count(&sms)) return -1;
    long long grid = (long long)sms * kMaxCtasPerSm;
    if (grid > p->n_pairs) grid = p->n_pairs;
    if (grid < 1) grid = 1;
    const size_t ut = align_up(sizeof(float) * (size_t)(p->n_pup + 1) * p->n_pup, 256) +
                      align_up(sizeof(float) * (size_t)p->n * p->n, 256);
    return (int64_t)(ut + (size_t)grid * p->n * (p->n_pup + 1) * sizeof(float2));
}

extern "C" int fastb_screen_detect(const FastbRunParams* p, const float* d_weight, const float* d_U,
                                   const float* d_chi, const float* d_noise, const FastbSubharm* sh,
                                   float* d_out_a, float* d_out_b, void* d_workspace,
                                   int64_t workspace_bytes, void* stream) {
    int rc = validate_run(p);
    if (rc) return rc;
    FASTB_REQUIRE(d_weight && d_U && d_out_a && d_out_b && d_workspace, "fastb_screen_detect: NULL pointer");
    FASTB_REQUIRE(((uintptr_t)d_weight & 15) == 0 && ((uintptr_t)d_workspace & 255) == 0,
                  "fastb_screen_detect: d_weight must be 16-byte and d_workspace 256-byte aligned");
    if (p->n_pairs == 0) return FASTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // [U table: transposed (P*P) or pair-interleaved ((P+1)*P) | weight * sqrt(2 ln 2) (N*N, radix
    //  kernel with device RNG) | scratch slots of N*(P+1) complex]
    const size_t ut0 = align_up(sizeof(float) * (size_t)(p->n_pup + 1) * p->n_pup, 256);
    const size_t ut = ut0 + align_up(sizeof(float) * (size_t)p->n * p->n, 256);
    const size_t slot = (size_t)p->n * (p->n_pup + 1) * sizeof(float2);
    FASTB_REQUIRE(workspace_bytes >= (int64_t)(ut + slot), "fastb_screen_detect: workspace too small (%lld B)",
                  (long long)workspace_bytes);
    long long max_grid = (long long)(((size_t)workspace_bytes - ut) / slot);
    if (max_grid > (1 << 20)) max_grid = 1 << 20;

    RunArgs a;
    a.n = p->n;
    a.n_pup = p->n_pup;
    a.lo = p->lo;
    a.coherent = p->coherent;
    a.n_pairs = p->n_pairs;
    a.first_pair = p->first_pair;
    a.ppc = p->pairs_per_chunk;
    a.seed = p->seed;
    a.inv_usum = (float)(1.0 / p->u_sum);
    a.sigma_chi = p->sigma_chi;
    a.weight = d_weight;
    a.weight_s = nullptr;
    a.u_t = (const float*)d_workspace;
    a.u_p = (const float2*)d_workspace;
    a.chi = d_chi;
    a.noise = (const float2*)d_noise;
    a.out_a = d_out_a;
    a.out_b = d_out_b;
    a.scratch = (float2*)((char*)d_workspace + ut);
    a.rows_per_block = direct_rows(p->n);
    a.stage_shift = 0;
    a.phs = nullptr;
#ifdef FASTB_TUNE_DBG
    a.dbg = getenv("FASTB_DBG") ? atoi(getenv("FASTB_DBG")) : 0;
#endif
    a.sh_weight = nullptr;
    a.sh_noise = a.sh_ex = a.sh_ey = a.sh_mean = nullptr;
    if (sh) {
        FASTB_REQUIRE(sh->d_weight && sh->d_ex && sh->d_ey && sh->d_mean,
                      "fastb_screen_detect: sub-harmonic tables must not be NULL");
        FASTB_REQUIRE((sh->d_noise == nullptr) == (d_noise == nullptr),
                      "fastb_screen_detect: d_noise and sh->d_noise must both be given or both be NULL");
        a.sh_weight = sh->d_weight;
        a.sh_noise = (const float2*)sh->d_noise;
        a.sh_ex = (const float2*)sh->d_ex;
        a.sh_ey = (const float2*)sh->d_ey;
        a.sh_mean = (const float2*)sh->d_mean;
    }

    const bool rng = d_noise == nullptr;
    // AUTO picks the one-line radix kernel: the line-pair kernel executes 22 % fewer instructions
    // but measured 4-10 % slower in its best shape (profiles/experiments_r01.txt)
    const bool use_pair = p->algo == FASTB_ALGO_RADIX_PAIR;
    if (use_pair) {
        const int cnt = ((p->n_pup + 1) / 2) * p->n_pup * 2;
        pair_u_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(d_U, p->n_pup, (float*)d_workspace);
        if ((rc = check_launch("pair_u_kernel"))) return rc;
        switch (p->n) {
            case 64: return launch_pair<6>(a, rng, (int)max_grid, st);
            case 128: return launch_pair<7>(a, rng, (int)max_grid, st);
            case 256: return launch_pair<8>(a, rng, (int)max_grid, st);
            case 512: return launch_pair<9>(a, rng, (int)max_grid, st);
            case 1024: return launch_pair<10>(a, rng, (int)max_grid, st);
            case 2048: return launch_pair<11>(a, rng, (int)max_grid, st);
            default: break;
        }
    }
    const int pp = p->n_pup * p->n_pup;
    transpose_u_kernel<<<(pp + 255) / 256, 256, 0, st>>>(d_U, p->n_pup, (float*)d_workspace);
    if ((rc = check_launch("transpose_u_kernel"))) return rc;

    const bool use_radix = p->algo == FASTB_ALGO_RADIX || (p->algo == FASTB_ALGO_AUTO && radix_ok(p->n));
    if (use_radix) {
        a.weight_s = (float*)((char*)d_workspace + ut0);      // filled by launch_radix_e when rng
        switch (p->n) {
            case 64: return launch_radix<6>(a, rng, (int)max_grid, st);
            case 128: return launch_radix<7>(a, rng, (int)max_grid, st);
            case 256: return launch_radix<8>(a, rng, (int)max_grid, st);
            case 512: return launch_radix<9>(a, rng, (int)max_grid, st);
            case 1024: return launch_radix<10>(a, rng, (int)max_grid, st);
            case 2048: return launch_radix<11>(a, rng, (int)max_grid, st);
            default: break;
        }
    }
    const bool has_sh = sh != nullptr;
    const size_t smem = direct_smem_bytes(p->n, has_sh, p->n_pup);
    void (*kern)(RunArgs) = nullptr;
    if (has_sh) kern = rng ? screen_detect_direct<true, true> : screen_detect_direct<false, true>;
    else kern = rng ? screen_detect_direct<true, false> : screen_detect_direct<false, false>;
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    if ((rc = sm_count(&sms))) return rc;
    long long grid = 2LL * sms;
    if (grid > p->n_pairs) grid = p->n_pairs;
    if (grid > max_grid) grid = max_grid;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(a);
    return check_launch("screen_detect_direct");
}

extern "C" int fastb_rng_dump(uint64_t seed, int64_t pair, int32_t n, float* d_noise_tile,
                              int64_t chi_first, int64_t chi_count, float* d_chi_normals, void* stream) {
    FASTB_REQUIRE(n >= 2 && (n % 2) == 0, "fastb_rng_dump: n must be even");
    FASTB_REQUIRE(pair >= 0 && chi_first >= 0 && chi_count >= 0, "fastb_rng_dump: negative index");
    long long work = d_noise_tile ? (long long)n * ((n + 15) / 16) : 0;
    if (d_chi_normals && chi_count > work) work = chi_count;
    if (work == 0) return FASTB_OK;
    rng_dump_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        seed, (unsigned long long)pair, n, (float2*)d_noise_tile, chi_first, chi_count, d_chi_normals);
    return check_launch("rng_dump_kernel");
}

extern "C" int fastb_screens_crop(const FastbRunParams* p, const float* d_weight, const float* d_noise,
                                  const FastbSubharm* sh, float* d_phs, void* d_workspace,
                                  int64_t workspace_bytes, void* stream) {
    int rc = validate_run(p);
    if (rc) return rc;
    FASTB_REQUIRE(d_weight && d_phs && d_workspace, "fastb_screens_crop: NULL pointer");
    if (p->n_pairs == 0) return FASTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ut = align_up(sizeof(float) * (size_t)(p->n_pup + 1) * p->n_pup, 256);
    const size_t slot = (size_t)p->n * (p->n_pup + 1) * sizeof(float2);
    FASTB_REQUIRE(workspace_bytes >= (int64_t)(ut + slot), "fastb_screens_crop: workspace too small");
    long long max_grid = (long long)(((size_t)workspace_bytes - ut) / slot);
    RunArgs a = {};
    a.n = p->n;
    a.n_pup = p->n_pup;
    a.lo = p->lo;
    a.n_pairs = p->n_pairs;
    a.first_pair = p->first_pair;
    a.ppc = p->pairs_per_chunk;
    a.seed = p->seed;
    a.weight = d_weight;
    a.noise = (const float2*)d_noise;
    a.scratch = (float2*)((char*)d_workspace + ut);
    a.rows_per_block = direct_rows(p->n);
    a.phs = d_phs;
    if (sh) {
        FASTB_REQUIRE(sh->d_weight && sh->d_ex && sh->d_ey && sh->d_mean,
                      "fastb_screens_crop: sub-harmonic tables must not be NULL");
        FASTB_REQUIRE((sh->d_noise == nullptr) == (d_noise == nullptr),
                      "fastb_screens_crop: d_noise and sh->d_noise must both be given or both be NULL");
        a.sh_weight = sh->d_weight;
        a.sh_noise = (const float2*)sh->d_noise;
        a.sh_ex = (const float2*)sh->d_ex;
        a.sh_ey = (const float2*)sh->d_ey;
        a.sh_mean = (const float2*)sh->d_mean;
    }
    const bool rng = d_noise == nullptr, has_sh = sh != nullptr;
    const size_t smem = direct_smem_bytes(p->n, has_sh, p->n_pup);
    void (*kern)(RunArgs) = nullptr;
    if (has_sh) kern = rng ? screen_detect_direct<true, true> : screen_detect_direct<false, true>;
    else kern = rng ? screen_detect_direct<true, false> : screen_detect_direct<false, false>;
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    if ((rc = sm_count(&sms))) return rc;
    long long grid = 2LL * sms;
    if (grid > p->n_pairs) grid = p->n_pairs;
    if (grid > max_grid) grid = max_grid;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(a);
    return check_launch("screen_detect_direct(screens)");
}
