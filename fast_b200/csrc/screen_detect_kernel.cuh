// K2 device code: phase-screen synthesis fused with the fibre-overlap detector (and, optionally,
// the K3 statistics).  Contract: include/fastb.h.  Included by the per-grid-size translation units
// (screen_detect_radix.cu, compiled once per LOG2N), by screen_detect.cu (direct kernel + C ABI) and
// by the tuning-only unit tune/screen_detect_tune.cu.
//
// One persistent CTA owns one complex transform ("pair" = two realisations) at a time:
//   pass 1  for every frequency row r': white noise (Philox + Box-Muller in registers, or the
//           caller's noise) x weight -> N-point line FFT in registers -> keep the n_pup output
//           columns of the pupil crop -> CTA-private scratch T[c][r'] (L2 resident)
//   pass 2  for every kept column c: N-point line FFT over r' -> keep the n_pup rows of the crop
//           -> U (cos phi, sin phi) accumulated in registers for Re and Im screens
//   final   fixed-order block reduction, exp(chi), normalisation -> 1 scalar per realisation,
//           moments / extrema / dB histogram of the results accumulated on the fly.
// Signs: the weight carries (-1)^(r'+c') and the output (-1)^(r+c), which turns the reference's
// centred (fftshift-ed) inverse DFT (fast/funcs.py:218 via aotools.ift2) into a plain one.
#pragma once
#include "fastb_common.cuh"
#include "fft_core.cuh"

namespace fastb {

constexpr int kThreads = 256;          // direct kernel

enum { kRngHost = 0, kRngPhilox = 1, kRngFast = 2 };

struct RunArgs {
    int n, n_pup, lo, coherent;
    long long n_pairs, first_pair, ppc;
    unsigned long long seed;
    float inv_usum, sigma_chi;
    const float* weight;      // N*N signed weight (n_items stacked tables in a batch)
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
    int noise_stride;         // direct kernel: noise blocks per row S of the device RNG (include/fastb.h)
    int stage_shift;          // radix kernel: 1 = stage two rows per line slot before storing, 0 = store directly
    // batch of configurations sharing the grid and U (FastbRunBatch): flattened pair index
    // q = item * ppi + g;  n_items <= 1: a single configuration, q = g
    int n_items;
    long long ppi;
    const float* item_sigma;              // n_items
    const unsigned long long* item_seed;  // n_items
    // fused K3 (FastbRunStats; st_sums == NULL = off).  Buffers are per item: [item][8], [item][2],
    // [item][nbins + 2]
    double* st_sums;
    double* st_minmax;
    unsigned long long* st_hist;
    double st_lo, st_hi;
    int st_nbins;
    // sub-harmonics (NULL weight = off)
    const float* sh_weight;   // 27
    const float2* sh_noise;   // n_pairs*27 or NULL
    const float2* sh_ex;      // 3*n_pup
    const float2* sh_ey;      // 3*n_pup
    const float2* sh_mean;    // 27
    float* phs;               // direct kernel only: write the cropped screens instead of detecting
    int dbg;                  // tuning builds: timing experiments (results wrong); 0 in the product
    // phase staggering (radix kernel; 0 = off): co-resident CTAs start `stagger` cycles apart per residency
    // slot (blockIdx / sms), the warps that share a scheduler inside one CTA `wstagger` cycles apart per pass
    int stagger, wstagger, sms;
};
#ifdef FASTB_TUNE_DBG
#define FASTB_DBG(a, bit) ((a).dbg & (bit))
#else
#define FASTB_DBG(a, bit) 0
#endif

// what the ABI layer asks of a per-size launcher
struct RadixRequest {
    int rng;          // kRng*
    int max_grid;     // scratch slots available
    bool prepared;    // weight_s is already valid in the workspace
};
int launch_radix_n(int log2n, const RunArgs& a, const RadixRequest& rq, cudaStream_t st);   // 6..11
int launch_pair_n(int log2n, const RunArgs& a, const RadixRequest& rq, cudaStream_t st);
int prepare_weight_s(int n, int n_items, const float* weight, float* weight_s, cudaStream_t st);
int radix_ctas_per_sm(int log2n);     // design occupancy (scratch sizing)

// a tuning build may register a selector that is asked first (tune/screen_detect_tune.cu);
// returns < 0 when it does not handle the request
typedef int (*TuneHook)(int log2n, const RunArgs& a, const RadixRequest& rq, cudaStream_t st);
extern TuneHook g_tune_hook;
// L2 residency hint for the scratch slots (launch_kernel): 0 off, 1 persisting access-policy window
extern int g_l2_persist;
// phase staggering in cycles (RunArgs.stagger / wstagger), 0 = off
extern int g_stagger, g_wstagger;

namespace {

// identity of the pair a CTA is working on
struct PairId {
    unsigned long long g;     // pair index inside its configuration (RNG counter)
    unsigned long long seed;
    int item;
    float sigma_chi;
};
__device__ __forceinline__ PairId pair_id(const RunArgs& a, long long pair) {
    PairId id;
    const long long q = a.first_pair + pair;
    id.item = 0;
    id.g = (unsigned long long)q;
    id.seed = a.seed;
    id.sigma_chi = a.sigma_chi;
    if (a.n_items > 1) {
        id.item = (int)(q / a.ppi);
        id.g = (unsigned long long)(q - (long long)id.item * a.ppi);
        id.seed = a.item_seed[id.item];
        id.sigma_chi = a.item_sigma[id.item];
    }
    return id;
}

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
__device__ void sh_prepare(const RunArgs& a, long long pair, const PairId& id, float2* amp, float2* tab) {
    const int tid = threadIdx.x;
    const unsigned long long g = id.g;
    if (tid < 14) {
        float2 n0, n1;
        if (a.sh_noise) {
            n0 = a.sh_noise[pair * 27 + 2 * tid];
            n1 = (2 * tid + 1 < 27) ? a.sh_noise[pair * 27 + 2 * tid + 1] : make_float2(0.f, 0.f);
        } else {
            const uint4 w = philox4x32_10((uint32_t)tid, (uint32_t)g, (uint32_t)(g >> 32), kStreamSubharm,
                                          (uint32_t)id.seed, (uint32_t)(id.seed >> 32));
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

// ---- fused K3: running statistics of one CTA (thread 0 only), kept in shared memory ----------
// st[0..5] = n, sum r, sum r^2, sum dB, sum dB^2, non-positive count; st[6] = min, st[7] = max;
// st[8] holds the item the sums belong to (as a double).  Flushed with atomics when the item
// changes and when the CTA is done: a handful of atomics per CTA and item instead of per pair.
constexpr int kStatWords = 9;
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void stats_reset(double* st, int item) {
#pragma unroll
    for (int k = 0; k < 6; ++k) st[k] = 0.0;
    st[6] = INFINITY;
    st[7] = -INFINITY;
    st[8] = (double)item;
}
__device__ __noinline__ void stats_flush(const RunArgs& a, double* st) {
    if (st[0] == 0.0) return;
    const int item = (int)st[8];
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(&a.st_sums[item * 8 + k], st[k]);
    atomic_min_double(&a.st_minmax[item * 2], st[6]);
    atomic_max_double(&a.st_minmax[item * 2 + 1], st[7]);
}
// one result (same arithmetic and bin rule as stats_kernel, fast_b200/csrc/stats.cu)
__device__ __noinline__ void stats_add(const RunArgs& a, double* st, int item, float r) {
    if ((int)st[8] != item) {
        stats_flush(a, st);
        stats_reset(st, item);
    }
    const double v = (double)r;
    st[0] += 1.0;
    st[1] += v;
    st[2] += v * v;
    st[6] = fmin(st[6], v);
    st[7] = fmax(st[7], v);
    unsigned long long* hist = a.st_hist + (size_t)item * (a.st_nbins + 2);
    if (v > 0.0) {
        const double db = 10.0 * log10(v);
        st[3] += db;
        st[4] += db * db;
        int b;
        if (db < a.st_lo) b = a.st_nbins;
        else if (db >= a.st_hi) b = a.st_nbins + 1;
        else {
            b = (int)((db - a.st_lo) * ((double)a.st_nbins / (a.st_hi - a.st_lo)));
            if (b >= a.st_nbins) b = a.st_nbins - 1;
        }
        atomicAdd(&hist[b], 1ULL);
    } else {
        st[5] += 1.0;
        atomicAdd(&hist[a.st_nbins], 1ULL);
    }
}

// fixed-order reduction of the 4 accumulators over the CTA, then the per-pair epilogue.  There is no
// barrier after the epilogue: `red` is next written in the following pair's reduction, which every
// thread reaches only after that pair's CTA-wide barrier between the passes, so thread 0's epilogue
// (exp, the stores, the statistics) overlaps the other warps' next rows.
template <int THREADS = kThreads>
__device__ void finish_pair(const RunArgs& a, long long pair, const PairId& id, float (&acc)[4], float* red,
                            double* st) {
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
        const long long g = (long long)id.g;
        const long long chunk = g / a.ppc, pp = g % a.ppc;
        const long long ia = chunk * 2 * a.ppc + pp, ib = ia + a.ppc;
        float chia, chib;
        if (a.chi) {
            chia = a.chi[ia];
            chib = a.chi[ib];
        } else {
            chia = id.sigma_chi * chi_normal(id.seed, (uint64_t)ia);
            chib = id.sigma_chi * chi_normal(id.seed, (uint64_t)ib);
        }
        const float ea = expf(chia) * a.inv_usum, eb = expf(chib) * a.inv_usum;
        const float zar = ea * t[0], zai = ea * t[1], zbr = eb * t[2], zbi = eb * t[3];
        const float ra = zar * zar + zai * zai, rb = zbr * zbr + zbi * zbi;
        if (a.coherent) {
            a.out_a[2 * pair] = zar;
            a.out_a[2 * pair + 1] = zai;
            a.out_b[2 * pair] = zbr;
            a.out_b[2 * pair + 1] = zbi;
        } else {
            a.out_a[pair] = ra;
            a.out_b[pair] = rb;
        }
        if (a.st_sums) {
            stats_add(a, st, id.item, ra);
            stats_add(a, st, id.item, rb);
        }
    }
}

// ---- TMA bulk copy (cp.async.bulk, 1-D) + mbarrier helpers (tuning flavour) ---------------------
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
//   N <= 256 : 128 threads x 4 CTAs/SM (128 registers) -- small CTAs keep the per-pair
//              barriers cheap and balance the pupil columns over 8 lines per iteration
//   N >= 512 : 512 threads x 1 CTA/SM (128 registers): one scratch slot per SM keeps the pass-1 ->
//              pass-2 intermediate of all SMs (148 x 8 n_pup N bytes) in or near the 126 MB L2
//              (same-box A/B, profiles/experiments_r02.txt: +3.4 % at N = 512 over 256 x 2,
//              +6.2 % at N = 1024 over 256 x 3)
// All use the 16-elements-per-thread FFT.  A 32-elements-per-thread flavour (512 = 32 x 16,
// 1024 = 32 x 32: one exchange per line, 168 registers) is kept for tuning builds: it measured the
// same throughput in every CTA shape.
template <int LOG2N, int E>
struct RadixCfg;
// FASTB_SPLIT=1 (tuning builds): N >= 512 uses the split-first line FFT (fft_core.cuh, LineFFTSplit) instead of
// LineFFT with its shuffle / shared-memory last stage.  Measured slower (C4 -8 %, C5 -5 %: fewer FFT instructions,
// but 16 more shared-memory stores and 8 more loads per line and thread on an L1 data pipe that is already 71 % busy)
#ifndef FASTB_SPLIT
#define FASTB_SPLIT 0
#endif
template <int LOG2N, bool SPLIT = (FASTB_SPLIT != 0 && LOG2N >= 9)>
struct RadixFft {
    using type = LineFFT<LOG2N>;
};
template <int LOG2N>
struct RadixFft<LOG2N, true> {
    using type = LineFFTSplit<LOG2N>;
};
template <int LOG2N>
struct RadixCfg<LOG2N, 16> {
    using F = typename RadixFft<LOG2N>::type;
    static constexpr int kThreadsPerCta = LOG2N <= 8 ? 128 : 512;
    static constexpr int kMinBlocks = LOG2N <= 8 ? 4 : 1;
};
template <int LOG2N>
struct RadixCfg<LOG2N, 32> {
    using F = LineFFT32<LOG2N>;
    static constexpr int kThreadsPerCta = 128;
    static constexpr int kMinBlocks = 3;
};

// One loop body serves both passes (keeps the kernel inside the instruction cache): iterations
// [0, n1) are frequency rows (noise -> FFT -> pruned store to T[c][r']), iterations [n1, n1+n2)
// are kept columns (load T[c][:] -> FFT -> detector accumulation).  No CTA-wide barrier inside
// a pass.  TMA != 0 (tuning builds, E = 16, lines inside a warp): the warp's contiguous input
// block of the next iteration is fetched by one cp.async.bulk into a per-warp stage
// (1: weights and scratch, 2: scratch only, 3: weights only) -- measured slower, off by default.
// WIN (compile time) promises that the crop lies inside the centred window of half-width
// window_half<N>(WIN) (0: no promise): only the registers keep_mask<F>() names can then hold a kept
// output, and the compiler drops the last-stage butterflies (and shared loads) that feed the others.
// SHFL: the last radix-2 / radix-4 stage of N = 512 / 1024 runs on warp shuffles (fft_core.cuh,
// phase_c_shfl) instead of a second shared-memory exchange.
// FASTB_PF=1 (tuning builds): pass 2 reads its columns through L1 and touches the next column one iteration ahead.
// Measured without effect (C2 -0.05 %, C4 -1.4 %, C5 +0.2 %; profiles/experiments_r02.txt, 15): the long-scoreboard
// samples of pass 2 (8 % of all at C2) are latency the other warps already cover.
#ifndef FASTB_PF
#define FASTB_PF 0
#endif
template <int N>
constexpr int window_half(int win) { return win == 1 ? N / 8 : win == 2 ? 3 * N / 16 : win == 3 ? N / 4 : N; }

// ONCHIP (tuning flavour, N = 256): the pass-1 -> pass-2 intermediate T lives in shared memory
// (n_pup x (N + 1) complex, one CTA per SM) instead of the CTA-private global slot.
// STAGE: two-row store staging of pass 1 fixed at compile time (0 / 1), or -1 = RunArgs.stage_shift decides.
template <class F, int RNG, bool SH, int THREADS, int MINB, int TMA = 0, int WIN = 0, bool SHFL = true,
          bool ONCHIP = false, int STAGE = -1>
__global__ void __launch_bounds__(THREADS, MINB) screen_detect_radix(const __grid_constant__ RunArgs a) {
    constexpr int N = F::N, S1 = F::S1, E = F::E, LPB = THREADS / S1;
    constexpr int NP = N + 1;                                   // row stride of the on-chip T
    static_assert(THREADS % S1 == 0 && LPB >= 1 && (S1 <= 32 || LPB <= 15), "line/barrier layout");
    static_assert(E == 16 || E == 32, "elements per thread");
    constexpr unsigned kKeep = WIN == 0 ? 0xffffffffu : keep_mask<F>(window_half<N>(WIN));
    constexpr bool kTma = (S1 <= 32) && TMA != 0 && E == 16;
    constexpr bool kTmaW = kTma && (TMA == 1 || TMA == 3);      // weight rows through the stage
    constexpr bool kTmaT = kTma && (TMA == 1 || TMA == 2);      // scratch columns through the stage
    constexpr bool kShfl = SHFL && F::kShflC;
    constexpr bool kPrefetch = FASTB_PF != 0 && !kTma && !ONCHIP;
    constexpr int kLinesPerWarp = S1 <= 32 ? 32 / S1 : 1;
    constexpr int kWarps = THREADS / 32;
    constexpr int kStageBytes = 32 * E * 8;

    using Tw = typename F::Tw;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tw* twa = reinterpret_cast<Tw*>(smem_raw);
    Tw* twb = twa + F::kTwA;
    float2* bufs = reinterpret_cast<float2*>(twb + F::kTwB);
    // pass-1 output staging: per line slot R = 2^stage_shift (1 or 2) planes of n_pup kept outputs
    const int rs = (kTma || ONCHIP) ? 0 : (STAGE >= 0 ? STAGE : a.stage_shift), R = 1 << rs;
    float2* tiles = bufs + LPB * F::kBuf;
    unsigned char* stage_all = reinterpret_cast<unsigned char*>(tiles + (rs ? LPB * R * a.n_pup : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + (kTma ? kWarps * kStageBytes : 0));
    double* st = reinterpret_cast<double*>(bars + (kTma ? kWarps : 0));    // 8-byte aligned by construction
    float* red = reinterpret_cast<float*>(st + kStatWords);
    float2* sh_amp = reinterpret_cast<float2*>(red + 4 * kWarps);     // SH only
    float2* sh_tab = sh_amp + 28;
    float2* Ts = sh_amp + (SH ? 28 + kShTab * a.n_pup : 0);           // ONCHIP only: n_pup x NP

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
    if (tid == 0) stats_reset(st, 0);
    if (kTma) fence_proxy_async();            // mbarrier init visible to the async proxy
    __syncthreads();
    auto spin = [](long long cycles) {
        const long long t0 = clock64();
        while (clock64() - t0 < cycles) {}
    };
    if (a.stagger) spin((long long)(blockIdx.x / a.sms) * a.stagger);

    float2* T = a.scratch + (size_t)blockIdx.x * N * P;
    const int n1 = (N + LPB - 1) / LPB, n2 = (P + LPB - 1) / LPB;
    static_assert(ONCHIP || N % LPB == 0, "rows per iteration");

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
    const float* wtab = a.weight;
    auto prefetch = [&](int itn) {
        const bool rows_n = itn < n1;
        if ((rows_n && !kTmaW) || (!rows_n && !kTmaT)) return;
        const int line0 = (rows_n ? itn : itn - n1) * LPB + warp * kLinesPerWarp;
        const int nlines = rows_n ? kLinesPerWarp : min(kLinesPerWarp, P - line0);
        if (nlines <= 0 || lane != 0) return;
        const uint32_t bytes = (uint32_t)nlines * N * (rows_n ? 4u : 8u);
        const void* src = rows_n ? (const void*)(wtab + (size_t)line0 * N) : (const void*)(T + (size_t)line0 * N);
        fence_proxy_async();                  // earlier generic reads of the stage precede the async write
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stage, src, bytes, bar);
    };

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const PairId id = pair_id(a, pair);
        const unsigned long long g = id.g;
        const uint32_t k0 = (uint32_t)id.seed, k1 = (uint32_t)(id.seed >> 32);
        const size_t item_off = (size_t)id.item * N * N;
        wtab = a.weight + item_off;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (SH) sh_prepare(a, pair, id, sh_amp, sh_tab);   // table visible after the barrier at it == n1
        if (kTma) prefetch(0);
        if (kWarps > 4 && a.wstagger) spin((long long)(warp >> 2) * a.wstagger);
        for (int it = 0; it < n1 + n2; ++it) {
            const bool rows = it < n1;
            if (it == n1) {
                if (kTmaT) fence_proxy_async();       // T was written through the generic proxy
                __syncthreads();                      // every row of T is stored before a column is read
                if (kTma) prefetch(n1);
                if (kWarps > 4 && a.wstagger) spin((long long)(warp >> 2) * a.wstagger);
            }
            // pass 1: a line slot takes R (1 or 2) consecutive rows in R consecutive iterations, so
            // that its kept outputs leave as 16-byte stores of two adjacent rows per column
            const int sub = it & (R - 1);
            const int line = rows ? (((it >> rs) * LPB + ln) << rs) + sub       // r'
                                  : (it - n1) * LPB + ln;                        // c
            // last column iteration: warps whose lines all lie beyond the crop have nothing to do
            // (line barriers involve only the threads of that line)
            if (!rows && line - (ln % kLinesPerWarp) >= P) continue;
            if (ONCHIP && rows && line - (ln % kLinesPerWarp) >= N) continue;     // partial last row iteration

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
                } else if (RNG != kRngHost) {
                    const float4* wq = reinterpret_cast<const float4*>(a.weight_s + item_off + (size_t)line * N) + u;
#pragma unroll
                    for (int j = 0; j < E / 4; ++j) {
                        const float4 t = __ldg(wq + j * S1);
                        w[4 * j] = t.x;
                        w[4 * j + 1] = t.y;
                        w[4 * j + 2] = t.z;
                        w[4 * j + 3] = t.w;
                    }
                } else {
                    const float* wrow = wtab + (size_t)line * N;
#pragma unroll
                    for (int m = 0; m < E; ++m) w[m] = FASTB_DBG(a, 2) ? 1.f + m : __ldg(wrow + u + S1 * m);
                }
                if (RNG != kRngHost && FASTB_DBG(a, 16)) {
#pragma unroll
                    for (int m = 0; m < E; ++m) v[m] = make_float2(w[m], w[m] * u);
                } else if (RNG != kRngHost) {
                    // thread (line, u) owns noise blocks t' = u + S1 h of its row: cell j of block h
                    // is element m = (E/16) j + h (include/fastb.h)
#pragma unroll
                    for (int h = 0; h < E / 16; ++h) {
                        uint32_t mr[16], ma[16];
                        const uint32_t block = (uint32_t)(line * (N / 16) + u + S1 * h);
                        if (FASTB_DBG(a, 64)) {                 // timing only: a cheap hash instead of Philox
                            uint32_t x = block * 0x9E3779B9u + (uint32_t)g;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                x = x * 1664525u + 1013904223u;
                                mr[j] = x >> 9;
                                ma[j] = (x * 0x85EBCA6Bu) >> 9;
                            }
                        } else if (RNG == kRngFast) {
                            noise_block_fields_fast(block, g, k0, k1, mr, ma);
                        } else {
                            noise_block_fields(block, g, k0, k1, mr, ma);
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
            } else if (ONCHIP) {
                const float2* tcol = Ts + (size_t)(line < P ? line : 0) * NP;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = tcol[u + S1 * m];
            } else if (kPrefetch) {
                // the column through L1, and one word of each 128-byte line of the NEXT column touched now, so that
                // its loads hit L1 an iteration later instead of waiting for L2 (pass 2 spent 18 % of its FFT
                // samples and 35 % of its detector samples on the long scoreboard).  T is only read after the
                // barrier that follows its last store, and the SM's own stores keep its L1 coherent.
                const float2* tcol = T + (size_t)(line < P ? line : 0) * N;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = __ldca(tcol + u + S1 * m);
                const int nxt = line + LPB < P ? line + LPB : P - 1;
                unsigned touch;
                asm volatile("ld.global.ca.b32 %0, [%1];" : "=r"(touch) : "l"(T + (size_t)nxt * N + (E * u)));
                if (u < (P * 4 + 127) / 128)
                    asm volatile("ld.global.ca.b32 %0, [%1];" : "=r"(touch) : "l"(a.u_t + (size_t)nxt * P + min(32 * u, P - 1)));
            } else {
                const float2* tcol = T + (size_t)(line < P ? line : 0) * N;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = FASTB_DBG(a, 4) ? make_float2(m, u) : __ldcg(tcol + u + S1 * m);
            }

            if constexpr (kShfl) F::template run_shfl<kKeep>(u, v, twa, twb, buf, sync);
            else F::run(u, v, twa, twb, buf, sync);

            if (ONCHIP && rows) {
                float2* tb = Ts + (kb * NP + line);           // &Ts[(k - lo) * NP + r'] at k_off = 0
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (((kKeep >> e) & 1u) && (need & (1u << e))) tb[F::k_off(e) * NP] = v[e];
            } else if (rows && rs == 0) {
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
                    if constexpr (WIN != 0) {
                        // the window class bounds the crop width: a fixed number of predicated column
                        // steps instead of a counted loop (its control flow was 2.3 % of all instructions)
                        constexpr int kSteps = (2 * window_half<N>(WIN) + S1 - 1) / S1;
#pragma unroll
                        for (int j = 0; j < kSteps; ++j) {
                            const int c = u + j * S1;
                            if (c < P) {
                                const float2 x0 = tile[c], x1 = tile[P + c];
                                __stcg(reinterpret_cast<float4*>(tr + (long long)c * N),
                                       make_float4(x0.x, x0.y, x1.x, x1.y));
                            }
                        }
                    } else {
#pragma unroll 2
                        for (int c = u; c < P; c += S1) {
                            const float2 x0 = tile[c], x1 = tile[P + c];
                            __stcg(reinterpret_cast<float4*>(tr + (long long)c * N), make_float4(x0.x, x0.y, x1.x, x1.y));
                        }
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
        finish_pair<THREADS>(a, pair, id, acc, red, st);
    }
    if (a.st_sums && tid == 0) stats_flush(a, st);
}

// ---- line-PAIR kernel: two adjacent lines per thread group, planar packed FP32 -----------------
// Same algorithm and results as screen_detect_radix, but every thread carries the same position
// of TWO adjacent lines (rows 2p, 2p+1 in pass 1; pupil columns 2q, 2q+1 in pass 2) as planar
// pairs (fft_core.cuh, value type pc), so that all FFT arithmetic and most of Box-Muller are
// packed FP32 (FADD2 / FMUL2 / FFMA2).  Scratch layout: T4[q][r'] = float4 (re(2q), re(2q+1),
// im(2q), im(2q+1)), which pass 2 reads with one 128-bit load per element pair.  Cross-check
// flavour (FASTB_ALGO_RADIX_PAIR): single configuration, Philox4x32-10 stream or host noise.
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
    double* st = reinterpret_cast<double*>(bufs + LPB * F::kBuf);
    float* red = reinterpret_cast<float*>(st + kStatWords);
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
    if (tid == 0) stats_reset(st, 0);
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
        const PairId id = pair_id(a, pair);
        const unsigned long long g = id.g;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (SH) sh_prepare(a, pair, id, sh_amp, sh_tab);   // table visible after the barrier at it == n1
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
        finish_pair<THREADS>(a, pair, id, acc, red, st);
    }
    if (a.st_sums && tid == 0) stats_flush(a, st);
}

// ---- host-side helpers shared by the launchers ---------------------------------------------------
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline size_t sh_smem_bytes(bool sh, int n_pup) { return sh ? sizeof(float2) * (28 + (size_t)kShTab * n_pup) : 0; }

template <class F>
size_t radix_smem_bytes(bool sh, int n_pup, int threads, bool use_tma, int stage_shift = 0) {
    const int LPB = threads / F::S1;
    const size_t tma = (F::S1 <= 32 && use_tma) ? (size_t)(threads / 32) * (32 * F::E * 8 + sizeof(uint64_t)) : 0;
    const size_t tile = stage_shift ? sizeof(float2) * (size_t)LPB * ((size_t)n_pup << stage_shift) : 0;
    return sizeof(typename F::Tw) * ((size_t)F::kTwA + F::kTwB) + sizeof(float2) * (size_t)LPB * F::kBuf + tma + tile +
           sizeof(double) * kStatWords + sizeof(float) * 4 * (threads / 32) + sh_smem_bytes(sh, n_pup);
}

inline int sm_count(int* out) {
    int dev = 0;
    FASTB_CUDA(cudaGetDevice(&dev));
    FASTB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return FASTB_OK;
}

constexpr int kMaxCtasPerSm = 12;

inline int launch_kernel(void (*kern)(RunArgs), const RunArgs& args, int threads, size_t smem, int max_grid,
                         cudaStream_t st, const char* what = "screen_detect_radix") {
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
    int per_sm = 0, sms = 0;
    FASTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) {
        set_error("%s: kernel does not fit (smem %zu B)", what, smem);
        return FASTB_ERR_UNSUPPORTED;
    }
    if (per_sm > kMaxCtasPerSm) per_sm = kMaxCtasPerSm;
    int rc = sm_count(&sms);
    if (rc) return rc;
    long long grid = (long long)per_sm * sms;
    if (grid > args.n_pairs) grid = args.n_pairs;
    if (grid > max_grid) grid = max_grid;
    RunArgs launch_args = args;
    launch_args.sms = sms;
    launch_args.stagger = g_stagger;
    launch_args.wstagger = g_wstagger;
    bool windowed = false;
    if (g_l2_persist) {
        // keep the CTA-private pass-1 -> pass-2 intermediate in L2: persisting window over the slots in use
        int dev = 0, max_win = 0, max_persist = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        size_t bytes = (size_t)grid * args.n * args.n_pup * sizeof(float2);
        if (max_win > 0 && max_persist > 0) {
            static int limit_set_for = -1;
            if (limit_set_for != dev) {
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
                limit_set_for = dev;
            }
            cudaStreamAttrValue attr = {};
            attr.accessPolicyWindow.base_ptr = (void*)args.scratch;
            attr.accessPolicyWindow.num_bytes = bytes < (size_t)max_win ? bytes : (size_t)max_win;
            attr.accessPolicyWindow.hitRatio = bytes <= (size_t)max_persist ? 1.0f : (float)max_persist / (float)bytes;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            windowed = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
            cudaGetLastError();
        }
    }
    kern<<<(unsigned)grid, threads, smem, st>>>(launch_args);
    const int rc_launch = check_launch(what);
    if (windowed) {
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
    }
    return rc_launch;
}

// Launch one radix instance; decides on the two-row store staging (N <= 512: +1 % at N = 256, +4 % at
// N = 512, -2 % at N = 1024 in same-box A/B) unless the extra shared memory would cost a resident CTA.
template <class F>
int launch_radix_instance(void (*kern)(RunArgs), const RunArgs& args, int threads, bool use_tma, int want_stage,
                          int max_grid, cudaStream_t st, bool stage_fixed = false) {
    const bool sh = args.sh_weight != nullptr;
    RunArgs a2 = args;
    a2.stage_shift = 0;
    size_t smem = radix_smem_bytes<F>(sh, args.n_pup, threads, use_tma);
    const size_t smem2 = radix_smem_bytes<F>(sh, args.n_pup, threads, false, 1);
    if (stage_fixed) {           // the instance was compiled with STAGE = want_stage
        a2.stage_shift = want_stage;
        return launch_kernel(kern, a2, threads, want_stage ? smem2 : smem, max_grid, st);
    }
    if (!use_tma && want_stage && (F::N / (threads / F::S1)) % 2 == 0 && smem2 <= 227 * 1024) {
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

}  // namespace
}  // namespace fastb
