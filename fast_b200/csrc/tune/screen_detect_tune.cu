// TUNING BUILDS ONLY (FASTB_TUNE=1 python build_fastb.py): environment-driven alternative shapes and
// variants of the K2 radix kernel for same-box A/B runs.  Never part of the product library: the
// default build does not compile this unit, and the product selection lives in screen_detect_radix.cu.
// The unit registers a selector (g_tune_hook) that screen_detect.cu asks first.
//
//   FASTB_SHAPE=<threads>x<min CTAs/SM>   e.g. 128x4, 128x5, 256x2, 256x3, 512x1
//   FASTB_SHFL=0                          last radix-2/4 stage through shared memory instead of shuffles
//   FASTB_KEEP=0                          generic instance instead of the window-specialised one
//   FASTB_STAGE=0|1                       two-row store staging off / on
//   FASTB_TMA=1|2|3                       cp.async.bulk staging of weights+scratch / scratch / weights
//   FASTB_E=32                            32 elements per thread (N = 512, 1024)
//   FASTB_L2PERSIST=1                     persisting-L2 access-policy window over the scratch slots
//   FASTB_STAGGER=<cycles>, FASTB_WSTAGGER=<cycles>   phase staggering of co-resident CTAs / same-scheduler warps
//   FASTB_ONCHIP=<threads>                N = 256 only: pass-1 -> pass-2 intermediate in shared memory, one CTA
//                                         of 256 / 384 threads per SM (no global scratch traffic at all)
// Serves the device-RNG, no-sub-harmonics instances of N = 256, 512, 1024 whose crop is the standard
// centred one (window class 2, 2, 1); anything else falls through to the product selection.
#include "../screen_detect_kernel.cuh"

#include <stdlib.h>
#include <string.h>

namespace fastb {
namespace {

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int LOG2N, int RNG, int WIN, bool SHFL>
void (*shape_kernel(int threads, int minb, int* t_out))(RunArgs) {
    using F = LineFFT<LOG2N>;
#define FASTB_SHAPE_CASE(T, M)                                                     \
    if (threads == T && minb == M) {                                                \
        *t_out = T;                                                                  \
        return screen_detect_radix<F, RNG, false, T, M, 0, WIN, SHFL>;               \
    }
    if constexpr (LOG2N <= 9) {
        FASTB_SHAPE_CASE(128, 4) FASTB_SHAPE_CASE(128, 5) FASTB_SHAPE_CASE(128, 6)
    }
    if constexpr (LOG2N == 10) { FASTB_SHAPE_CASE(128, 4) }
    FASTB_SHAPE_CASE(256, 2) FASTB_SHAPE_CASE(256, 3) FASTB_SHAPE_CASE(256, 4)
    FASTB_SHAPE_CASE(512, 1)
#undef FASTB_SHAPE_CASE
    return nullptr;
}

template <int LOG2N, int WIN>
int tune_size(const RunArgs& a, const RadixRequest& rq, cudaStream_t st) {
    using Cfg = RadixCfg<LOG2N, 16>;
    using F = typename Cfg::F;
    if (rq.rng == kRngHost || a.sh_weight) return -1;
    const char* shape = getenv("FASTB_SHAPE");
    const int shfl = env_int("FASTB_SHFL", 1), keep = env_int("FASTB_KEEP", 1);
    const int stage = env_int("FASTB_STAGE", LOG2N <= 9 ? 1 : 0), tma = env_int("FASTB_TMA", 0);
    const int e = env_int("FASTB_E", 16);
    if constexpr (LOG2N == 8) {
        const int onchip = env_int("FASTB_ONCHIP", 0);
        if (onchip) {
            const int c0 = F::N / 2, h0 = (c0 - a.lo) > (a.lo + a.n_pup - c0) ? (c0 - a.lo) : (a.lo + a.n_pup - c0);
            if (!(a.lo <= c0 && a.lo + a.n_pup >= c0 && h0 <= window_half<F::N>(WIN))) return -1;
            void (*ko)(RunArgs) = nullptr;
            if (onchip == 384) ko = rq.rng == kRngFast ? screen_detect_radix<F, kRngFast, false, 384, 1, 0, WIN, true, true>
                                                        : screen_detect_radix<F, kRngPhilox, false, 384, 1, 0, WIN, true, true>;
            else if (onchip == 256) ko = rq.rng == kRngFast ? screen_detect_radix<F, kRngFast, false, 256, 1, 0, WIN, true, true>
                                                            : screen_detect_radix<F, kRngPhilox, false, 256, 1, 0, WIN, true, true>;
            if (!ko) return -1;
            const size_t smem = radix_smem_bytes<F>(false, a.n_pup, onchip, false) + sizeof(float2) * (size_t)a.n_pup * (F::N + 1);
            if (smem > 227 * 1024) return -1;
            RunArgs a2 = a;
            a2.stage_shift = 0;
            return launch_kernel(ko, a2, onchip, smem, rq.max_grid, st, "screen_detect_radix(onchip)");
        }
    }
    if (!shape && shfl && keep && stage == (LOG2N <= 9 ? 1 : 0) && !tma && e == 16) return -1;
    // the crop must be the one the window class was chosen for
    const int c = F::N / 2, half = (c - a.lo) > (a.lo + a.n_pup - c) ? (c - a.lo) : (a.lo + a.n_pup - c);
    const bool win_ok = a.lo <= c && a.lo + a.n_pup >= c && half <= window_half<F::N>(WIN);
    int threads = Cfg::kThreadsPerCta, minb = Cfg::kMinBlocks;
    if (shape && sscanf(shape, "%dx%d", &threads, &minb) != 2) return -1;
    void (*k)(RunArgs) = nullptr;
    int t = threads;
    bool use_tma = false;
    if constexpr (LOG2N == 9 || LOG2N == 10) {
        if (e == 32) {
            using F32 = typename RadixCfg<LOG2N, 32>::F;
            k = rq.rng == kRngFast ? screen_detect_radix<F32, kRngFast, false, 128, 3>
                                   : screen_detect_radix<F32, kRngPhilox, false, 128, 3>;
            return launch_radix_instance<F32>(k, a, 128, false, 0, rq.max_grid, st);
        }
    }
    if (tma) {
        constexpr int T = Cfg::kThreadsPerCta, M = Cfg::kMinBlocks;
        if (tma == 1) k = screen_detect_radix<F, kRngPhilox, false, T, M, 1>;
        else if (tma == 2) k = screen_detect_radix<F, kRngPhilox, false, T, M, 2>;
        else k = screen_detect_radix<F, kRngPhilox, false, T, M, 3>;
        t = T;
        use_tma = true;
    } else if (keep && win_ok) {
        if (rq.rng == kRngFast) k = shfl ? shape_kernel<LOG2N, kRngFast, WIN, true>(threads, minb, &t)
                                         : shape_kernel<LOG2N, kRngFast, WIN, false>(threads, minb, &t);
        else k = shfl ? shape_kernel<LOG2N, kRngPhilox, WIN, true>(threads, minb, &t)
                      : shape_kernel<LOG2N, kRngPhilox, WIN, false>(threads, minb, &t);
    } else {
        k = shfl ? shape_kernel<LOG2N, kRngPhilox, 0, true>(threads, minb, &t)
                 : shape_kernel<LOG2N, kRngPhilox, 0, false>(threads, minb, &t);
    }
    if (!k) return -1;
    return launch_radix_instance<F>(k, a, t, use_tma, stage, rq.max_grid, st);
}

int tune_hook(int log2n, const RunArgs& a, const RadixRequest& rq, cudaStream_t st) {
    switch (log2n) {
        case 8: return tune_size<8, 2>(a, rq, st);
        case 9: return tune_size<9, 2>(a, rq, st);
        case 10: return tune_size<10, 1>(a, rq, st);
        default: break;
    }
    return -1;
}

struct Registrar {
    Registrar() {
        g_tune_hook = tune_hook;
        g_l2_persist = env_int("FASTB_L2PERSIST", 0);      // FASTB_L2PERSIST=1: persisting L2 window over the scratch
        g_stagger = env_int("FASTB_STAGGER", 0);           // cycles between the start of co-resident CTAs
        g_wstagger = env_int("FASTB_WSTAGGER", 0);         // cycles between same-scheduler warps of a CTA, per pass
    }
} g_registrar;

}  // namespace
}  // namespace fastb
