// K2 radix kernels of ONE grid size: compiled once per LOG2N = 6..11 (-DFASTB_LOG2N=k, see
// build_fastb.py) so that the ~20 instances per size build in parallel.  Device code:
// screen_detect_kernel.cuh; C ABI and size dispatch: screen_detect.cu.
#include "screen_detect_kernel.cuh"

#ifndef FASTB_LOG2N
#error "compile with -DFASTB_LOG2N=6..11"
#endif
#define FASTB_CAT_(a, b) a##b
#define FASTB_CAT(a, b) FASTB_CAT_(a, b)

namespace fastb {
namespace {

constexpr int kLog2N = FASTB_LOG2N;
using Cfg = RadixCfg<kLog2N, 16>;
using F = Cfg::F;
constexpr int T = Cfg::kThreadsPerCta, M = Cfg::kMinBlocks;

// N <= 512: two rows staged per line slot and stored as one 16-byte word per column (compile-time: the
// staging tiles of the product shapes always fit, 32 KB / 201 KB at the widest crop)
constexpr int kStage = kLog2N <= 9 ? 1 : 0;
static_assert((F::N / (T / F::S1)) % 2 == 0, "row iterations come in pairs");

// LineFFT flavour only (FASTB_SPLIT=0): last FFT stage on warp shuffles where it measured faster, the radix-2
// stage of N = 512 (+5.4 %); the radix-4 stage of N = 1024 is 0.5 - 1 % slower than the shared-memory exchange
// (profiles/experiments_r02.txt).  The split-first flavour has no such stage.
constexpr bool kShuffle = F::kShflC && F::S1 == 32;

template <int RNG, bool SH, int WIN>
constexpr auto kern() { return screen_detect_radix<F, RNG, SH, T, M, 0, WIN, kShuffle, false, kStage>; }

// Window-specialised instances (no sub-harmonics): the smallest centred window class that
// contains the crop.  FAST's pupil crop is centred and 1/6 .. 1/3 of the grid wide.
int window_class(const RunArgs& a) {
    if (kLog2N < 8 || a.sh_weight) return 0;
    const int lo = a.lo, hi = a.lo + a.n_pup, c = F::N / 2;
    if (!(lo <= c && hi >= c)) return 0;
    const int half = (c - lo) > (hi - c) ? (c - lo) : (hi - c);
    for (int w = 1; w <= 3; ++w)
        if (half <= window_half<F::N>(w)) return w;
    return 0;
}

template <int RNG>
void (*pick(bool sh, int win))(RunArgs) {
    if (sh) return kern<RNG, true, 0>();
    if constexpr (kLog2N >= 8) {
        switch (win) {
            case 1: return kern<RNG, false, 1>();
            case 2: return kern<RNG, false, 2>();
            case 3: return kern<RNG, false, 3>();
            default: break;
        }
    }
    return kern<RNG, false, 0>();
}

}  // namespace

int FASTB_CAT(launch_radix_, FASTB_LOG2N)(const RunArgs& a, const RadixRequest& rq, cudaStream_t st) {
    const bool sh = a.sh_weight != nullptr;
    const int win = window_class(a);
    void (*k)(RunArgs) = rq.rng == kRngHost ? pick<kRngHost>(sh, win)
                         : rq.rng == kRngFast ? pick<kRngFast>(sh, win)
                                              : pick<kRngPhilox>(sh, win);
    if (rq.rng != kRngHost && !rq.prepared) {
        const int rc = prepare_weight_s(F::N, a.n_items > 1 ? a.n_items : 1, a.weight, a.weight_s, st);
        if (rc) return rc;
    }
    return launch_radix_instance<F>(k, a, T, false, kStage, rq.max_grid, st, true);
}

// line-pair kernel: N <= 256: 128 threads x 5 CTAs/SM (96 registers); above: 256 x 2 (128 registers)
int FASTB_CAT(launch_pair_, FASTB_LOG2N)(const RunArgs& a, const RadixRequest& rq, cudaStream_t st) {
    using FP = LineFFT<kLog2N, pc>;
    constexpr int TP = kLog2N <= 8 ? 128 : 256, MP = kLog2N <= 8 ? 5 : 2;
    const bool sh = a.sh_weight != nullptr, rng = rq.rng != kRngHost;
    void (*k)(RunArgs) = nullptr;
    if (sh) k = rng ? screen_detect_pair<kLog2N, true, true, TP, MP> : screen_detect_pair<kLog2N, false, true, TP, MP>;
    else k = rng ? screen_detect_pair<kLog2N, true, false, TP, MP> : screen_detect_pair<kLog2N, false, false, TP, MP>;
    return launch_kernel(k, a, TP, radix_smem_bytes<FP>(sh, a.n_pup, TP, false), rq.max_grid, st, "screen_detect_pair");
}

}  // namespace fastb
