// CPU emulation of fast_b200/csrc/fft_core.cuh: runs the exact per-thread phases of the line
// FFT sequentially (all threads of a line between two sync points, then the next phase) and
// checks against a direct float64 DFT -- for the one-line value type (float2), the line-pair
// type (pc: two different lines, planar shared-memory layout) and the 32-elements-per-thread
// tuning flavour.  Built and run by tests/test_host_fft.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fast_b200/csrc/fft_core.cuh"

using namespace fastb;

static float2 lane(const float2& v, int) { return v; }
static float2 lane(const pc& v, int l) { return l == 0 ? make_float2(v.re.x, v.im.x) : make_float2(v.re.y, v.im.y); }
static void set_lanes(float2& v, const float2& a, const float2&) { v = a; }
static void set_lanes(pc& v, const float2& a, const float2& b) {
    v.re = make_float2(a.x, b.x);
    v.im = make_float2(a.y, b.y);
}

template <class Tw>
static std::vector<Tw> table(int n, int N, int (*expo)(int)) {
    std::vector<Tw> t(n + 1);
    for (int i = 0; i < n; ++i) {
        const int ex = expo(i);
        t[i] = make_tw((float)cos(2.0 * M_PI * ex / N), (float)sin(2.0 * M_PI * ex / N), (Tw*)nullptr);
    }
    return t;
}

// compare the outputs X[l][k] of `lines` lines with the direct DFT of x[l]
static double check(const std::vector<float2> (&x)[2], const std::vector<float2> (&X)[2],
                    const std::vector<int>& seen, int lines, const char* label) {
    const int N = (int)x[0].size();
    for (int k = 0; k < N; ++k)
        if (seen[k] != 1) {
            printf("%s: output %d produced %d times\n", label, k, seen[k]);
            return 1e9;
        }
    double worst = 0, scale = 0;
    for (int l = 0; l < lines; ++l)
        for (int k = 0; k < N; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < N; ++n) {
                const double ang = 2.0 * M_PI * (double)(((long long)n * k) % N) / N;
                re += x[l][n].x * cos(ang) - x[l][n].y * sin(ang);
                im += x[l][n].x * sin(ang) + x[l][n].y * cos(ang);
            }
            worst = fmax(worst, hypot(X[l][k].x - re, X[l][k].y - im));
            scale = fmax(scale, hypot(re, im));
        }
    printf("%s  max_err/max_abs=%.3e\n", label, worst / scale);
    return worst / scale;
}

template <int LOG2N, class V>
double run_one(unsigned seed) {
    using F = LineFFT<LOG2N, V>;
    using Tw = typename F::Tw;
    constexpr int N = F::N, L = F::kLines;
    std::vector<float2> x[2], X[2];
    std::vector<int> seen(N, 0);
    srand(seed);
    for (int l = 0; l < 2; ++l) {
        x[l].resize(N);
        X[l].resize(N);
        for (int i = 0; i < N; ++i)
            x[l][i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    }
    std::vector<Tw> twa = table<Tw>(F::kTwA, N, F::twa_exponent), twb = table<Tw>(F::kTwB, N, F::twb_exponent);
    std::vector<float2> buf(F::kBuf);
    std::vector<V> keep((size_t)16 * F::S1);
    V v[16];
    auto emit = [&](int u) {
        for (int e = 0; e < 16; ++e) {
            for (int l = 0; l < L; ++l) X[l][F::k_out(u, e)] = lane(v[e], l);
            seen[F::k_out(u, e)]++;
        }
    };
    for (int t = 0; t < F::S1; ++t) {
        for (int m = 0; m < 16; ++m) set_lanes(v[m], x[0][F::n_in(t, m)], x[1][F::n_in(t, m)]);
        F::phase_a(t, v, twa.data(), buf.data());
    }
    if (F::kThree) {
        for (int u = 0; u < F::S1; ++u) {
            F::phase_b(u, v, twb.data(), buf.data());
            for (int e = 0; e < 16; ++e) keep[u * 16 + e] = v[e];
        }
        if (F::S2 > 1) {
            for (int u = 0; u < F::S1; ++u) {
                for (int e = 0; e < 16; ++e) v[e] = keep[u * 16 + e];
                F::phase_b_store(u, v, buf.data());
            }
            for (int u = 0; u < F::S1; ++u) {
                F::phase_c(u, v, buf.data());
                emit(u);
            }
        } else {
            for (int u = 0; u < F::S1; ++u) {
                for (int e = 0; e < 16; ++e) v[e] = keep[u * 16 + e];
                emit(u);
            }
        }
    } else {
        for (int u = 0; u < F::S1; ++u) {
            F::phase_c(u, v, buf.data());
            emit(u);
        }
    }
    char label[96];
    snprintf(label, sizeof label, "N=%d lines=%d S1=%d S2=%d SF=%d buf=%d", N, L, F::S1, F::S2, F::SF, F::kBuf);
    return check(x, X, seen, L, label);
}

template <int LOG2N>
double run_32(unsigned seed) {
    using F = LineFFT32<LOG2N>;
    constexpr int N = F::N;
    std::vector<float2> x[2], X[2];
    std::vector<int> seen(N, 0);
    srand(seed);
    x[0].resize(N);
    X[0].resize(N);
    for (int i = 0; i < N; ++i)
        x[0][i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<float2> twa = table<float2>(F::kTwA, N, F::twa_exponent), buf(F::kBuf);
    float2 v[32];
    for (int t = 0; t < F::S1; ++t) {
        for (int m = 0; m < 32; ++m) v[m] = x[0][F::n_in(t, m)];
        F::phase_a(t, v, twa.data(), buf.data());
    }
    for (int u = 0; u < F::S1; ++u) {
        F::phase_b(u, v, buf.data());
        for (int e = 0; e < 32; ++e) {
            X[0][F::k_out(u, e)] = v[e];
            seen[F::k_out(u, e)]++;
        }
    }
    char label[96];
    snprintf(label, sizeof label, "N=%d E=32 S1=%d buf=%d", N, F::S1, F::kBuf);
    return check(x, X, seen, 1, label);
}

// the split-first flavour (N = 512, 1024, 2048): stage 0 of every thread, then the 256-point transforms
template <int LOG2N>
double run_split(unsigned seed) {
    using F = LineFFTSplit<LOG2N>;
    constexpr int N = F::N;
    std::vector<float2> x[2], X[2];
    std::vector<int> seen(N, 0);
    srand(seed);
    x[0].resize(N);
    X[0].resize(N);
    for (int i = 0; i < N; ++i)
        x[0][i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<float2> twa = table<float2>(F::kTwA, N, F::twa_exponent), twb = table<float2>(F::kTwB, N, F::twb_exponent);
    std::vector<float2> buf(F::kBuf);
    std::vector<float2> keep((size_t)16 * F::S1);
    float2 v[16];
    for (int u = 0; u < F::S1; ++u) {
        for (int m = 0; m < 16; ++m) v[m] = x[0][F::n_in(u, m)];
        F::stage0(u, v, twb.data(), buf.data());
    }
    for (int u = 0; u < F::S1; ++u) {                       // all gathers before any phase-A store (same regions)
        F::gather0(u, v, buf.data());
        for (int e = 0; e < 16; ++e) keep[u * 16 + e] = v[e];
    }
    for (int u = 0; u < F::S1; ++u) {
        for (int e = 0; e < 16; ++e) v[e] = keep[u * 16 + e];
        F::Sub::phase_a(u % 16, v, twa.data(), buf.data() + (u / 16) * F::kRegion);
    }
    for (int u = 0; u < F::S1; ++u) {
        F::Sub::phase_b(u % 16, v, nullptr, buf.data() + (u / 16) * F::kRegion);
        for (int e = 0; e < 16; ++e) {
            X[0][F::k_out(u, e)] = v[e];
            seen[F::k_out(u, e)]++;
        }
    }
    char label[96];
    snprintf(label, sizeof label, "N=%d split R=%d S1=%d buf=%d", N, F::R, F::S1, F::kBuf);
    return check(x, X, seen, 1, label);
}

// keep_mask<F>(half) (compile-time output pruning) against a brute-force scan of k_out, and the
// register counts DESIGN.md quotes for the bench crops
template <int LOG2N>
int check_keep_masks() {
    using F = LineFFT<LOG2N, float2>;
    constexpr int N = F::N;
    int bad = 0;
    const int halves[3] = {N / 8, 3 * N / 16, N / 4};
    constexpr unsigned m1 = keep_mask<F>(N / 8), m2 = keep_mask<F>(3 * N / 16), m3 = keep_mask<F>(N / 4);
    const unsigned got[3] = {m1, m2, m3};
    for (int i = 0; i < 3; ++i) {
        unsigned want = 0;
        for (int u = 0; u < F::S1; ++u)
            for (int e = 0; e < 16; ++e) {
                const int k = F::k_out(u, e);
                if (k >= N / 2 - halves[i] && k < N / 2 + halves[i]) want |= 1u << e;
            }
        if (want != got[i]) ++bad;
        printf("N=%d window half=%d keeps %d of 16 registers (mask 0x%04x)%s\n", N, halves[i],
               __builtin_popcount(got[i]), got[i], want == got[i] ? "" : "  MISMATCH");
    }
    if ((m1 & ~m2) || (m2 & ~m3)) ++bad;          // nested windows -> nested masks
    return bad;
}

int main() {
    int bad = check_keep_masks<6>() + check_keep_masks<7>() + check_keep_masks<8>() + check_keep_masks<9>() +
              check_keep_masks<10>() + check_keep_masks<11>();
    if (__builtin_popcount(keep_mask<LineFFT<8, float2>>(48)) != 6) ++bad;     // C2: 82 of 256 -> 6 of 16
    if (__builtin_popcount(keep_mask<LineFFT<9, float2>>(96)) != 6) ++bad;     // C4: 162 of 512 -> 6 of 16
    if (__builtin_popcount(keep_mask<LineFFT<10, float2>>(128)) != 4) ++bad;   // C5: 162 of 1024 -> 4 of 16
    if (bad) {
        printf("keep_mask check failed (%d)\n", bad);
        return 2;
    }
    double w = 0;
    w = fmax(w, run_one<6, float2>(1));
    w = fmax(w, run_one<7, float2>(2));
    w = fmax(w, run_one<8, float2>(3));
    w = fmax(w, run_one<9, float2>(4));
    w = fmax(w, run_one<10, float2>(5));
    w = fmax(w, run_one<11, float2>(6));
    w = fmax(w, run_one<6, pc>(11));
    w = fmax(w, run_one<7, pc>(12));
    w = fmax(w, run_one<8, pc>(13));
    w = fmax(w, run_one<9, pc>(14));
    w = fmax(w, run_one<10, pc>(15));
    w = fmax(w, run_one<11, pc>(16));
    w = fmax(w, run_32<9>(7));
    w = fmax(w, run_32<10>(8));
    w = fmax(w, run_split<9>(21));
    w = fmax(w, run_split<10>(22));
    w = fmax(w, run_split<11>(23));
    // the split flavour keeps fewer registers for the bench crops: C4 6 of 16 (as before), C5 4 of 16, and the
    // kept ones are whole outputs of the last 16-point DFT
    if (__builtin_popcount(keep_mask<LineFFTSplit<9>>(96)) != 6) return 3;
    if (__builtin_popcount(keep_mask<LineFFTSplit<10>>(128)) != 4) return 3;
    printf("worst %.3e\n", w);
    return w < 2e-6 ? 0 : 1;
}
