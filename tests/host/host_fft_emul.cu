// CPU emulation of fast_b200/csrc/fft_core.cuh: runs the exact per-thread phases of the line
// FFT sequentially (threads emulated between sync points) and checks against a direct
// float64 DFT.  Built and run by tests/test_host_fft.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fast_b200/csrc/fft_core.cuh"

using namespace fastb;

template <int LOG2N>
double run_one(unsigned seed) {
    using F = LineFFT<LOG2N>;
    constexpr int N = F::N;
    std::vector<float2> x(N), twa(F::kTwA), twb(F::kTwB + 1), buf(F::kBuf), X(N);
    std::vector<int> seen(N, 0);
    srand(seed);
    for (int i = 0; i < N; ++i) {
        x[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    }
    for (int i = 0; i < F::kTwA; ++i) {
        const int ex = F::twa_exponent(i);
        twa[i] = make_float2((float)cos(2.0 * M_PI * ex / N), (float)sin(2.0 * M_PI * ex / N));
    }
    for (int i = 0; i < F::kTwB; ++i) {
        const int ex = F::twb_exponent(i);
        twb[i] = make_float2((float)cos(2.0 * M_PI * ex / N), (float)sin(2.0 * M_PI * ex / N));
    }
    float2 v[16];
    for (int t = 0; t < F::S1; ++t) {
        for (int m = 0; m < 16; ++m) v[m] = x[F::n_in(t, m)];
        F::phase_a(t, v, twa.data(), buf.data());
    }
    if (F::kThree) {
        std::vector<float2> keep(16 * F::S1);
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
                for (int e = 0; e < 16; ++e) { X[F::k_out(u, e)] = v[e]; seen[F::k_out(u, e)]++; }
            }
        } else {
            for (int u = 0; u < F::S1; ++u)
                for (int e = 0; e < 16; ++e) { X[F::k_out(u, e)] = keep[u * 16 + e]; seen[F::k_out(u, e)]++; }
        }
    } else {
        for (int u = 0; u < F::S1; ++u) {
            F::phase_c(u, v, buf.data());
            for (int e = 0; e < 16; ++e) { X[F::k_out(u, e)] = v[e]; seen[F::k_out(u, e)]++; }
        }
    }
    double worst = 0, scale = 0;
    for (int k = 0; k < N; ++k) {
        if (seen[k] != 1) { printf("N=%d: output %d produced %d times\n", N, k, seen[k]); return 1e9; }
        double re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            const double ang = 2.0 * M_PI * (double)(((long long)n * k) % N) / N;
            re += x[n].x * cos(ang) - x[n].y * sin(ang);
            im += x[n].x * sin(ang) + x[n].y * cos(ang);
        }
        worst = fmax(worst, hypot(X[k].x - re, X[k].y - im));
        scale = fmax(scale, hypot(re, im));
    }
    printf("N=%d S1=%d S2=%d SF=%d buf=%d  max_err/max_abs=%.3e\n", N, F::S1, F::S2, F::SF, F::kBuf, worst / scale);
    return worst / scale;
}

template <int LOG2N>
double run_32(unsigned seed) {
    using F = LineFFT32<LOG2N>;
    constexpr int N = F::N;
    std::vector<float2> x(N), twa(F::kTwA), buf(F::kBuf), X(N);
    std::vector<int> seen(N, 0);
    srand(seed);
    for (int i = 0; i < N; ++i)
        x[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    for (int i = 0; i < F::kTwA; ++i) {
        const int ex = F::twa_exponent(i);
        twa[i] = make_float2((float)cos(2.0 * M_PI * ex / N), (float)sin(2.0 * M_PI * ex / N));
    }
    float2 v[32];
    for (int t = 0; t < F::S1; ++t) {
        for (int m = 0; m < 32; ++m) v[m] = x[F::n_in(t, m)];
        F::phase_a(t, v, twa.data(), buf.data());
    }
    for (int u = 0; u < F::S1; ++u) {
        F::phase_b(u, v, buf.data());
        for (int e = 0; e < 32; ++e) { X[F::k_out(u, e)] = v[e]; seen[F::k_out(u, e)]++; }
    }
    double worst = 0, scale = 0;
    for (int k = 0; k < N; ++k) {
        if (seen[k] != 1) { printf("N=%d (E=32): output %d produced %d times\n", N, k, seen[k]); return 1e9; }
        double re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            const double ang = 2.0 * M_PI * (double)(((long long)n * k) % N) / N;
            re += x[n].x * cos(ang) - x[n].y * sin(ang);
            im += x[n].x * sin(ang) + x[n].y * cos(ang);
        }
        worst = fmax(worst, hypot(X[k].x - re, X[k].y - im));
        scale = fmax(scale, hypot(re, im));
    }
    printf("N=%d E=32 S1=%d buf=%d  max_err/max_abs=%.3e\n", N, F::S1, F::kBuf, worst / scale);
    return worst / scale;
}

int main() {
    double w = 0;
    w = fmax(w, run_32<9>(7));
    w = fmax(w, run_32<10>(8));
    w = fmax(w, run_one<6>(1));
    w = fmax(w, run_one<7>(2));
    w = fmax(w, run_one<8>(3));
    w = fmax(w, run_one<9>(4));
    w = fmax(w, run_one<10>(5));
    w = fmax(w, run_one<11>(6));
    printf("worst %.3e\n", w);
    return w < 2e-6 ? 0 : 1;
}
