// CPU emulation of one line of the chirp-z K2 kernel (fast_b200/csrc/screen_detect_bluestein_m.cu): the same
// geometry helpers (bluestein.cuh), the same register line FFT (fft_core.cuh) run thread after thread between the
// sync points, the Bhat table in register order, the shifted kernel, the class pruning -- against a direct float64
// DFT of the wanted outputs.  Built and run by tests/test_host_fft.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fast_b200/csrc/bluestein.cuh"

using namespace fastb;

static void chirp(long long m, int N, double* c, double* s) {
    const double ph = M_PI * (double)(((m % (2LL * N)) * (m % (2LL * N))) % (2LL * N)) / N;
    *c = cos(ph);
    *s = sin(ph);
}

// the line FFT of all S1 threads; regs[u][m] in, regs[u][e] out (register e of thread u holds k_out(u, e))
template <class F>
static void fft_all(std::vector<float2>& regs, const std::vector<float2>& twa, const std::vector<float2>& twb,
                    std::vector<float2>& buf) {
    float2 v[16];
    for (int t = 0; t < F::S1; ++t) {
        for (int m = 0; m < 16; ++m) v[m] = regs[t * 16 + m];
        F::phase_a(t, v, twa.data(), buf.data());
    }
    if (F::kThree) {
        for (int u = 0; u < F::S1; ++u) {
            F::phase_b(u, v, twb.data(), buf.data());
            for (int e = 0; e < 16; ++e) regs[u * 16 + e] = v[e];
        }
        if (F::S2 > 1) {
            for (int u = 0; u < F::S1; ++u) {
                for (int e = 0; e < 16; ++e) v[e] = regs[u * 16 + e];
                F::phase_b_store(u, v, buf.data());
            }
            for (int u = 0; u < F::S1; ++u) {
                F::phase_c(u, v, buf.data());
                for (int e = 0; e < 16; ++e) regs[u * 16 + e] = v[e];
            }
        }
    } else {
        for (int u = 0; u < F::S1; ++u) {
            F::phase_c(u, v, buf.data());
            for (int e = 0; e < 16; ++e) regs[u * 16 + e] = v[e];
        }
    }
}

template <int LOG2M>
static double run_case(int N, int lo, int P, unsigned seed) {
    using F = LineFFT<LOG2M>;
    constexpr int M = F::N, S1 = F::S1;
    if (N + P - 1 > M || (LOG2M > 6 && N + P - 1 <= M / 2)) {
        printf("bad case N=%d P=%d for M=%d\n", N, P, M);
        return 1e9;
    }
    const int C = blue_cell_pairs(N, S1);
    const unsigned keep = blue_keep_mask_low<F>(blue_output_bound(C, S1));
    constexpr bool ident = blue_identity_layout<F>();
    // class invariants: no cell at or beyond m = 2C lies inside the grid; every wanted output sits in a kept register
    for (int u = 0; u < S1; ++u)
        for (int m = 2 * C; m < 16; ++m)
            if (u + S1 * m < N) return 1e9;
    for (int u = 0; u < S1; ++u)
        for (int e = 0; e < 16; ++e)
            if (F::k_out(u, e) < P && !((keep >> e) & 1u)) return 1e9;

    srand(seed);
    std::vector<float2> x(N);
    for (int i = 0; i < N; ++i) x[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<float2> twa(F::kTwA + 1), twb(F::kTwB + 1), buf(F::kBuf);
    for (int i = 0; i < F::kTwA; ++i) {
        const int ex = F::twa_exponent(i);
        twa[i] = make_float2((float)cos(2.0 * M_PI * ex / M), (float)sin(2.0 * M_PI * ex / M));
    }
    for (int i = 0; i < F::kTwB; ++i) {
        const int ex = F::twb_exponent(i);
        twb[i] = make_float2((float)cos(2.0 * M_PI * ex / M), (float)sin(2.0 * M_PI * ex / M));
    }
    // Bhat[q] = (1/M) sum_d conj(c[d + lo]) e^{2 pi i (d mod M) q / M}, -N < d < P, tabulated as (u, e) -> k_out(u, e)
    std::vector<float2> bhatp((size_t)S1 * 16);
    for (int u = 0; u < S1; ++u)
        for (int e = 0; e < 16; ++e) {
            const int q = F::k_out(u, e);
            double sr = 0, si = 0;
            for (int d = 1 - N; d <= P - 1; ++d) {
                const int j = ((d % M) + M) % M;
                double c, s;
                chirp(llabs((long long)d + lo), N, &c, &s);
                const double a = 2.0 * M_PI * (double)(((long long)j * q) % M) / M;
                // conj(c) e^{ia} = (c cos a + s sin a) + i (c sin a - s cos a)
                sr += c * cos(a) + s * sin(a);
                si += c * sin(a) - s * cos(a);
            }
            bhatp[u * 16 + e] = make_float2((float)(sr / M), (float)(si / M));
        }
    // a[n] = x[n] c[n] in the input registers of the transform; cells m >= 2C are structural zeros
    std::vector<float2> regs((size_t)S1 * 16);
    for (int u = 0; u < S1; ++u)
        for (int m = 0; m < 16; ++m) {
            const int n = u + S1 * m;
            float2 z = make_float2(0.f, 0.f);
            if (m < 2 * C && n < N) {
                double c, s;
                chirp(n, N, &c, &s);
                z = make_float2((float)(x[n].x * c - x[n].y * s), (float)(x[n].x * s + x[n].y * c));
            }
            regs[u * 16 + m] = z;
        }
    fft_all<F>(regs, twa, twb, buf);
    for (int u = 0; u < S1; ++u)
        for (int e = 0; e < 16; ++e) {
            const float2 z = cmul(regs[u * 16 + e], bhatp[u * 16 + e]);
            regs[u * 16 + e] = make_float2(z.x, -z.y);
        }
    if (!ident) {                                   // back to the input order through the line buffer
        std::vector<float2> nat(M);
        for (int u = 0; u < S1; ++u)
            for (int e = 0; e < 16; ++e) nat[F::k_out(u, e)] = regs[u * 16 + e];
        for (int u = 0; u < S1; ++u)
            for (int m = 0; m < 16; ++m) regs[u * 16 + m] = nat[u + S1 * m];
    }
    fft_all<F>(regs, twa, twb, buf);
    double worst = 0, scale = 0;
    int seen = 0;
    for (int u = 0; u < S1; ++u)
        for (int e = 0; e < 16; ++e) {
            const int kp = F::k_out(u, e);
            if (!((keep >> e) & 1u) || kp >= P) continue;
            ++seen;
            double c, s;
            chirp(kp + lo, N, &c, &s);
            const float2 y = regs[u * 16 + e];              // X = c conj(y)
            const double gr = y.x * c + y.y * s, gi = y.x * s - y.y * c;
            double re = 0, im = 0;
            for (int n = 0; n < N; ++n) {
                const double ang = 2.0 * M_PI * (double)(((long long)n * (kp + lo)) % N) / N;
                re += x[n].x * cos(ang) - x[n].y * sin(ang);
                im += x[n].x * sin(ang) + x[n].y * cos(ang);
            }
            worst = fmax(worst, hypot(gr - re, gi - im));
            scale = fmax(scale, hypot(re, im));
        }
    if (seen != P) return 1e9;
    printf("chirp-z N=%d lo=%d P=%d M=%d C=%d keep=0x%04x identity=%d  max_err/max_abs=%.3e\n", N, lo, P, M, C, keep,
           (int)ident, worst / scale);
    return worst / scale;
}

int main() {
    double w = 0;
    w = fmax(w, run_case<8>(164, 41, 82, 1));      // the reference's auto-sized example grid: class 6
    w = fmax(w, run_case<8>(100, 35, 30, 2));      // class 5
    w = fmax(w, run_case<8>(200, 78, 44, 3));      // class 7
    w = fmax(w, run_case<8>(236, 108, 20, 4));     // class 8
    w = fmax(w, run_case<6>(20, 0, 20, 5));
    w = fmax(w, run_case<6>(58, 26, 6, 6));
    w = fmax(w, run_case<7>(104, 40, 24, 7));
    w = fmax(w, run_case<7>(120, 50, 8, 8));
    w = fmax(w, run_case<9>(300, 60, 180, 9));
    w = fmax(w, run_case<9>(460, 210, 50, 10));
    w = fmax(w, run_case<10>(700, 250, 200, 11));
    w = fmax(w, run_case<10>(900, 400, 120, 12));
    w = fmax(w, run_case<11>(1500, 500, 549, 13));
    static_assert(blue_identity_layout<LineFFT<8>>() && !blue_identity_layout<LineFFT<9>>() &&
                      !blue_identity_layout<LineFFT<7>>(), "only M = 256 chains in registers");
    static_assert(blue_cell_pairs(164, 16) == 6 && blue_cell_pairs(20, 4) == 5 && blue_cell_pairs(236, 16) == 8, "classes");
    printf("worst %.3e\n", w);
    return w < 5e-6 ? 0 : 1;
}
