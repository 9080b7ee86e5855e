// Chirp-z (Bluestein) building blocks shared by the K2 kernel for grids that are not a power of two
// (screen_detect_bluestein.cu, screen_detect_bluestein_m.cu) and the TEMPORAL layer screens
// (layer_screens_fft.cu).
//   X[k] = c[k] sum_n (x[n] c[n]) conj(c[k - n]),  c[m] = e^{i pi m^2 / N}
// The circular convolution of length M = 2^LOG2M runs on the register line FFT G (inverse sign,
// unnormalised):  y = conj(G(conj(G(a) G(b)))) / M.
#pragma once
#include "fastb_common.cuh"
#include "fft_core.cuh"

namespace fastb {

int bluestein_log2m(int n, int n_out);                       // smallest M = 2^l >= n + n_out - 1 (l >= 6)

// ---- layer screens (full-size outputs): tables chirp[N] then bhat[M] = G(b) / M, natural order ----
size_t bluestein_table_bytes(int n, int n_out);
// tables for outputs k in [lo, lo + n_out) of an n-point transform, float64 arithmetic
int bluestein_prepare(int n, int n_out, int lo, void* tables, cudaStream_t st);

// ---- K2: geometry of an (N, n_pup) problem on the length-M transform ------------------------------
// S1 = M / 16 threads serve a line and thread u holds the inputs n = u + S1 m, m < 16.  Only the
// MC = ceil(N / S1) first of them can be non-zero; the kernels are compiled per class C = cell PAIRS per
// thread (5: up to 10 cells and any crop, 6, 7, 8: exactly 2C - 1 or 2C cells), which also bounds the
// crop: n_pup <= M - N + 1 <= (18 - 2C) S1 for C > 5, so the second transform of the convolution is
// pruned to the low output window at compile time.
struct BlueGeom {
    int log2m, M, S1, C;
};
BlueGeom blue_geom(int n, int n_pup);
// class of an N-point grid on S1 threads per line: cell pairs per thread, at least 5
FASTB_HD constexpr int blue_cell_pairs(int n, int s1) {
    const int mc = (n + s1 - 1) / s1, c = (mc + 1) / 2;
    return c < 5 ? 5 : c;
}
// outputs a problem of class C can want: below (18 - 2C) S1 (any for the catch-all class 5)
FASTB_HD constexpr int blue_output_bound(int c, int s1) { return c == 5 ? 16 * s1 : (18 - 2 * c) * s1; }
// does the transform F leave output k = u + S1 e in register e of thread u, i.e. in its own input order?
template <class F>
constexpr bool blue_identity_layout() {
    for (int u = 0; u < F::S1; ++u)
        for (int e = 0; e < 16; ++e)
            if (F::k_out(u, e) != u + F::S1 * e) return false;
    return true;
}
// registers of F that can hold an output below `w`
template <class F>
constexpr unsigned blue_keep_mask_low(int w) {
    unsigned m = 0;
    for (int u = 0; u < F::S1; ++u)
        for (int e = 0; e < 16; ++e)
            if (F::k_out(u, e) < w) m |= 1u << e;
    return m;
}
// complex chirped weight copies, one float4 per cell pair: entry (row r, pair j < C, thread u) at
// (r C + j) S1 + u holds wc[r][u + S1 2j], wc[r][u + S1 (2j + 1)], wc = weight c[col] c[row] (0 beyond N)
size_t bluestein_weight_bytes(int n, int n_pup, int n_items);
// tables of the K2 kernel: bhat in the register order of the line FFT (S1 rows of 18 float2, 16 used:
// row u, entry e = G(b')[k_out(u, e)] / M for the kernel shifted to the crop, b'[d] = conj(c[d + lo]))
// followed by the output chirp c[lo + k'], k' < n_pup
size_t bluestein_k2_table_bytes(int n, int n_pup);
struct RunArgs;
struct RadixRequest;
bool bluestein_ok(int n, int n_pup);           // even N, not a radix size, N + n_pup - 1 <= 2048
int bluestein_ctas_per_sm(int n, int n_pup);   // design occupancy (scratch sizing)
// fills the K2 tables and, unless weight == NULL, the chirped weight copies of n_items stacked tables
int bluestein_prepare_k2(int n, int n_pup, int lo, int n_items, const float* weight, void* wcq, void* tables,
                         cudaStream_t st);
int launch_bluestein(const RunArgs& a, const RadixRequest& rq, const void* tables, cudaStream_t st);

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// chirp-z of the line held as v[m] = a[u + S1 m] (a = x c, zero beyond N): on return v[e] holds
// sum_n a[n] conj(c[k - n]) for k = k_out(u, e) (valid for the outputs the tables were built for)
template <class F, class Sync>
__device__ __forceinline__ void chirp_convolve(int u, float2 (&v)[16], const typename F::Tw* twa,
                                               const typename F::Tw* twb, float2* buf, const float2* bhat,
                                               Sync sync) {
    // ONE copy of the line FFT in the instruction stream, executed twice (the kernels that inline this are
    // instruction-cache bound otherwise: ncu `no_instruction` stalls of 1.0 per issue with two copies per pass)
#pragma unroll 1
    for (int rep = 0; rep < 2; ++rep) {
        F::run(u, v, twa, twb, buf, sync);
        if (rep == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int k = F::k_base(u) + F::k_off(e);
                const float2 z = cmul(v[e], bhat[k]);
                buf[k] = make_float2(z.x, -z.y);              // conj(A B), natural order
            }
            sync();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = buf[u + F::S1 * m];
            sync();
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e].y = -v[e].y;            // y = conj(G(.)); the 1/M sits in bhat
}

}  // namespace fastb
