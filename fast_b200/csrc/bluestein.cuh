// Chirp-z (Bluestein) building blocks shared by the K2 kernel for grids that are not a power of two
// (screen_detect_bluestein.cu) and the TEMPORAL layer screens (layer_screens_fft.cu).
//   X[k] = c[k] sum_n (x[n] c[n]) conj(c[k - n]),  c[m] = e^{i pi m^2 / N}
// The circular convolution of length M = 2^LOG2M runs on the register line FFT G (inverse sign,
// unnormalised):  y = conj(G(conj(G(a) G(b)))) / M.  Tables (float2): chirp[N] then bhat[M] = G(b) / M.
#pragma once
#include "fastb_common.cuh"
#include "fft_core.cuh"

namespace fastb {

int bluestein_log2m(int n, int n_out);                       // smallest M = 2^l >= n + n_out - 1 (l >= 6)
size_t bluestein_table_bytes(int n, int n_out);
// tables for outputs k in [lo, lo + n_out) of an n-point transform, float64 arithmetic
int bluestein_prepare(int n, int n_out, int lo, void* tables, cudaStream_t st);
// complex weight copies of the K2 chirp-z kernel: wc = weight * c[col] * c[row], n_items stacked tables
int bluestein_prepare_weights(int n, int n_items, const float* weight, void* wc, cudaStream_t st);

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
