// K4a through the line FFT: the L full-size real layer screens of TEMPORAL mode
// (Fast.compute_phs_temporal, fast/fast.py:609-614; funcs.make_phase_fft(double=False),
// fast/funcs.py:210-223) as two passes of N-point line transforms instead of the O(N^3) direct DFT:
//   pass 0  row r' of layer l: noise x signed weight -> transform over c' -> T[l][c][r'] (transposed)
//   pass 1  column c: T[l][c][:] -> transform over r' -> screen[l][r][c] = (-1)^(r+c) Re(.)
// N = 64..2048 powers of two use the register radix FFT directly; any other even N <= 1024 goes through
// the chirp-z convolution (bluestein.cuh) with M >= 2N - 1.  Same RNG contract as the direct kernel
// (layer l = pair index FASTB_LAYER_PAIR_BASE + l).
#include "screen_detect_kernel.cuh"
#include "bluestein.cuh"

namespace fastb {
namespace {

template <int LOG2M, bool BLUE, int PASS, bool RNG, int THREADS>
__global__ void __launch_bounds__(THREADS) layer_lines_kernel(int N, unsigned long long seed,
                                                               const float* __restrict__ weight,
                                                               const float2* __restrict__ noise,
                                                               const float2* __restrict__ tables,
                                                               float2* __restrict__ T, float* __restrict__ out) {
    using F = LineFFT<LOG2M>;
    using Tw = typename F::Tw;
    constexpr int M = F::N, S1 = F::S1, LPB = THREADS / S1;
    static_assert(THREADS % S1 == 0 && LPB >= 1 && (S1 <= 32 || LPB <= 15), "line/barrier layout");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tw* twa = reinterpret_cast<Tw*>(smem_raw);
    Tw* twb = twa + F::kTwA;
    float2* bufs = reinterpret_cast<float2*>(twb + F::kTwB);
    float2* bhat = bufs + LPB * F::kBuf;                    // BLUE: M
    float2* chirp = bhat + (BLUE ? M : 0);                  // BLUE: N
    float2* rows = chirp + (BLUE ? N : 0);                  // PASS 0: LPB x N staged inputs

    const int tid = threadIdx.x, ln = tid / S1, u = tid % S1, l = blockIdx.y;
    const int line0 = blockIdx.x * LPB, line = line0 + ln;
    float2* buf = bufs + ln * F::kBuf;
    const LineSync<S1> sync{ln};
    for (int j = tid; j < F::kTwA + F::kTwB; j += THREADS) {
        const int ex = j < F::kTwA ? F::twa_exponent(j) : F::twb_exponent(j - F::kTwA);
        double s, c;
        sincospi(2.0 * (double)ex / (double)M, &s, &c);
        twa[j] = make_tw((float)c, (float)s, (Tw*)nullptr);
    }
    if (BLUE) {
        for (int j = tid; j < N; j += THREADS) chirp[j] = tables[j];
        for (int j = tid; j < M; j += THREADS) bhat[j] = tables[N + j];
    }
    const size_t layer = (size_t)l * N * N;
    if (PASS == 0) {
        __syncthreads();                                     // chirp visible
        const int nr = min(LPB, N - line0), S = (N + 15) / 16;
        if (RNG) {
            const unsigned long long g = FASTB_LAYER_PAIR_BASE + (unsigned long long)l;
            for (int idx = tid; idx < nr * S; idx += THREADS) {
                const int rl = idx / S, t = idx % S, r = line0 + rl;
                uint32_t mr[16], ma[16];
                noise_block_fields((uint32_t)(r * S + t), g, (uint32_t)seed, (uint32_t)(seed >> 32), mr, ma);
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int j = t + S * m;
                    if (j < N) {
                        const float2 x = weighted_normal_m(mr[m], ma[m], weight[layer + (size_t)r * N + j]);
                        rows[rl * N + j] = BLUE ? cmulf(x, chirp[j]) : x;
                    }
                }
            }
        } else {
            for (int idx = tid; idx < nr * N; idx += THREADS) {
                const int rl = idx / N, j = idx % N;
                const size_t o = layer + (size_t)(line0 + rl) * N + j;
                const float2 z = noise[o];
                const float w = weight[o];
                const float2 x = make_float2(z.x * w, z.y * w);
                rows[rl * N + j] = BLUE ? cmulf(x, chirp[j]) : x;
            }
        }
    }
    __syncthreads();

    float2 v[16];
    const bool live = line < N;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = u + S1 * m;
        float2 x = make_float2(0.f, 0.f);
        if (live && n < N) {
            if (PASS == 0) x = rows[ln * N + n];
            else {
                x = T[layer + (size_t)line * N + n];
                if (BLUE) x = cmulf(x, chirp[n]);
            }
        }
        v[m] = x;
    }
    if (BLUE) chirp_convolve<F>(u, v, twa, twb, buf, bhat, sync);
    else F::run(u, v, twa, twb, buf, sync);
    if (!live) return;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int k = F::k_base(u) + F::k_off(e);
        if (k >= N) continue;
        const float2 y = BLUE ? cmulf(v[e], chirp[k]) : v[e];
        if (PASS == 0) T[layer + (size_t)k * N + line] = y;                               // T[l][c = k][r' = line]
        else out[layer + (size_t)k * N + line] = ((k + line) & 1) ? -y.x : y.x;           // screen[l][r = k][c = line]
    }
}

template <int LOG2M, bool BLUE>
int run_layers(int n, int L, unsigned long long seed, const float* weight, const float2* noise, const float2* tables,
               float2* T, float* out, cudaStream_t st) {
    using F = LineFFT<LOG2M>;
    constexpr int THREADS = F::S1 >= 64 ? 256 : 128, LPB = THREADS / F::S1;
    const size_t common = sizeof(float2) * ((size_t)F::kTwA + F::kTwB + (size_t)LPB * F::kBuf + (BLUE ? (size_t)F::N + n : 0));
    const size_t smem0 = common + sizeof(float2) * (size_t)LPB * n;
    if (smem0 > 227 * 1024) {
        set_error("fastb_layer_screens: n=%d needs %zu B of shared memory", n, smem0);
        return FASTB_ERR_UNSUPPORTED;
    }
    dim3 grid((n + LPB - 1) / LPB, L);
    void (*k0)(int, unsigned long long, const float*, const float2*, const float2*, float2*, float*) =
        noise ? layer_lines_kernel<LOG2M, BLUE, 0, false, THREADS> : layer_lines_kernel<LOG2M, BLUE, 0, true, THREADS>;
    void (*k1)(int, unsigned long long, const float*, const float2*, const float2*, float2*, float*) =
        layer_lines_kernel<LOG2M, BLUE, 1, false, THREADS>;
    FASTB_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
    FASTB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)common));
    k0<<<grid, THREADS, smem0, st>>>(n, seed, weight, noise, tables, T, out);
    int rc = check_launch("layer_lines_kernel(rows)");
    if (rc) return rc;
    k1<<<grid, THREADS, common, st>>>(n, seed, weight, noise, tables, T, out);
    return check_launch("layer_lines_kernel(columns)");
}

bool pow2(int n) { return (n & (n - 1)) == 0; }

}  // namespace

bool layer_fft_ok(int n) {
    if (n < 4 || (n % 2)) return false;
    if (pow2(n)) return n >= 64 && n <= 2048;
    return 2 * n - 1 <= 2048;
}

size_t layer_fft_workspace_bytes(int n, int L) {
    size_t b = sizeof(float2) * (size_t)L * n * n;
    if (!pow2(n)) b += bluestein_table_bytes(n, n);
    return b;
}

int layer_screens_fft(int n, int L, unsigned long long seed, const float* weight, const float* noise, float* screens,
                      void* workspace, cudaStream_t st) {
    float2* T = (float2*)workspace;
    const float2* nz = (const float2*)noise;
    if (pow2(n)) {
        switch (n) {
            case 64: return run_layers<6, false>(n, L, seed, weight, nz, nullptr, T, screens, st);
            case 128: return run_layers<7, false>(n, L, seed, weight, nz, nullptr, T, screens, st);
            case 256: return run_layers<8, false>(n, L, seed, weight, nz, nullptr, T, screens, st);
            case 512: return run_layers<9, false>(n, L, seed, weight, nz, nullptr, T, screens, st);
            case 1024: return run_layers<10, false>(n, L, seed, weight, nz, nullptr, T, screens, st);
            default: return run_layers<11, false>(n, L, seed, weight, nz, nullptr, T, screens, st);
        }
    }
    float2* tables = T + (size_t)L * n * n;
    int rc = bluestein_prepare(n, n, 0, tables, st);
    if (rc) return rc;
    switch (bluestein_log2m(n, n)) {
        case 6: return run_layers<6, true>(n, L, seed, weight, nz, tables, T, screens, st);
        case 7: return run_layers<7, true>(n, L, seed, weight, nz, tables, T, screens, st);
        case 8: return run_layers<8, true>(n, L, seed, weight, nz, tables, T, screens, st);
        case 9: return run_layers<9, true>(n, L, seed, weight, nz, tables, T, screens, st);
        case 10: return run_layers<10, true>(n, L, seed, weight, nz, tables, T, screens, st);
        default: return run_layers<11, true>(n, L, seed, weight, nz, tables, T, screens, st);
    }
}

}  // namespace fastb
