// K2 for grids that are not a power of two (any even N with N + n_pup - 1 <= 2048): the reference's
// NPXLS 'auto' rule (fast/fast.py:166-187) produces sizes such as 164.  Geometry, table sizes and the
// dispatch to the per-length units (screen_detect_bluestein_m.cu, one per M = 2^6 .. 2^11) live here,
// together with the natural-order chirp tables of the TEMPORAL layer screens (layer_screens_fft.cu).
//   X[k] = sum_n x[n] e^{+2 pi i n k / N} = c[k] sum_n (x[n] c[n]) conj(c[k - n]),   c[m] = e^{i pi m^2 / N}
// Only the n_pup outputs k in [lo, lo + n_pup) are wanted, so the kernel b[d] = conj(c[d]) is needed on
// d = k - n in (lo - N, lo + n_pup): N + n_pup - 1 consecutive values, alias-free on a circle of length M.
#include "screen_detect_kernel.cuh"
#include "bluestein.cuh"

namespace fastb {
namespace {

int blue_log2m(int n, int n_pup) {
    int l = 6;
    while ((1 << l) < n + n_pup - 1) ++l;
    return l;
}

// layer-screen tables: chirp[N] then bhat[M] (float2 each), natural order
__global__ void __launch_bounds__(256) bluestein_tables_kernel(int N, int M, int lo, int P, float2* chirp,
                                                               float2* bhat) {
    __shared__ double red[2][8];
    const int q = blockIdx.x;
    if (q >= M) {                                  // the last blocks write the chirp
        for (int n = (q - M) * 256 + threadIdx.x; n < N; n += 256 * (gridDim.x - M)) {
            double s, c;
            sincospi((double)(((long long)n * n) % (2LL * N)) / (double)N, &s, &c);
            chirp[n] = make_float2((float)c, (float)s);
        }
        return;
    }
    // bhat[q] = (1/M) sum_j b[j] e^{+2 pi i j q / M},  b[j] = conj(c[d]) for the d in (lo - N, lo + P)
    // congruent to j mod M, 0 when there is none
    double sr = 0.0, si = 0.0;
    const int d_lo = lo - N + 1, d_hi = lo + P - 1;
    for (int d = d_lo + threadIdx.x; d <= d_hi; d += 256) {
        const int j = ((d % M) + M) % M;
        const long long d2 = ((long long)d * d) % (2LL * N);
        // phase / pi = -d^2 / N + 2 j q / M
        const double ph = -(double)d2 / (double)N + 2.0 * (double)(((long long)j * q) % M) / (double)M;
        double s, c;
        sincospi(ph, &s, &c);
        sr += c;
        si += s;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sr;
        red[1][threadIdx.x >> 5] = si;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, ti = 0.0;
        for (int w = 0; w < 8; ++w) {
            tr += red[0][w];
            ti += red[1][w];
        }
        bhat[q] = make_float2((float)(tr / M), (float)(ti / M));
    }
}

}  // namespace

#define FASTB_DECL_BLUE(k)                                                                                          \
    int prepare_blue_##k(int n, int n_pup, int lo, int C, int n_items, const float* weight, void* wcq, void* tables, \
                         cudaStream_t st);                                                                          \
    int launch_blue_##k(const RunArgs& a, const RadixRequest& rq, const void* tables, int C, cudaStream_t st);
FASTB_DECL_BLUE(6) FASTB_DECL_BLUE(7) FASTB_DECL_BLUE(8) FASTB_DECL_BLUE(9) FASTB_DECL_BLUE(10) FASTB_DECL_BLUE(11)
#undef FASTB_DECL_BLUE

int bluestein_log2m(int n, int n_pup) { return blue_log2m(n, n_pup); }

BlueGeom blue_geom(int n, int n_pup) {
    BlueGeom g;
    g.log2m = blue_log2m(n, n_pup);
    g.M = 1 << g.log2m;
    g.S1 = g.M / 16;
    g.C = blue_cell_pairs(n, g.S1);
    return g;
}

// the chirp-z path serves every even grid the radix kernels do not (powers of two 64..2048 are theirs: the noise
// stride of the device RNG follows the kernel that owns the size, include/fastb.h)
bool bluestein_ok(int n, int n_pup) {
    const bool radix = n >= 64 && n <= 2048 && (n & (n - 1)) == 0;
    return !radix && n >= 4 && (n % 2) == 0 && n + n_pup - 1 <= 2048;
}

int bluestein_ctas_per_sm(int n, int n_pup) { return blue_log2m(n, n_pup) <= 8 ? 4 : 2; }

size_t bluestein_table_bytes(int n, int n_pup) {
    return sizeof(float2) * ((size_t)n + ((size_t)1 << blue_log2m(n, n_pup)));
}

size_t bluestein_weight_bytes(int n, int n_pup, int n_items) {
    const BlueGeom g = blue_geom(n, n_pup);
    return sizeof(float4) * (size_t)n * g.C * g.S1 * (n_items > 1 ? n_items : 1);
}

size_t bluestein_k2_table_bytes(int n, int n_pup) {
    const BlueGeom g = blue_geom(n, n_pup);
    return sizeof(float2) * ((size_t)18 * g.S1 + n_pup);
}

int bluestein_prepare(int n, int n_pup, int lo, void* tables, cudaStream_t st) {
    const int M = 1 << blue_log2m(n, n_pup);
    float2* t = (float2*)tables;
    bluestein_tables_kernel<<<M + 4, 256, 0, st>>>(n, M, lo, n_pup, t, t + n);
    return check_launch("bluestein_tables_kernel");
}

// K2: weight == NULL leaves the weight copies alone (tables only)
int bluestein_prepare_k2(int n, int n_pup, int lo, int n_items, const float* weight, void* wcq, void* tables,
                         cudaStream_t st) {
    const BlueGeom g = blue_geom(n, n_pup);
    switch (g.log2m) {
        case 6: return prepare_blue_6(n, n_pup, lo, g.C, n_items, weight, wcq, tables, st);
        case 7: return prepare_blue_7(n, n_pup, lo, g.C, n_items, weight, wcq, tables, st);
        case 8: return prepare_blue_8(n, n_pup, lo, g.C, n_items, weight, wcq, tables, st);
        case 9: return prepare_blue_9(n, n_pup, lo, g.C, n_items, weight, wcq, tables, st);
        case 10: return prepare_blue_10(n, n_pup, lo, g.C, n_items, weight, wcq, tables, st);
        case 11: return prepare_blue_11(n, n_pup, lo, g.C, n_items, weight, wcq, tables, st);
        default: break;
    }
    set_error("screen_detect_bluestein: N=%d, n_pup=%d out of range", n, n_pup);
    return FASTB_ERR_UNSUPPORTED;
}

int launch_bluestein(const RunArgs& a, const RadixRequest& rq, const void* tables, cudaStream_t st) {
    const BlueGeom g = blue_geom(a.n, a.n_pup);
    switch (g.log2m) {
        case 6: return launch_blue_6(a, rq, tables, g.C, st);
        case 7: return launch_blue_7(a, rq, tables, g.C, st);
        case 8: return launch_blue_8(a, rq, tables, g.C, st);
        case 9: return launch_blue_9(a, rq, tables, g.C, st);
        case 10: return launch_blue_10(a, rq, tables, g.C, st);
        case 11: return launch_blue_11(a, rq, tables, g.C, st);
        default: break;
    }
    set_error("screen_detect_bluestein: N=%d, n_pup=%d out of range", a.n, a.n_pup);
    return FASTB_ERR_UNSUPPORTED;
}

}  // namespace fastb
