// K3: per-rank result statistics (moments, extrema, dB histogram) for the multi-GPU reduction.
// Contract: include/fastb.h.
#include "fastb_common.cuh"

namespace fastb {
namespace {

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

__global__ void __launch_bounds__(256) stats_kernel(const float* __restrict__ r, long long n, double db_lo,
                                                    double db_hi, int nbins, double* sums, double* minmax,
                                                    unsigned long long* hist) {
    __shared__ double red[6][8];
    double s[5] = {0, 0, 0, 0, 0};   // n, sum r, sum r^2, sum dB, sum dB^2
    double nonpos = 0;
    double mn = INFINITY, mx = -INFINITY;
    const double inv_w = (double)nbins / (db_hi - db_lo);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)r[i];
        s[0] += 1.0;
        s[1] += v;
        s[2] += v * v;
        mn = fmin(mn, v);
        mx = fmax(mx, v);
        if (v > 0.0) {
            const double db = 10.0 * log10(v);
            s[3] += db;
            s[4] += db * db;
            int b;
            if (db < db_lo) b = nbins;
            else if (db >= db_hi) b = nbins + 1;
            else {
                b = (int)((db - db_lo) * inv_w);
                if (b >= nbins) b = nbins - 1;
            }
            atomicAdd(&hist[b], 1ULL);
        } else {
            nonpos += 1.0;
            atomicAdd(&hist[nbins], 1ULL);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double vals[6] = {s[0], s[1], s[2], s[3], s[4], nonpos};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
        if (lane == 0) red[k][warp] = vals[k];
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        atomic_min_double(&minmax[0], mn);
        atomic_max_double(&minmax[1], mx);
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

}  // namespace
}  // namespace fastb

using namespace fastb;

extern "C" int fastb_stats(const float* d_r, int64_t n, double db_lo, double db_hi, int32_t nbins,
                           double* d_sums, double* d_minmax, unsigned long long* d_hist, void* stream) {
    FASTB_REQUIRE(d_r && d_sums && d_minmax && d_hist, "fastb_stats: NULL pointer");
    FASTB_REQUIRE(n >= 0 && nbins >= 1 && db_hi > db_lo, "fastb_stats: bad range");
    if (n == 0) return FASTB_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_r, n, db_lo, db_hi, nbins, d_sums,
                                                                   d_minmax, d_hist);
    return check_launch("stats_kernel");
}
