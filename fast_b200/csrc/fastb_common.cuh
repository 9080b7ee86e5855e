// Shared helpers of libfastb (sm_100a): error plumbing, launch counting, Philox4x32-10 and
// Box-Muller.  See include/fastb.h for the contracts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/fastb.h"
#include "fft_core.cuh"

namespace fastb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);          // cudaGetLastError -> FASTB_* (+ counts the launch)
int fail_cuda(cudaError_t e, const char* what);

#define FASTB_REQUIRE(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            ::fastb::set_error(__VA_ARGS__);                      \
            return FASTB_ERR_ARG;                                 \
        }                                                         \
    } while (0)

#define FASTB_CUDA(call)                                          \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return ::fastb::fail_cuda(e__, #call); \
    } while (0)

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kStreamNoise = 0x5CE7E000u;
constexpr uint32_t kStreamChi = 0x10CA3900u;
constexpr uint32_t kStreamSubharm = 0x5AB4A200u;

// Philox4x32-R (Salmon et al. SC'11; R = 10 is the standard generator, R = 7 the smallest round
// count Random123 documents as passing BigCrush).  Key schedule is uniform across the warp.
template <int R>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}

// log2 of a normal (non-denormal) float as the bare MUFU.LG2: __log2f wraps the same instruction in
// a compare, a predicated 2^24 scaling and a predicated -24 correction for denormal arguments --
// three more issued instructions per call, and the Box-Muller argument is never below 2^-23.
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Box-Muller from two 23-bit mantissa fields, scaled by w: radius argument 1 - mr 2^-23 in (0, 1],
// angle 2 pi ma 2^-23.  The mantissa trick builds 1+u in [1, 2) without an int->float conversion;
// sin/cos are 2 pi-periodic so 2 pi (1+u) is used directly.
// 4 MUFU (lg2, sqrt, sin, cos) + 4 FMUL + 1 FMUL2 + 1 FADD + 2 LEA.HI per complex sample.
__device__ __forceinline__ float2 weighted_normal_m(uint32_t mr, uint32_t ma, float w) {
    const float u1 = 2.0f - __uint_as_float(0x3f800000u | mr);
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * lg2_ftz(u1)));
    rad *= w;
    const float ang = 6.283185307179586f * __uint_as_float(0x3f800000u | ma);
    float2 cs;
    __sincosf(ang, &cs.y, &cs.x);
    return mul2(cs, bc2(rad));            // (rad cos, rad sin): one FMUL2 with a broadcast operand
}

// The same sample for a weight that already carries the factor sqrt(2 ln 2) (kBoxMullerScale):
// radius = sqrt(-log2 u1) * ws, one multiply less per sample.
constexpr float kBoxMullerScale = 1.1774100225154747f;
__device__ __forceinline__ float2 weighted_normal_s(uint32_t mr, uint32_t ma, float ws) {
    const float u1 = 2.0f - __uint_as_float(0x3f800000u | mr);
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-lg2_ftz(u1)));
    rad *= ws;
    const float ang = 6.283185307179586f * __uint_as_float(0x3f800000u | ma);
    float2 cs;
    __sincosf(ang, &cs.y, &cs.x);
    return mul2(cs, bc2(rad));
}

// top 23 bits of each word (used by the chi and sub-harmonic streams)
__device__ __forceinline__ float2 box_muller(uint32_t wa, uint32_t wb) {
    return weighted_normal_m(wa >> 9, wb >> 9, 1.0f);
}

// ---- phase-noise stream ------------------------------------------------------------------
// Noise block b = r * S + t (S = ceil(N/16)) holds the 16 cells (r, t + S m), m < 16, of pair g.
// It is fed by SIX Philox calls q < 6 with counter (b, g lo, g hi, kStreamNoise + q): 24 words
// W[4q + j].  Each group of three words G < 8 (W[3G], W[3G+1], W[3G+2]) yields four 23-bit
// fields: the top 23 bits of each word, plus one field mixed from their low 9 bits, so that
// 24 words feed 32 uniforms = 16 Box-Muller pairs (8 calls would otherwise be needed):
//   cell m = 2G   : radius W[3G]   >> 9, angle W[3G+1] >> 9
//   cell m = 2G+1 : radius W[3G+2] >> 9, angle (W[3G]&511) << 14 | (W[3G+1]&511) << 5 | (W[3G+2] >> 4)&31
__device__ __forceinline__ void noise_block_fields(uint32_t block, unsigned long long g, uint32_t k0,
                                                   uint32_t k1, uint32_t (&mr)[16], uint32_t (&ma)[16]) {
    uint32_t W[24];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const uint4 w = philox4x32_10(block, (uint32_t)g, (uint32_t)(g >> 32), kStreamNoise + q, k0, k1);
        W[4 * q] = w.x;
        W[4 * q + 1] = w.y;
        W[4 * q + 2] = w.z;
        W[4 * q + 3] = w.w;
    }
#pragma unroll
    for (int G = 0; G < 8; ++G) {
        const uint32_t a = W[3 * G], b = W[3 * G + 1], c = W[3 * G + 2];
        mr[2 * G] = a >> 9;
        ma[2 * G] = b >> 9;
        mr[2 * G + 1] = c >> 9;
        ma[2 * G + 1] = ((a & 0x1FFu) << 14) | ((b & 0x1FFu) << 5) | ((c >> 4) & 0x1Fu);
    }
}

// ---- 'device-fast' phase-noise stream (opt-in, RNG='device-fast') ----------------------------
// Same block / cell mapping, fed by FIVE Philox4x32-7 calls q < 5 with counter
// (b, g lo, g hi, kStreamNoiseFast + q): 20 words W[4q + j].  Cell m < 16 owns word W[m] and byte
// m % 4 of the extra word X = W[16 + m / 4]  (40 bits per complex sample instead of 46):
//   radius field = W[m] & 0x7FFFFF                                        (23 bits)
//   angle  field = ((W[m] >> 9) & 0x7FC000) ^ (byte << 8)                  (bits 22..8: 15 bits)
// 35 Philox rounds per 16 samples instead of 60.
constexpr uint32_t kStreamNoiseFast = 0x5CE7F000u;
__device__ __forceinline__ void noise_block_fields_fast(uint32_t block, unsigned long long g, uint32_t k0,
                                                        uint32_t k1, uint32_t (&mr)[16], uint32_t (&ma)[16]) {
    uint32_t W[20];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const uint4 w = philox4x32<7>(block, (uint32_t)g, (uint32_t)(g >> 32), kStreamNoiseFast + q, k0, k1);
        W[4 * q] = w.x;
        W[4 * q + 1] = w.y;
        W[4 * q + 2] = w.z;
        W[4 * q + 3] = w.w;
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const uint32_t w = W[m];
        mr[m] = w & 0x7FFFFFu;
        // byte m % 4 of the extra word moved to bits 15..8 (one PRMT), xor-ed under the word's top 9 bits
        ma[m] = ((w >> 9) & 0x7FC000u) ^ __byte_perm(W[16 + m / 4], 0u, 0x4404u | ((m % 4) << 4));
    }
}

// standard normal n_i used for the log-amplitude of global realisation index i
__device__ __forceinline__ float chi_normal(uint64_t seed, uint64_t i) {
    const uint64_t call = i >> 2;
    const uint4 w = philox4x32_10((uint32_t)call, (uint32_t)(call >> 32), 0u, kStreamChi,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
    const int sel = (int)(i & 3);
    const float2 n = (sel < 2) ? box_muller(w.x, w.y) : box_muller(w.z, w.w);
    return (sel & 1) ? n.y : n.x;
}

}  // namespace fastb
