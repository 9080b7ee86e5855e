// Register-resident line FFT for N = 2^LOG2N (64..2048), 16 complex elements per thread,
// N/16 threads per line.  Inverse sign convention: X[k] = sum_n x[n] exp(+2 pi i n k / N).
//
// Decomposition (decimation in frequency), S1 = N/16 threads per line:
//   phase A  thread t holds x[t + S1 m], m = 0..15: 16-point DFT over m -> a = k mod 16,
//            twiddle by w_N^(t a), exchange through the line buffer.
//   N >= 256 (S1 = 16 S2): thread u = a S2 + t2 gathers t = t2 + S2 m2 and does a second
//   phase B  16-point DFT over m2 -> a2, twiddle by w_S1^(t2 a2).  S2 == 1: done,
//            k = a + 16 a2.  Otherwise a last exchange and
//   phase C  16/S2 DFTs of length S2 over t2 -> b2,  k = a + 16 (a2 + 16 b2).
//   N = 64, 128 (S1 = 4, 8): phase C directly after phase A with length S1 over t,
//            k = a + 16 b.
// The functions are __host__ __device__ so that tests/host_fft_emul.cu can run the exact
// index algebra on the CPU (threads emulated sequentially between the sync points).
#pragma once
#include <cuda_runtime.h>

#ifndef FASTB_HD
#define FASTB_HD __host__ __device__ __forceinline__
#endif

namespace fastb {

FASTB_HD float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// Complex add / subtract.  On sm_100a the (re, im) pair is handled by ONE packed FP32
// instruction (add.f32x2 / sub.f32x2 -> SASS FADD2): the kernel is issue-bound, and complex
// adds are ~30 % of its instructions.  The host build (CPU emulation test) uses scalar code.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(FASTB_NO_F32X2)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
#else
FASTB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FASTB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
FASTB_HD float2 cmuli(float2 a) { return make_float2(-a.y, a.x); }          // * (+i)

// inverse radix-4: y_k = sum_n x_n i^(n k)
FASTB_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 s02 = cadd(x0, x2), d02 = csub(x0, x2);
    const float2 s13 = cadd(x1, x3), d13 = cmuli(csub(x1, x3));
    x0 = cadd(s02, s13);
    x2 = csub(s02, s13);
    x1 = cadd(d02, d13);
    x3 = csub(d02, d13);
}

FASTB_HD void dft2(float2& x0, float2& x1) {
    const float2 s = cadd(x0, x1), d = csub(x0, x1);
    x0 = s;
    x1 = d;
}

#define FASTB_C8 0.70710678118654752f
#define FASTB_C16 0.92387953251128674f
#define FASTB_S16 0.38268343236508977f

// inverse 8-point DFT, natural order in and out (n = n1 + 2 n2, k = k2 + 4 k1)
FASTB_HD void dft8(float2 (&v)[8]) {
    dft4(v[0], v[2], v[4], v[6]);        // n1 = 0 : y[0][k2] in v[2 k2]
    dft4(v[1], v[3], v[5], v[7]);        // n1 = 1 : y[1][k2] in v[2 k2 + 1]
    // twiddle y[1][k2] *= w8^k2
    v[3] = cmul(v[3], make_float2(FASTB_C8, FASTB_C8));
    v[5] = cmuli(v[5]);
    v[7] = cmul(v[7], make_float2(-FASTB_C8, FASTB_C8));
    float2 o[8];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        o[k2] = cadd(v[2 * k2], v[2 * k2 + 1]);
        o[k2 + 4] = csub(v[2 * k2], v[2 * k2 + 1]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = o[i];
}

// inverse 16-point DFT, natural order in and out (n = n1 + 4 n2, k = k2 + 4 k1)
FASTB_HD void dft16(float2 (&v)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);  // -> v[n1+4k2]
    // twiddles w16^(n1 k2)
    v[5] = cmul(v[5], make_float2(FASTB_C16, FASTB_S16));      // (1,1) w^1
    v[9] = cmul(v[9], make_float2(FASTB_C8, FASTB_C8));        // (1,2) w^2
    v[13] = cmul(v[13], make_float2(FASTB_S16, FASTB_C16));    // (1,3) w^3
    v[6] = cmul(v[6], make_float2(FASTB_C8, FASTB_C8));        // (2,1) w^2
    v[10] = cmuli(v[10]);                                      // (2,2) w^4 = i
    v[14] = cmul(v[14], make_float2(-FASTB_C8, FASTB_C8));     // (2,3) w^6
    v[7] = cmul(v[7], make_float2(FASTB_S16, FASTB_C16));      // (3,1) w^3
    v[11] = cmul(v[11], make_float2(-FASTB_C8, FASTB_C8));     // (3,2) w^6
    v[15] = cmul(v[15], make_float2(-FASTB_C16, -FASTB_S16));  // (3,3) w^9
    float2 o[16];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        float2 a = v[4 * k2], b = v[4 * k2 + 1], c = v[4 * k2 + 2], d = v[4 * k2 + 3];
        dft4(a, b, c, d);                                       // over n1 -> k1
        o[k2] = a;
        o[k2 + 4] = b;
        o[k2 + 8] = c;
        o[k2 + 12] = d;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// inverse 32-point DFT, natural order in and out: two 16-point DFTs of the even / odd samples
// combined by one radix-2 stage, X[k] = E[k] + w32^k O[k], X[k+16] = E[k] - w32^k O[k]
FASTB_HD void dft32(float2 (&v)[32]) {
    constexpr float c[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                             0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f,
                             0.19509032201612825f, 0.f, -0.19509032201612825f, -0.38268343236508977f,
                             -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                             -0.92387953251128674f, -0.98078528040323043f};
    constexpr float sn[16] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                              0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f,
                              0.98078528040323043f, 1.f, 0.98078528040323043f, 0.92387953251128674f,
                              0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                              0.38268343236508977f, 0.19509032201612825f};
    float2 ev[16], od[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        ev[j] = v[2 * j];
        od[j] = v[2 * j + 1];
    }
    dft16(ev);
    dft16(od);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float2 t = (k == 0) ? od[0] : (k == 8) ? cmuli(od[8]) : cmul(od[k], make_float2(c[k], sn[k]));
        v[k] = cadd(ev[k], t);
        v[k + 16] = csub(ev[k], t);
    }
}

// Line FFT with 32 complex elements per thread for N = 512 (32 x 16) and N = 1024 (32 x 32):
// S1 = N/32 threads per line, ONE shared-memory exchange per line.
//   phase A  thread t holds x[t + S1 m], m < 32: 32-point DFT over m -> a = k mod 32, twiddle
//            w_N^(t a), store buf[a][t]
//   phase B  N = 1024: thread u = a gathers t < 32 and does a 32-point DFT -> b, k = a + 32 b
//            N = 512 : thread u handles a = u and a = u + 16, a 16-point DFT each, k = a + 32 b
template <int LOG2N>
struct LineFFT32 {
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = 32;                            // elements per thread
    static constexpr int S1 = N / 32;                       // threads per line (16 or 32)
    static_assert(LOG2N == 9 || LOG2N == 10, "LineFFT32 serves N = 512 and 1024");
    static constexpr int kRowA = S1 + 2;                    // buf[a * kRowA + t]: 16-byte aligned rows,
    static constexpr int kBuf = 32 * kRowA;                 // conflict-free 128-bit gathers
    static constexpr int kTwRow = 34;                       // twa[t * 34 + a], a < 32
    static constexpr int kTwA = kTwRow * S1;
    static constexpr int kTwB = 0;
    FASTB_HD static int n_in(int t, int m) { return t + S1 * m; }
    FASTB_HD static int twa_exponent(int idx) {
        const int t = idx / kTwRow, a = idx % kTwRow;
        return a < 32 ? (t * a) & (N - 1) : 0;
    }
    FASTB_HD static int twb_exponent(int) { return 0; }

    FASTB_HD static void phase_a(int t, float2 (&v)[32], const float2* twa, float2* buf) {
        dft32(v);
        const float4* q = reinterpret_cast<const float4*>(twa + t * kTwRow);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 w = q[j];
            if (j > 0) v[2 * j] = cmul(v[2 * j], make_float2(w.x, w.y));
            v[2 * j + 1] = cmul(v[2 * j + 1], make_float2(w.z, w.w));
        }
#pragma unroll
        for (int a = 0; a < 32; ++a) buf[a * kRowA + t] = v[a];
    }

    FASTB_HD static void phase_b(int u, float2 (&v)[32], const float2* buf) {
        if (S1 == 32) {
            const float4* q = reinterpret_cast<const float4*>(buf + u * kRowA);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 w = q[j];
                v[2 * j] = make_float2(w.x, w.y);
                v[2 * j + 1] = make_float2(w.z, w.w);
            }
            dft32(v);
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float2 x[16];
                const float4* q = reinterpret_cast<const float4*>(buf + (u + 16 * h) * kRowA);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w = q[j];
                    x[2 * j] = make_float2(w.x, w.y);
                    x[2 * j + 1] = make_float2(w.z, w.w);
                }
                dft16(x);
#pragma unroll
                for (int b = 0; b < 16; ++b) v[16 * h + b] = x[b];
            }
        }
    }

    // the whole line FFT; `sync` synchronises the S1 threads of the line
    template <typename Sync>
    FASTB_HD static void run(int u, float2 (&v)[32], const float2* twa, const float2*, float2* buf, Sync sync) {
        phase_a(u, v, twa, buf);
        sync();
        phase_b(u, v, buf);
        sync();                 // buffer may be rewritten by the next line
    }

    FASTB_HD static int k_base(int u) { return u; }
    FASTB_HD static constexpr int k_off(int e) { return S1 == 32 ? 32 * e : 16 * (e / 16) + 32 * (e % 16); }
    FASTB_HD static int k_out(int u, int e) { return k_base(u) + k_off(e); }
    FASTB_HD static constexpr bool k_off_all_even() {
        for (int e = 0; e < 32; ++e)
            if (k_off(e) & 1) return false;
        return true;
    }
};

template <int LOG2N>
struct LineFFT {
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = 16;                            // elements per thread
    static constexpr int S1 = N / 16;                       // threads per line
    static constexpr bool kThree = (S1 >= 16);              // N >= 256
    static constexpr int S2 = kThree ? S1 / 16 : 1;
    static constexpr int SF = kThree ? S2 : S1;             // length of the last small DFTs
    static constexpr bool kHasC = (SF > 1);
    // exchange-A padding: S2 == 1 uses 2 so that each thread's 16 gathered values are 16-byte
    // aligned (128-bit shared loads, conflict-free at a 144-byte lane stride)
    static constexpr int kPadA = kThree ? (S2 == 1 ? 2 : S2) : 0;
    static constexpr int kBufA = 16 * (S1 + kPadA);         // exchange A layout: a*(S1+pad)+t
    static constexpr int kBufC = 17 * S1;                   // exchange C layout: u*17 + e
    static constexpr int kBuf = (kBufA > kBufC) ? kBufA : kBufC;   // float2 per line
    static_assert(LOG2N >= 6 && LOG2N <= 11, "N must be 64..2048");

    // element index held in register m of thread t before phase A
    FASTB_HD static int n_in(int t, int m) { return t + S1 * m; }

    // Twiddle tables, one row of kTwRow = 18 float2 per thread (16 used): the 144-byte row
    // stride makes 128-bit shared loads of consecutive threads conflict-free.
    //   twa[t * 18 + a]   = exp(+2 pi i (t a) / N)          t < S1, a < 16
    //   twb[t2 * 18 + a2] = exp(+2 pi i (16 t2 a2) / N)     t2 < S2, a2 < 16   (S2 > 1 only)
    static constexpr int kTwRow = 18;
    static constexpr int kTwA = kTwRow * S1;
    static constexpr int kTwB = (kThree && S2 > 1) ? kTwRow * S2 : 0;
    FASTB_HD static int twa_exponent(int idx) {
        const int t = idx / kTwRow, a = idx % kTwRow;
        return a < 16 ? (t * a) & (N - 1) : 0;
    }
    FASTB_HD static int twb_exponent(int idx) {
        const int t2 = idx / kTwRow, a2 = idx % kTwRow;
        return a2 < 16 ? (16 * t2 * a2) & (N - 1) : 0;
    }

    // multiply v[1..15] by the 16 twiddles of one table row, fetched as 8 x 128-bit loads
    FASTB_HD static void apply_twiddle_row(float2 (&v)[16], const float2* row) {
        const float4* q = reinterpret_cast<const float4*>(row);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 w = q[j];
            if (j > 0) v[2 * j] = cmul(v[2 * j], make_float2(w.x, w.y));
            v[2 * j + 1] = cmul(v[2 * j + 1], make_float2(w.z, w.w));
        }
    }

    // phase A: dft16 over m, twiddle, write exchange buffer
    FASTB_HD static void phase_a(int t, float2 (&v)[16], const float2* twa, float2* buf) {
        dft16(v);
        apply_twiddle_row(v, twa + t * kTwRow);
        if (kThree) {
#pragma unroll
            for (int a = 0; a < 16; ++a) buf[a * (S1 + kPadA) + t] = v[a];
        } else {
#pragma unroll
            for (int a = 0; a < 16; ++a) buf[t * 17 + a] = v[a];
        }
    }

    // phase B (N >= 256): gather, dft16 over m2, twiddle.  If S2 > 1 the caller must sync
    // and then call phase_b_store before phase C.
    FASTB_HD static void phase_b(int u, float2 (&v)[16], const float2* twb, const float2* buf) {
        const int a = u / S2, t2 = u % S2;
        if (S2 == 1) {
            // 16 contiguous, 16-byte aligned values: 8 x 128-bit loads
            const float4* q = reinterpret_cast<const float4*>(buf + a * (S1 + kPadA));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 w = q[j];
                v[2 * j] = make_float2(w.x, w.y);
                v[2 * j + 1] = make_float2(w.z, w.w);
            }
        } else {
#pragma unroll
            for (int m2 = 0; m2 < 16; ++m2) v[m2] = buf[a * (S1 + kPadA) + t2 + S2 * m2];
        }
        dft16(v);
        if (S2 > 1) apply_twiddle_row(v, twb + t2 * kTwRow);
    }

    FASTB_HD static void phase_b_store(int u, const float2 (&v)[16], float2* buf) {
#pragma unroll
        for (int e = 0; e < 16; ++e) buf[u * 17 + e] = v[e];
    }

    // phase C: 16/SF DFTs of length SF across the SF threads of a group
    FASTB_HD static void phase_c(int u, float2 (&v)[16], const float2* buf) {
        constexpr int G = 16 / SF;
        const int g0 = (u / SF) * SF, j = u % SF;
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int t2 = 0; t2 < SF; ++t2) v[i * SF + t2] = buf[(g0 + t2) * 17 + i * SF + j];
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            if (SF == 2) {
                dft2(v[i * SF], v[i * SF + 1]);
            } else if (SF == 4) {
                dft4(v[i * SF], v[i * SF + 1], v[i * SF + 2], v[i * SF + 3]);
            } else if (SF == 8) {
                float2 w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) w[q] = v[i * SF + q];
                dft8(w);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[i * SF + q] = w[q];
            }
        }
    }

    // the whole line FFT; `sync` synchronises the S1 threads of the line
    template <typename Sync>
    FASTB_HD static void run(int u, float2 (&v)[16], const float2* twa, const float2* twb, float2* buf, Sync sync) {
        phase_a(u, v, twa, buf);
        sync();
        if (kThree) {
            phase_b(u, v, twb, buf);
            if (S2 > 1) {
                sync();
                phase_b_store(u, v, buf);
                sync();
                phase_c(u, v, buf);
            }
        } else {
            phase_c(u, v, buf);
        }
        sync();                 // buffer may be rewritten by the next line
    }

    // Output index k held in register e of thread u after the last phase, split into a
    // per-thread base and a compile-time offset: k = k_base(u) + k_off(e).  Offsets are even.
    FASTB_HD static int k_base(int u) {
        if (kThree) return (S2 == 1) ? u : (u / S2) + 16 * (u % S2);
        return u % SF;
    }
    FASTB_HD static constexpr int k_off(int e) {
        return kThree ? ((S2 == 1) ? 16 * e : 16 * S2 * (e / S2) + 256 * (e % S2))
                      : SF * (e / SF) + 16 * (e % SF);
    }
    FASTB_HD static int k_out(int u, int e) { return k_base(u) + k_off(e); }
    FASTB_HD static constexpr bool k_off_all_even() {
        for (int e = 0; e < 16; ++e)
            if (k_off(e) & 1) return false;
        return true;
    }
};

}  // namespace fastb
