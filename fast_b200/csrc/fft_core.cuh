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
//
// The code is generic over the VALUE TYPE V a thread operates on:
//   float2  one line, (re, im) interleaved.  Complex adds are one packed FADD2 (add.f32x2).
//   pc      a PAIR of lines held planar: re = (line A, line B), im = (line A, line B).  Every
//           operation -- adds, twiddle multiplies (2 FMUL2 + 2 FFMA2 for both lines, broadcast
//           immediates for the constant twiddles), rotations by i (free) -- is packed FP32, which
//           halves the multiply instruction count per line; the kernel is issue-bound.
// The functions are __host__ __device__ so that tests/host/host_fft_emul.cu can run the exact
// index algebra on the CPU (threads emulated sequentially between the sync points; the packed
// intrinsics fall back to scalar code on the host).
#pragma once
#include <cuda_runtime.h>

#ifndef FASTB_HD
#define FASTB_HD __host__ __device__ __forceinline__
#endif

namespace fastb {

// ---- packed FP32 pairs (sm_100: FADD2 / FMUL2 / FFMA2) ------------------------------------
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
#define FASTB_PACKED 1
#else
#define FASTB_PACKED 0
#endif
FASTB_HD float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
FASTB_HD float2 bc2(float s) { return make_float2(s, s); }
FASTB_HD float2 add2(float2 a, float2 b) {
#if FASTB_PACKED
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
FASTB_HD float2 sub2(float2 a, float2 b) {
#if FASTB_PACKED
    return __fadd2_rn(a, neg2(b));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
FASTB_HD float2 mul2(float2 a, float2 b) {
#if FASTB_PACKED
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
FASTB_HD float2 fma2(float2 a, float2 b, float2 c) {           // a b + c
#if FASTB_PACKED
    return __ffma2_rn(a, b, c);
#else
    return make_float2(a.x * b.x + c.x, a.y * b.y + c.y);
#endif
}
FASTB_HD float2 fnma2(float2 a, float2 b, float2 c) {          // c - a b
#if FASTB_PACKED
    return __ffma2_rn(neg2(a), b, c);
#else
    return make_float2(c.x - a.x * b.x, c.y - a.y * b.y);
#endif
}

// ---- value types and their complex arithmetic ----------------------------------------------
struct pc {            // two lines, planar: .x = line A, .y = line B
    float2 re, im;
};
template <class V>
struct TwOf;
template <>
struct TwOf<float2> {
    using type = float2;                   // (cos, sin)
    static constexpr int kLines = 1;
    static constexpr int kTwPerRow = 18;   // 16 used; 144-byte rows: conflict-free 128-bit loads
};
template <>
struct TwOf<pc> {
    using type = float4;                   // (cos, cos, sin, sin)
    static constexpr int kLines = 2;
    static constexpr int kTwPerRow = 17;   // 16 used; 272-byte rows: conflict-free 128-bit loads
};
FASTB_HD float2 make_tw(float c, float s, float2*) { return make_float2(c, s); }
FASTB_HD float4 make_tw(float c, float s, float4*) { return make_float4(c, c, s, s); }

// one line, (re, im) interleaved.  sm_100 packed operands take a free half swap (.LO_HI), a scalar
// broadcast (.F32) and a per-half negation (.NP), so a +- i b is ONE FADD2 and a complex
// multiply is FMUL2 + FFMA2 (2 instructions instead of 4 scalar ones).
FASTB_HD float2 cadd(float2 a, float2 b) { return add2(a, b); }
FASTB_HD float2 csub(float2 a, float2 b) { return sub2(a, b); }
FASTB_HD float2 caddi(float2 a, float2 b) { return add2(a, make_float2(-b.y, b.x)); }     // a + i b
FASTB_HD float2 csubi(float2 a, float2 b) { return add2(a, make_float2(b.y, -b.x)); }     // a - i b
FASTB_HD float2 cmuli(float2 a) { return make_float2(-a.y, a.x); }                         // i a
FASTB_HD float2 cmul(float2 a, float2 w) {                                                 // w = (cos, sin)
    return fma2(make_float2(-a.y, a.x), bc2(w.y), mul2(a, bc2(w.x)));
}
FASTB_HD float2 cmulc(float2 a, float c, float s) { return cmul(a, make_float2(c, s)); }
// two lines
FASTB_HD pc cadd(pc a, pc b) { return pc{add2(a.re, b.re), add2(a.im, b.im)}; }
FASTB_HD pc csub(pc a, pc b) { return pc{sub2(a.re, b.re), sub2(a.im, b.im)}; }
FASTB_HD pc caddi(pc a, pc b) { return pc{sub2(a.re, b.im), add2(a.im, b.re)}; }
FASTB_HD pc csubi(pc a, pc b) { return pc{add2(a.re, b.im), sub2(a.im, b.re)}; }
FASTB_HD pc cmuli(pc a) { return pc{neg2(a.im), a.re}; }
FASTB_HD pc cmul(pc a, float4 w) {
    const float2 cc = make_float2(w.x, w.y), ss = make_float2(w.z, w.w);
    return pc{fnma2(a.im, ss, mul2(a.re, cc)), fma2(a.re, ss, mul2(a.im, cc))};
}
FASTB_HD pc cmulc(pc a, float c, float s) {
    return pc{fnma2(a.im, bc2(s), mul2(a.re, bc2(c))), fma2(a.re, bc2(s), mul2(a.im, bc2(c)))};
}

// one line, scalar FP32 (tuning flavour FASTB_SCALAR_STAGES): the same arithmetic as float2 but as
// FADD / FMUL / FFMA, which may issue on either FMA pipe -- packed instructions occupy the heavy pipe alone
struct sc {
    float x, y;
};
FASTB_HD sc cadd(sc a, sc b) { return sc{a.x + b.x, a.y + b.y}; }
FASTB_HD sc csub(sc a, sc b) { return sc{a.x - b.x, a.y - b.y}; }
FASTB_HD sc caddi(sc a, sc b) { return sc{a.x - b.y, a.y + b.x}; }
FASTB_HD sc csubi(sc a, sc b) { return sc{a.x + b.y, a.y - b.x}; }
FASTB_HD sc cmuli(sc a) { return sc{-a.y, a.x}; }
FASTB_HD sc cmulc(sc a, float c, float s) { return sc{fmaf(-a.y, s, a.x * c), fmaf(a.x, s, a.y * c)}; }

// ---- small DFTs (inverse sign), natural order in and out -------------------------------------
template <class V>
FASTB_HD void dft4(V& x0, V& x1, V& x2, V& x3) {            // y_k = sum_n x_n i^(n k)
    const V s02 = cadd(x0, x2), d02 = csub(x0, x2);
    const V s13 = cadd(x1, x3), e13 = csub(x1, x3);
    x0 = cadd(s02, s13);
    x2 = csub(s02, s13);
    x1 = caddi(d02, e13);
    x3 = csubi(d02, e13);
}

template <class V>
FASTB_HD void dft2(V& x0, V& x1) {
    const V s = cadd(x0, x1), d = csub(x0, x1);
    x0 = s;
    x1 = d;
}

#define FASTB_C8 0.70710678118654752f
#define FASTB_C16 0.92387953251128674f
#define FASTB_S16 0.38268343236508977f

template <class V>
FASTB_HD void dft8(V (&v)[8]) {                             // n = n1 + 2 n2, k = k2 + 4 k1
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    v[3] = cmulc(v[3], FASTB_C8, FASTB_C8);                 // w8^1
    v[5] = cmuli(v[5]);                                     // w8^2
    v[7] = cmulc(v[7], -FASTB_C8, FASTB_C8);                // w8^3
    V o[8];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        o[k2] = cadd(v[2 * k2], v[2 * k2 + 1]);
        o[k2 + 4] = csub(v[2 * k2], v[2 * k2 + 1]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = o[i];
}

template <class V>
FASTB_HD void dft16(V (&v)[16]) {                           // n = n1 + 4 n2, k = k2 + 4 k1
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);  // -> v[n1 + 4 k2]
    v[5] = cmulc(v[5], FASTB_C16, FASTB_S16);               // (n1,k2) = (1,1): w16^1
    v[9] = cmulc(v[9], FASTB_C8, FASTB_C8);                 // (1,2) w^2
    v[13] = cmulc(v[13], FASTB_S16, FASTB_C16);             // (1,3) w^3
    v[6] = cmulc(v[6], FASTB_C8, FASTB_C8);                 // (2,1) w^2
    v[14] = cmulc(v[14], -FASTB_C8, FASTB_C8);              // (2,3) w^6
    v[7] = cmulc(v[7], FASTB_S16, FASTB_C16);               // (3,1) w^3
    v[11] = cmulc(v[11], -FASTB_C8, FASTB_C8);              // (3,2) w^6
    v[15] = cmulc(v[15], -FASTB_C16, -FASTB_S16);           // (3,3) w^9
    V o[16];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        V a = v[4 * k2], b = v[4 * k2 + 1], c = v[4 * k2 + 2], d = v[4 * k2 + 3];
        if (k2 == 2) {
            // (2,2): w16^4 = i on input c, folded into the butterfly: c' = i c
            const V s02 = caddi(a, c), d02 = csubi(a, c);   // a + i c, a - i c
            const V s13 = cadd(b, d), e13 = csub(b, d);
            a = cadd(s02, s13);
            c = csub(s02, s13);
            b = caddi(d02, e13);
            d = csubi(d02, e13);
        } else {
            dft4(a, b, c, d);                                // over n1 -> k1
        }
        o[k2] = a;
        o[k2 + 4] = b;
        o[k2 + 8] = c;
        o[k2 + 12] = d;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// inverse 32-point DFT of one line (tuning flavour): two 16-point DFTs + one radix-2 stage
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

// ---- shared-memory element access: interleaved (one line) or two planes (line pair) ---------
// `buf` holds kPlane float2 per plane; the pair type uses plane 0 for re = (A, B), plane 1 for im.
template <class V>
struct Smem;
template <>
struct Smem<float2> {
    static constexpr int kPlanes = 1;
    FASTB_HD static void put(float2* buf, int, int idx, float2 v) { buf[idx] = v; }
    FASTB_HD static float2 get(const float2* buf, int, int idx) { return buf[idx]; }
    // 16 consecutive elements starting at a 16-byte aligned index: 8 x 128-bit loads
    FASTB_HD static void get16(const float2* buf, int, int base, float2 (&v)[16]) {
        const float4* q = reinterpret_cast<const float4*>(buf + base);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 w = q[j];
            v[2 * j] = make_float2(w.x, w.y);
            v[2 * j + 1] = make_float2(w.z, w.w);
        }
    }
};
template <>
struct Smem<pc> {
    static constexpr int kPlanes = 2;
    FASTB_HD static void put(float2* buf, int plane, int idx, pc v) {
        buf[idx] = v.re;
        buf[plane + idx] = v.im;
    }
    FASTB_HD static pc get(const float2* buf, int plane, int idx) { return pc{buf[idx], buf[plane + idx]}; }
    FASTB_HD static void get16(const float2* buf, int plane, int base, pc (&v)[16]) {
        const float4* qr = reinterpret_cast<const float4*>(buf + base);
        const float4* qi = reinterpret_cast<const float4*>(buf + plane + base);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 r = qr[j], i = qi[j];
            v[2 * j] = pc{make_float2(r.x, r.y), make_float2(i.x, i.y)};
            v[2 * j + 1] = pc{make_float2(r.z, r.w), make_float2(i.z, i.w)};
        }
    }
};

// multiply v[1..15] by the 16 twiddles of one table row
FASTB_HD void apply_twiddle_row(float2 (&v)[16], const float2* row) {      // 8 x 128-bit loads
    const float4* q = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 w = q[j];
        if (j > 0) v[2 * j] = cmul(v[2 * j], make_float2(w.x, w.y));
        v[2 * j + 1] = cmul(v[2 * j + 1], make_float2(w.z, w.w));
    }
}
FASTB_HD void apply_twiddle_row(pc (&v)[16], const float4* row) {          // 15 x 128-bit loads
#pragma unroll
    for (int a = 1; a < 16; ++a) v[a] = cmul(v[a], row[a]);
}

template <int LOG2N, class V = float2>
struct LineFFT {
    using Value = V;
    using Tw = typename TwOf<V>::type;
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = 16;                            // elements per thread
    static constexpr int kLines = TwOf<V>::kLines;          // lines a thread group processes at once
    static constexpr int S1 = N / 16;                       // threads per line (pair)
    static constexpr bool kThree = (S1 >= 16);              // N >= 256
    static constexpr int S2 = kThree ? S1 / 16 : 1;
    static constexpr int SF = kThree ? S2 : S1;             // length of the last small DFTs
    // exchange-A padding: S2 == 1 uses 2 so that each thread's 16 gathered values are 16-byte
    // aligned (128-bit shared loads, conflict-free at a 144-byte lane stride)
    static constexpr int kPadA = kThree ? (S2 == 1 ? 2 : S2) : 0;
    static constexpr int kBufA = 16 * (S1 + kPadA);         // exchange A layout: a*(S1+pad)+t
    static constexpr int kBufC = 17 * S1;                   // exchange C layout: u*17 + e
    static constexpr int kPlane = (kBufA > kBufC) ? kBufA : kBufC;   // float2 per plane
    static constexpr int kBuf = kPlane * Smem<V>::kPlanes;           // float2 per line (pair)
    static_assert(LOG2N >= 6 && LOG2N <= 11, "N must be 64..2048");
    // the last stage can run on warp shuffles (phase_c_shfl): one line per group, 2 or 4 lanes per DFT
    static constexpr bool kShflC = kThree && (S2 == 2 || S2 == 4) && TwOf<V>::kLines == 1;

    // element index held in register m of thread t before phase A
    FASTB_HD static int n_in(int t, int m) { return t + S1 * m; }

    // Twiddle tables, one row of kTwRow entries per thread (16 used):
    //   twa[t * kTwRow + a]   = w_N^(t a)          t < S1, a < 16
    //   twb[t2 * kTwRow + a2] = w_N^(16 t2 a2)     t2 < S2, a2 < 16   (S2 > 1 only)
    static constexpr int kTwRow = TwOf<V>::kTwPerRow;
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

    // phase A: dft16 over m, twiddle, write exchange buffer
    // FASTB_SCALAR_STAGES (tuning builds): bit 0 = phase A, bit 1 = phase B butterflies in scalar FP32
    FASTB_HD static void dft16_stage(V (&v)[16], int stage_bit) {
#ifdef FASTB_SCALAR_STAGES
        if constexpr (TwOf<V>::kLines == 1) {
            if ((FASTB_SCALAR_STAGES) & stage_bit) {
                sc t[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = sc{v[i].x, v[i].y};
                dft16(t);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = make_float2(t[i].x, t[i].y);
                return;
            }
        }
#endif
        (void)stage_bit;
        dft16(v);
    }

    FASTB_HD static void phase_a(int t, V (&v)[16], const Tw* twa, float2* buf) {
        dft16_stage(v, 1);
        apply_twiddle_row(v, twa + t * kTwRow);
        if (kThree) {
#pragma unroll
            for (int a = 0; a < 16; ++a) Smem<V>::put(buf, kPlane, a * (S1 + kPadA) + t, v[a]);
        } else {
#pragma unroll
            for (int a = 0; a < 16; ++a) Smem<V>::put(buf, kPlane, t * 17 + a, v[a]);
        }
    }

    // phase B (N >= 256): gather, dft16 over m2, twiddle.  If S2 > 1 the caller must sync
    // and then call phase_b_store before phase C.
    FASTB_HD static void phase_b(int u, V (&v)[16], const Tw* twb, const float2* buf) {
        const int a = u / S2, t2 = u % S2;
        if (S2 == 1) {
            Smem<V>::get16(buf, kPlane, a * (S1 + kPadA), v);
        } else {
#pragma unroll
            for (int m2 = 0; m2 < 16; ++m2) v[m2] = Smem<V>::get(buf, kPlane, a * (S1 + kPadA) + t2 + S2 * m2);
        }
        dft16_stage(v, 2);
        if (S2 > 1) apply_twiddle_row(v, twb + t2 * kTwRow);
    }

    FASTB_HD static void phase_b_store(int u, const V (&v)[16], float2* buf) {
#pragma unroll
        for (int e = 0; e < 16; ++e) Smem<V>::put(buf, kPlane, u * 17 + e, v[e]);
    }

    // phase C: 16/SF DFTs of length SF across the SF threads of a group
    FASTB_HD static void phase_c(int u, V (&v)[16], const float2* buf) {
        constexpr int G = 16 / SF;
        const int g0 = (u / SF) * SF, j = u % SF;
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int t2 = 0; t2 < SF; ++t2) v[i * SF + t2] = Smem<V>::get(buf, kPlane, (g0 + t2) * 17 + i * SF + j);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            if (SF == 2) {
                dft2(v[i * SF], v[i * SF + 1]);
            } else if (SF == 4) {
                dft4(v[i * SF], v[i * SF + 1], v[i * SF + 2], v[i * SF + 3]);
            } else if (SF == 8) {
                V w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) w[q] = v[i * SF + q];
                dft8(w);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[i * SF + q] = w[q];
            }
        }
    }

#if defined(__CUDACC__)
    // Phase C through warp shuffles (S2 = 2 or 4, one line per thread group, device only): the S2
    // threads u, u^1 (, u^2, u^3) of a group are adjacent lanes, so the S2 x S2 transposes of the
    // last stage need no shared memory and no line barrier.  Thread j = u % S2 ends with
    //   v[i S2 + b2] = sum_t2 x_t2[i S2 + j] w_S2^(t2 b2)      (x_t2 = phase-B register file of lane t2)
    // exactly as phase_c leaves it.  KEEP (bit e = register e is consumed) prunes whole exchanges.
    template <unsigned KEEP>
    __device__ __forceinline__ static void phase_c_shfl(int u, float2 (&v)[16]) {
        static_assert(kThree && (S2 == 2 || S2 == 4), "shuffle stage serves N = 512 and 1024");
        constexpr unsigned kAll = 0xffffffffu;
        auto xchg = [](float2 s, int lane_xor) {
            return make_float2(__shfl_xor_sync(kAll, s.x, lane_xor), __shfl_xor_sync(kAll, s.y, lane_xor));
        };
        auto sel = [](bool c, float2 a, float2 b) { return make_float2(c ? a.x : b.x, c ? a.y : b.y); };
        const bool b0 = u & 1;
        const float s0 = b0 ? -1.f : 1.f;
        if (S2 == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool want_sum = (KEEP >> (2 * i)) & 1u, want_dif = (KEEP >> (2 * i + 1)) & 1u;
                if (!want_sum && !want_dif) continue;
                const float2 own = sel(b0, v[2 * i + 1], v[2 * i]);
                const float2 got = xchg(sel(b0, v[2 * i], v[2 * i + 1]), 1);
                if (want_sum) v[2 * i] = add2(own, got);
                if (want_dif) v[2 * i + 1] = mul2(sub2(own, got), bc2(s0));    // x_0 - x_1
            }
        } else {
            const bool b1 = u & 2;
            const float s1 = b1 ? -1.f : 1.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned k4 = (KEEP >> (4 * i)) & 15u;
                if (k4 == 0) continue;
                const bool even = k4 & 5u, odd = k4 & 10u;
                // step 1 (lane ^ 2): keep the two a2 whose owner shares my bit 1, trade the others
                const float2 kA = sel(b1, v[4 * i + 2], v[4 * i]), kB = sel(b1, v[4 * i + 3], v[4 * i + 1]);
                const float2 gA = xchg(sel(b1, v[4 * i], v[4 * i + 2]), 2);
                const float2 gB = xchg(sel(b1, v[4 * i + 1], v[4 * i + 3]), 2);
                // radix-2 over the t2 pair {t, t + 2}, t = u & 1: sums feed even b2, differences odd b2
                float2 y0 = make_float2(0.f, 0.f), y1 = y0, y2 = y0, y3 = y0;
                if (even) {
                    const float2 SA = add2(kA, gA), SB = add2(kB, gB);
                    const float2 mine = sel(b0, SB, SA), got = xchg(sel(b0, SA, SB), 1);   // lane ^ 1
                    y0 = add2(mine, got);
                    y2 = mul2(sub2(mine, got), bc2(s0));
                }
                if (odd) {
                    const float2 DA = mul2(sub2(kA, gA), bc2(s1)), DB = mul2(sub2(kB, gB), bc2(s1));
                    const float2 mine = sel(b0, DB, DA), got = xchg(sel(b0, DA, DB), 1);
                    const float2 lo = sel(b0, got, mine), hi = sel(b0, mine, got);       // t = 0, t = 1
                    y1 = caddi(lo, hi);
                    y3 = csubi(lo, hi);
                }
                if (k4 & 1u) v[4 * i] = y0;
                if (k4 & 2u) v[4 * i + 1] = y1;
                if (k4 & 4u) v[4 * i + 2] = y2;
                if (k4 & 8u) v[4 * i + 3] = y3;
            }
        }
    }

    // the whole line FFT with the shuffle last stage: two line syncs instead of four
    template <unsigned KEEP, typename Sync>
    __device__ __forceinline__ static void run_shfl(int u, float2 (&v)[16], const Tw* twa, const Tw* twb,
                                                    float2* buf, Sync sync) {
        phase_a(u, v, twa, buf);
        sync();
        phase_b(u, v, twb, buf);
        sync();                 // every gather is done: the next line may overwrite the buffer
        phase_c_shfl<KEEP>(u, v);
    }
#endif

    // the whole line FFT; `sync` synchronises the S1 threads of the line (pair)
    template <typename Sync>
    FASTB_HD static void run(int u, V (&v)[16], const Tw* twa, const Tw* twb, float2* buf, Sync sync) {
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
    FASTB_HD static constexpr int k_base(int u) {
        if (kThree) return (S2 == 1) ? u : (u / S2) + 16 * (u % S2);
        return u % SF;
    }
    FASTB_HD static constexpr int k_off(int e) {
        return kThree ? ((S2 == 1) ? 16 * e : 16 * S2 * (e / S2) + 256 * (e % S2))
                      : SF * (e / SF) + 16 * (e % SF);
    }
    FASTB_HD static constexpr int k_out(int u, int e) { return k_base(u) + k_off(e); }
    FASTB_HD static constexpr bool k_off_all_even() {
        for (int e = 0; e < 16; ++e)
            if (k_off(e) & 1) return false;
        return true;
    }
};

// Registers that can hold an output inside the centred window [N/2 - half, N/2 + half), over all
// threads of a line: bit e of the result.  A kernel that only consumes those registers lets the
// compiler drop the butterflies of the last stage that feed the others (output pruning at compile
// time; the crop is the same for every line of both passes).
template <class F>
FASTB_HD constexpr unsigned keep_mask(int half) {
    unsigned m = 0;
    for (int u = 0; u < F::S1; ++u)
        for (int e = 0; e < F::E; ++e) {
            const int k = F::k_out(u, e);
            if (k >= F::N / 2 - half && k < F::N / 2 + half) m |= 1u << e;
        }
    return m;
}

// Line FFT for N = 512, 1024, 2048 as R = N/256 interleaved 256-point transforms ("split first"):
//   stage 0  thread u holds x[u + S1 m]; the R elements n' + 256 q (q < R) of one residue n' < 256 sit in the
//            same thread, so the radix-R step over the TOP digit of n is done in registers:
//            Y_p[n'] = w_N^(n' p) sum_q x[n' + 256 q] w_R^(q p),   and   X[R k' + p] = DFT256(Y_p)[k'].
//            The R sequences are exchanged once through the line buffer (natural order, one region per p).
//   then     the 16 threads u = 16 p + t' run the 256-point transform of sequence p exactly as LineFFT<8> does
//            (16 x 16, one exchange inside the half-warp): register a2 of thread (p, t') holds k' = t' + 16 a2.
// Output k = R t' + p + 16 R a2: the register index is the TOP digit of k, so a centred crop window keeps only
// 6 (N = 512) / 4 (N = 1024) of the 16 outputs of the LAST 16-point DFT -- the compile-time pruning removes
// more than with the a + 16 (a2 + 16 b2) order of LineFFT -- and there is one twiddle pass less; the only
// line-wide synchronisations are the two around the first exchange.  TUNING FLAVOUR (FASTB_SPLIT=1): measured
// 5 - 8 % slower than LineFFT on B200 (profiles/experiments_r02.txt, 12): it executes fewer instructions but
// moves 1.5x the data through shared memory.
template <int LOG2N>
struct LineFFTSplit {
    using Value = float2;
    using Tw = float2;
    using Sub = LineFFT<8, float2>;
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = 16;
    static constexpr int kLines = 1;
    static constexpr int S1 = N / 16;                       // threads per line
    static constexpr int R = N / 256;                       // interleaved sub-transforms: 2, 4, 8
    static constexpr int kML = 16 / R;                      // radix-R steps per thread
    static_assert(LOG2N >= 9 && LOG2N <= 11, "split flavour serves N = 512, 1024, 2048");
    static constexpr int kRegion = Sub::kBuf;               // float2 per sub-transform (>= 256, padded)
    static constexpr int kBuf = R * kRegion;
    static constexpr int kTwRow = 18;                       // 16 used; 144-byte rows
    static constexpr int kTwA = Sub::kTwA;                  // w_256^(t a) = w_N^(R t a)
    static constexpr int kTwB = kTwRow * S1;                // stage-0 twiddles: row u, entry m_lo R + p
    static constexpr bool kShflC = false;
    template <unsigned KEEP, typename Sync>
    FASTB_HD static void run_shfl(int, float2 (&)[16], const float2*, const float2*, float2*, Sync) {}

    FASTB_HD static int n_in(int t, int m) { return t + S1 * m; }
    FASTB_HD static int twa_exponent(int idx) {
        const int t = idx / Sub::kTwRow, a = idx % Sub::kTwRow;
        return a < 16 ? (R * t * a) & (N - 1) : 0;
    }
    FASTB_HD static int twb_exponent(int idx) {
        const int u = idx / kTwRow, c = idx % kTwRow;
        return c < 16 ? ((u + 16 * R * (c / R)) * (c % R)) & (N - 1) : 0;
    }

    // radix-R step over q (register m = m_lo + kML q), twiddle, store sequence p at region p in natural order
    FASTB_HD static void stage0(int u, float2 (&v)[16], const float2* twb, float2* buf) {
        const float2* row = twb + u * kTwRow;
#pragma unroll
        for (int ml = 0; ml < kML; ++ml) {
            float2 y[R];
#pragma unroll
            for (int q = 0; q < R; ++q) y[q] = v[ml + kML * q];
            if constexpr (R == 2) dft2(y[0], y[1]);
            else if constexpr (R == 4) dft4(y[0], y[1], y[2], y[3]);
            else dft8(y);
            const int np = u + 16 * R * ml;
#pragma unroll
            for (int p = 0; p < R; ++p) {
                const float2 z = p == 0 ? y[0] : cmul(y[p], row[ml * R + p]);
                buf[p * kRegion + np] = z;
            }
        }
    }
    // thread u = 16 p + t' picks up sequence p in the 256-point transform's input layout
    FASTB_HD static void gather0(int u, float2 (&v)[16], const float2* buf) {
        const float2* reg = buf + (u / 16) * kRegion + (u % 16);
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = reg[16 * m];
    }

    template <typename Sync>
    FASTB_HD static void run(int u, float2 (&v)[16], const float2* twa, const float2* twb, float2* buf, Sync sync) {
        stage0(u, v, twb, buf);
        sync();
        gather0(u, v, buf);
        float2* reg = buf + (u / 16) * kRegion;
#if defined(__CUDA_ARCH__)
        __syncwarp();           // region p is read and rewritten by the same 16 lanes
#endif
        Sub::phase_a(u % 16, v, twa, reg);
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        Sub::phase_b(u % 16, v, nullptr, reg);
        sync();                 // the next line's stage 0 rewrites every region
    }

    FASTB_HD static constexpr int k_base(int u) { return R * (u % 16) + u / 16; }
    FASTB_HD static constexpr int k_off(int e) { return 16 * R * e; }
    FASTB_HD static constexpr int k_out(int u, int e) { return k_base(u) + k_off(e); }
    FASTB_HD static constexpr bool k_off_all_even() { return true; }
};

// Line FFT with 32 complex elements per thread for N = 512 (32 x 16) and N = 1024 (32 x 32):
// S1 = N/32 threads per line, ONE shared-memory exchange per line (tuning flavour, one line).
//   phase A  thread t holds x[t + S1 m], m < 32: 32-point DFT over m -> a = k mod 32, twiddle
//            w_N^(t a), store buf[a][t]
//   phase B  N = 1024: thread u = a gathers t < 32 and does a 32-point DFT -> b, k = a + 32 b
//            N = 512 : thread u handles a = u and a = u + 16, a 16-point DFT each, k = a + 32 b
template <int LOG2N>
struct LineFFT32 {
    using Value = float2;
    using Tw = float2;
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = 32;
    static constexpr int kLines = 1;
    static constexpr int S1 = N / 32;
    static_assert(LOG2N == 9 || LOG2N == 10, "LineFFT32 serves N = 512 and 1024");
    static constexpr int kRowA = S1 + 2;
    static constexpr int kBuf = 32 * kRowA;
    static constexpr int kTwRow = 34;
    static constexpr int kTwA = kTwRow * S1;
    static constexpr int kTwB = 0;
    static constexpr bool kShflC = false;
    template <unsigned KEEP, typename Sync>
    FASTB_HD static void run_shfl(int, float2 (&)[32], const float2*, const float2*, float2*, Sync) {}
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
                Smem<float2>::get16(buf, 0, (u + 16 * h) * kRowA, x);
                dft16(x);
#pragma unroll
                for (int b = 0; b < 16; ++b) v[16 * h + b] = x[b];
            }
        }
    }

    template <typename Sync>
    FASTB_HD static void run(int u, float2 (&v)[32], const float2* twa, const float2*, float2* buf, Sync sync) {
        phase_a(u, v, twa, buf);
        sync();
        phase_b(u, v, buf);
        sync();
    }

    FASTB_HD static constexpr int k_base(int u) { return u; }
    FASTB_HD static constexpr int k_off(int e) { return S1 == 32 ? 32 * e : 16 * (e / 16) + 32 * (e % 16); }
    FASTB_HD static constexpr int k_out(int u, int e) { return k_base(u) + k_off(e); }
    FASTB_HD static constexpr bool k_off_all_even() {
        for (int e = 0; e < 32; ++e)
            if (k_off(e) & 1) return false;
        return true;
    }
};

}  // namespace fastb
