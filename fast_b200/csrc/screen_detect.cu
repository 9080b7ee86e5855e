// K2 entry points (C ABI, include/fastb.h): argument checks, workspace layout, kernel selection.
// Device code: screen_detect_kernel.cuh; radix instances: screen_detect_radix.cu (one unit per grid
// size); general even N: screen_detect_bluestein.cu; this unit also holds the pruned direct-DFT
// kernel (cross-check / inspection path) and the small table-preparation kernels.
#include "screen_detect_kernel.cuh"
#include "bluestein.cuh"

namespace fastb {

TuneHook g_tune_hook = nullptr;
int g_l2_persist = 0;
int g_stagger = 0, g_wstagger = 0;

#define FASTB_DECL_SIZE(k)                                                                    \
    int launch_radix_##k(const RunArgs& a, const RadixRequest& rq, cudaStream_t st);         \
    int launch_pair_##k(const RunArgs& a, const RadixRequest& rq, cudaStream_t st);
FASTB_DECL_SIZE(6) FASTB_DECL_SIZE(7) FASTB_DECL_SIZE(8) FASTB_DECL_SIZE(9) FASTB_DECL_SIZE(10) FASTB_DECL_SIZE(11)
#undef FASTB_DECL_SIZE

int launch_radix_n(int log2n, const RunArgs& a, const RadixRequest& rq, cudaStream_t st) {
    if (g_tune_hook) {
        const int rc = g_tune_hook(log2n, a, rq, st);
        if (rc >= 0) return rc;
    }
    switch (log2n) {
        case 6: return launch_radix_6(a, rq, st);
        case 7: return launch_radix_7(a, rq, st);
        case 8: return launch_radix_8(a, rq, st);
        case 9: return launch_radix_9(a, rq, st);
        case 10: return launch_radix_10(a, rq, st);
        case 11: return launch_radix_11(a, rq, st);
        default: break;
    }
    set_error("radix kernel: unsupported size 2^%d", log2n);
    return FASTB_ERR_UNSUPPORTED;
}

int launch_pair_n(int log2n, const RunArgs& a, const RadixRequest& rq, cudaStream_t st) {
    switch (log2n) {
        case 6: return launch_pair_6(a, rq, st);
        case 7: return launch_pair_7(a, rq, st);
        case 8: return launch_pair_8(a, rq, st);
        case 9: return launch_pair_9(a, rq, st);
        case 10: return launch_pair_10(a, rq, st);
        case 11: return launch_pair_11(a, rq, st);
        default: break;
    }
    set_error("line-pair kernel: unsupported size 2^%d", log2n);
    return FASTB_ERR_UNSUPPORTED;
}

int radix_ctas_per_sm(int log2n) { return log2n <= 8 ? 4 : 1; }

// general even N through Bluestein's chirp-z on the radix line FFT (screen_detect_bluestein.cu)

namespace {

// ---- general even N: pruned direct DFT (slow path; also used for N not a power of two) ----
template <int RNG, bool SH>
__global__ void __launch_bounds__(kThreads) screen_detect_direct(const __grid_constant__ RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.n, P = a.n_pup, lo = a.lo, R = a.rows_per_block;
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float2* rows = tw + N;                       // R x N coloured noise
    double* st = reinterpret_cast<double*>(rows + (size_t)R * N);
    float* red = reinterpret_cast<float*>(st + kStatWords);
    float2* sh_amp = reinterpret_cast<float2*>(red + 4 * (kThreads / 32));     // SH only
    float2* sh_tab = sh_amp + 28;
    const int tid = threadIdx.x;

    for (int j = tid; j < N; j += kThreads) {
        double s, c;
        sincospi(2.0 * (double)j / (double)N, &s, &c);
        tw[j] = make_float2((float)c, (float)s);
    }
    if (tid == 0) stats_reset(st, 0);
    __syncthreads();

    float2* T = a.scratch + (size_t)blockIdx.x * N * P;

    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const PairId id = pair_id(a, pair);
        const unsigned long long g = id.g;
        const uint32_t k0 = (uint32_t)id.seed, k1 = (uint32_t)(id.seed >> 32);
        const float* weight = a.weight + (size_t)id.item * N * N;
        if (SH) sh_prepare(a, pair, id, sh_amp, sh_tab);    // followed by barriers in the row loop
        for (int row0 = 0; row0 < N; row0 += R) {
            const int nr = min(R, N - row0);
            if (RNG != kRngHost) {
                const int S = a.noise_stride;
                for (int idx = tid; idx < nr * S; idx += kThreads) {
                    const int rl = idx / S, t = idx % S, r = row0 + rl;
                    uint32_t mr[16], ma[16];
                    if (RNG == kRngFast) noise_block_fields_fast((uint32_t)(r * S + t), g, k0, k1, mr, ma);
                    else noise_block_fields((uint32_t)(r * S + t), g, k0, k1, mr, ma);
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int j = t + S * m;
                        if (j < N) rows[rl * N + j] = weighted_normal_m(mr[m], ma[m], weight[(size_t)r * N + j]);
                    }
                }
            } else {
                for (int idx = tid; idx < nr * N; idx += kThreads) {
                    const int rl = idx / N, j = idx % N, r = row0 + rl;
                    const float2 nz = a.noise[((size_t)pair * N + r) * N + j];
                    const float w0 = weight[(size_t)r * N + j];
                    rows[rl * N + j] = make_float2(nz.x * w0, nz.y * w0);
                }
            }
            __syncthreads();
            for (int idx = tid; idx < nr * P; idx += kThreads) {
                const int rl = idx / P, c = idx % P;
                const int kk = c + lo;                 // output column, 0 <= kk < N
                int ti = 0;
                float sr = 0.f, si = 0.f;
                const float2* row = rows + rl * N;
                for (int cp = 0; cp < N; ++cp) {
                    const float2 x = row[cp], t = tw[ti];
                    sr = fmaf(x.x, t.x, fmaf(-x.y, t.y, sr));
                    si = fmaf(x.x, t.y, fmaf(x.y, t.x, si));
                    ti += kk;
                    if (ti >= N) ti -= N;
                }
                T[(size_t)c * N + row0 + rl] = make_float2(sr, si);
            }
            __syncthreads();
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int idx = tid; idx < P * P; idx += kThreads) {
            const int c = idx / P, rr = idx % P;
            const int kk = rr + lo;
            const float2* col = T + (size_t)c * N;
            int ti = 0;
            float sr = 0.f, si = 0.f;
            for (int rp = 0; rp < N; ++rp) {
                const float2 x = __ldcg(col + rp), t = tw[ti];
                sr = fmaf(x.x, t.x, fmaf(-x.y, t.y, sr));
                si = fmaf(x.x, t.y, fmaf(x.y, t.x, si));
                ti += kk;
                if (ti >= N) ti -= N;
            }
            if (a.phs) {     // inspection seam (fastb_screens_crop): store phi, skip the detector
                const float sgn = ((kk + c + lo) & 1) ? -1.f : 1.f;
                float2 ph = make_float2(sgn * sr, sgn * si);
                if (SH) {
                    const float2 ex[3] = {a.sh_ex[c], a.sh_ex[P + c], a.sh_ex[2 * P + c]};
                    const float2 sp = sh_phase(sh_tab + rr * kShTab, ex);
                    ph.x += sp.x;
                    ph.y += sp.y;
                }
                float* dst = a.phs + (size_t)pair * 2 * P * P + (size_t)rr * P + c;
                dst[0] = ph.x;
                dst[(size_t)P * P] = ph.y;
                continue;
            }
            const float uu = a.u_t[(size_t)c * P + rr];
            if (SH) {
                const float sgn = ((kk + c + lo) & 1) ? -1.f : 1.f;
                const float2 ex[3] = {a.sh_ex[c], a.sh_ex[P + c], a.sh_ex[2 * P + c]};
                const float2 sp = sh_phase(sh_tab + rr * kShTab, ex);
                accumulate(make_float2(fmaf(sgn, sr, sp.x), fmaf(sgn, si, sp.y)), uu, uu, acc);
            } else {
                accumulate(make_float2(sr, si), uu, ((kk + c + lo) & 1) ? -uu : uu, acc);
            }
        }
        if (a.phs) {
            __syncthreads();          // scratch and tables are reused by the next pair
            continue;
        }
        finish_pair(a, pair, id, acc, red, st);
    }
    if (a.st_sums && tid == 0) stats_flush(a, st);
}

// weight * sqrt(2 ln 2) for the radix kernel's device-RNG path (weighted_normal_s), interleaved so
// that thread u of a line fetches its registers m = 4j .. 4j+3 (cells u + S1 m) with one 128-bit load
// per j and the threads of a line read consecutive 16-byte words
__global__ void scale_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int n, int s1,
                                    long long total) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const long long row = o / n;                 // row index over all stacked tables
    const int rem = (int)(o % n);
    const int j = rem / (4 * s1), u = (rem % (4 * s1)) / 4, q = rem % 4;
    out[o] = w[(size_t)row * n + u + s1 * (4 * j + q)] * kBoxMullerScale;
}

__global__ void transpose_u_kernel(const float* __restrict__ U, int P, float* __restrict__ u_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * P) return;
    const int r = i / P, c = i % P;
    u_t[(size_t)c * P + r] = U[i];
}

// u_p[(q * P + r) * 2 + j] = U[r][2q + j] (0 beyond the last column): U for the line-pair kernel
__global__ void pair_u_kernel(const float* __restrict__ U, int P, float* __restrict__ u_p) {
    const int PP = (P + 1) >> 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PP * P * 2) return;
    const int j = i & 1, r = (i >> 1) % P, q = (i >> 1) / P, c = 2 * q + j;
    u_p[i] = c < P ? U[(size_t)r * P + c] : 0.f;
}

__global__ void rng_dump_kernel(unsigned long long seed, unsigned long long g, int N, int S, int fast, float2* tile,
                                long long chi_first, long long chi_count, float* chi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile && i < (long long)N * S) {
        const int r = (int)(i / S), t = (int)(i % S);
        uint32_t mr[16], ma[16];
        if (fast) noise_block_fields_fast((uint32_t)(r * S + t), g, (uint32_t)seed, (uint32_t)(seed >> 32), mr, ma);
        else noise_block_fields((uint32_t)(r * S + t), g, (uint32_t)seed, (uint32_t)(seed >> 32), mr, ma);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int j = t + S * m;
            if (j < N) tile[(size_t)r * N + j] = weighted_normal_m(mr[m], ma[m], 1.0f);
        }
    }
    if (chi && i < chi_count) chi[i] = chi_normal(seed, (uint64_t)(chi_first + i));
}

int direct_rows(int n) {
    int r = (int)((96 * 1024) / ((size_t)n * sizeof(float2)));
    return r < 1 ? 1 : (r > 8 ? 8 : r);
}
size_t direct_smem_bytes(int n, bool sh, int n_pup) {
    return sizeof(float2) * ((size_t)n + (size_t)direct_rows(n) * n) + sizeof(double) * kStatWords +
           sizeof(float) * 4 * (kThreads / 32) + sh_smem_bytes(sh, n_pup);
}

bool radix_ok(int n) { return n >= 64 && n <= 2048 && (n & (n - 1)) == 0; }
// noise blocks per row of the device RNG (include/fastb.h): the threads per line of the transform that owns the grid
int noise_stride(int n, int n_pup) {
    if (radix_ok(n)) return n / 16;
    if (bluestein_ok(n, n_pup)) return blue_geom(n, n_pup).S1;
    return (n + 15) / 16;
}
int ilog2(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

int launch_direct(const RunArgs& a, int rng, bool has_sh, long long max_grid, cudaStream_t st, const char* what) {
    const size_t smem = direct_smem_bytes(a.n, has_sh, a.n_pup);
    void (*kern)(RunArgs) = nullptr;
    if (has_sh) kern = rng == kRngHost ? screen_detect_direct<kRngHost, true>
                       : rng == kRngFast ? screen_detect_direct<kRngFast, true> : screen_detect_direct<kRngPhilox, true>;
    else kern = rng == kRngHost ? screen_detect_direct<kRngHost, false>
                : rng == kRngFast ? screen_detect_direct<kRngFast, false> : screen_detect_direct<kRngPhilox, false>;
    FASTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    const int rc = sm_count(&sms);
    if (rc) return rc;
    long long grid = 2LL * sms;
    if (grid > a.n_pairs) grid = a.n_pairs;
    if (grid > max_grid) grid = max_grid;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(a);
    return check_launch(what);
}

// ---- which implementation serves a request ----------------------------------------------------
enum { kImplDirect = 0, kImplRadix = 1, kImplPair = 2, kImplBluestein = 3 };
int pick_impl(const FastbRunParams* p) {
    switch (p->algo) {
        case FASTB_ALGO_DIRECT: return kImplDirect;
        case FASTB_ALGO_RADIX: return kImplRadix;
        case FASTB_ALGO_RADIX_PAIR: return kImplPair;
        case FASTB_ALGO_BLUESTEIN: return kImplBluestein;
        default: break;
    }
    if (radix_ok(p->n)) return kImplRadix;
    if (bluestein_ok(p->n, p->n_pup)) return kImplBluestein;
    return kImplDirect;
}

// ---- workspace layout ---------------------------------------------------------------------------
// [U table: transposed (P*P) or pair-interleaved ((P+1)*P) | per item: weight * sqrt(2 ln 2),
//  interleaved (radix kernel, device RNG) or the Bluestein tables | scratch slots of N*(P+1) complex]
struct Layout {
    size_t u_bytes, tab_bytes, slot_bytes;
    size_t scratch_off() const { return u_bytes + tab_bytes; }
};
Layout layout_of(const FastbRunParams* p, int n_items, int impl) {
    Layout l;
    l.u_bytes = align_up(sizeof(float) * (size_t)(p->n_pup + 1) * p->n_pup, 256);
    // radix: real pre-scaled copies; chirp-z: complex chirped copies followed by the chirp tables
    if (impl == kImplBluestein)
        l.tab_bytes = align_up(bluestein_weight_bytes(p->n, p->n_pup, n_items), 256) +
                      align_up(bluestein_k2_table_bytes(p->n, p->n_pup), 256);
    else
        l.tab_bytes = align_up(sizeof(float) * (size_t)p->n * p->n * (n_items > 1 ? n_items : 1), 256);
    l.slot_bytes = (size_t)p->n * (p->n_pup + 1) * sizeof(float2);
    return l;
}

}  // namespace

int prepare_weight_s(int n, int n_items, const float* weight, float* weight_s, cudaStream_t st) {
    const long long total = (long long)n * n * n_items;
    scale_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(weight, weight_s, n, n / 16, total);
    return check_launch("scale_weight_kernel");
}

}  // namespace fastb

using namespace fastb;

static int validate_run(const FastbRunParams* p, const char* who) {
    FASTB_REQUIRE(p, "%s: NULL params", who);
    FASTB_REQUIRE(p->n >= 4 && (p->n % 2) == 0, "%s: n=%d must be even and >= 4", who, p->n);
    FASTB_REQUIRE(p->n_pup >= 1 && p->n_pup <= p->n, "%s: n_pup=%d outside 1..n", who, p->n_pup);
    FASTB_REQUIRE(p->lo >= 0 && p->lo + p->n_pup <= p->n, "%s: crop [%d,%d) outside grid", who, p->lo,
                  p->lo + p->n_pup);
    FASTB_REQUIRE(p->n_pairs >= 0 && p->first_pair >= 0, "%s: negative pair range", who);
    FASTB_REQUIRE(p->pairs_per_chunk > 0, "%s: pairs_per_chunk must be > 0", who);
    FASTB_REQUIRE(p->algo >= FASTB_ALGO_AUTO && p->algo <= FASTB_ALGO_BLUESTEIN, "%s: bad algo", who);
    FASTB_REQUIRE((p->flags & ~(FASTB_RUN_PREPARED | FASTB_RUN_RNG_FAST)) == 0, "%s: unknown flags 0x%x", who, p->flags);
    FASTB_REQUIRE(p->u_sum != 0.0, "%s: u_sum is zero", who);
    if ((p->algo == FASTB_ALGO_RADIX || p->algo == FASTB_ALGO_RADIX_PAIR) && !radix_ok(p->n)) {
        set_error("%s: radix path needs N = 64..2048 power of two, got %d", who, p->n);
        return FASTB_ERR_UNSUPPORTED;
    }
    if (p->algo == FASTB_ALGO_BLUESTEIN && !bluestein_ok(p->n, p->n_pup)) {
        set_error("%s: chirp-z path serves grids other than the powers of two 64..2048 with n + n_pup - 1 <= 2048, "
                  "got n=%d n_pup=%d", who, p->n, p->n_pup);
        return FASTB_ERR_UNSUPPORTED;
    }
    if (p->n > 4096) {
        set_error("%s: N=%d > 4096 not supported", who, p->n);
        return FASTB_ERR_UNSUPPORTED;
    }
    return FASTB_OK;
}

extern "C" int64_t fastb_screen_detect_batch_workspace_bytes(const FastbRunParams* p, int32_t n_items) {
    if (validate_run(p, "fastb_screen_detect_workspace_bytes")) return -1;
    if (n_items < 1) n_items = 1;
    int sms = 0;
    if (sm_count(&sms)) return -1;
    // scratch slots for the design occupancy of the implementation that will run (the launcher
    // clamps its grid to the slots it is given, so a smaller workspace still works)
    const int impl = pick_impl(p);
    int per_sm = 2;
    if (impl == kImplRadix) per_sm = radix_ctas_per_sm(ilog2(p->n));
    else if (impl == kImplPair) per_sm = p->n <= 256 ? 5 : 2;
    else if (impl == kImplBluestein) per_sm = bluestein_ctas_per_sm(p->n, p->n_pup);
    long long grid = (long long)sms * per_sm;
    if (grid > p->n_pairs) grid = p->n_pairs;
    if (grid < 1) grid = 1;
    const Layout l = layout_of(p, n_items, impl);
    return (int64_t)(l.scratch_off() + (size_t)grid * l.slot_bytes);
}

extern "C" int64_t fastb_screen_detect_workspace_bytes(const FastbRunParams* p) {
    return fastb_screen_detect_batch_workspace_bytes(p, 1);
}

// fills the derived tables of the workspace: U transposed, the pre-scaled interleaved weight copies
// (radix kernel) or the chirp tables (Bluestein).  Runs implicitly in every fastb_screen_detect*
// call unless FASTB_RUN_PREPARED is set.
static int prepare_tables(const FastbRunParams* p, int n_items, int impl, const float* d_weight, const float* d_U,
                          void* d_workspace, const Layout& l, cudaStream_t st) {
    int rc;
    if (impl == kImplPair) {
        const int cnt = ((p->n_pup + 1) / 2) * p->n_pup * 2;
        pair_u_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(d_U, p->n_pup, (float*)d_workspace);
        return check_launch("pair_u_kernel");
    }
    const int pp = p->n_pup * p->n_pup;
    transpose_u_kernel<<<(pp + 255) / 256, 256, 0, st>>>(d_U, p->n_pup, (float*)d_workspace);
    if ((rc = check_launch("transpose_u_kernel"))) return rc;
    char* tab = (char*)d_workspace + l.u_bytes;
    if (impl == kImplRadix && d_weight)
        return prepare_weight_s(p->n, n_items, d_weight, (float*)tab, st);
    if (impl == kImplBluestein) {
        const size_t w_bytes = align_up(bluestein_weight_bytes(p->n, p->n_pup, n_items), 256);
        return bluestein_prepare_k2(p->n, p->n_pup, p->lo, n_items, d_weight, tab, tab + w_bytes, st);
    }
    return FASTB_OK;
}

extern "C" int fastb_screen_detect_prepare(const FastbRunParams* p, int32_t n_items, const float* d_weight,
                                           const float* d_U, void* d_workspace, int64_t workspace_bytes,
                                           void* stream) {
    int rc = validate_run(p, "fastb_screen_detect_prepare");
    if (rc) return rc;
    if (n_items < 1) n_items = 1;
    FASTB_REQUIRE(d_weight && d_U && d_workspace, "fastb_screen_detect_prepare: NULL pointer");
    FASTB_REQUIRE(((uintptr_t)d_weight & 15) == 0 && ((uintptr_t)d_workspace & 255) == 0,
                  "fastb_screen_detect_prepare: d_weight must be 16-byte and d_workspace 256-byte aligned");
    const int impl = pick_impl(p);
    const Layout l = layout_of(p, n_items, impl);
    FASTB_REQUIRE(workspace_bytes >= (int64_t)(l.scratch_off() + l.slot_bytes),
                  "fastb_screen_detect_prepare: workspace too small (%lld B)", (long long)workspace_bytes);
    return prepare_tables(p, n_items, impl, d_weight, d_U, d_workspace, l, (cudaStream_t)stream);
}

static int run_impl(const char* who, const FastbRunParams* p, const FastbRunBatch* batch, const FastbRunStats* stats,
                    const float* d_weight, const float* d_U, const float* d_chi, const float* d_noise,
                    const FastbSubharm* sh, float* d_out_a, float* d_out_b, void* d_workspace, int64_t workspace_bytes,
                    void* stream) {
    int rc = validate_run(p, who);
    if (rc) return rc;
    FASTB_REQUIRE(d_weight && d_U && d_out_a && d_out_b && d_workspace, "%s: NULL pointer", who);
    FASTB_REQUIRE(((uintptr_t)d_weight & 15) == 0 && ((uintptr_t)d_workspace & 255) == 0,
                  "%s: d_weight must be 16-byte and d_workspace 256-byte aligned", who);
    const int n_items = batch ? batch->n_items : 1;
    if (batch) {
        FASTB_REQUIRE(n_items >= 1 && batch->pairs_per_item > 0, "%s: bad batch geometry", who);
        FASTB_REQUIRE(batch->d_sigma_chi && batch->d_seeds, "%s: batch tables must not be NULL", who);
        FASTB_REQUIRE(p->first_pair + p->n_pairs <= (int64_t)n_items * batch->pairs_per_item,
                      "%s: pair range beyond the batch (%lld items x %lld pairs)", who, (long long)n_items,
                      (long long)batch->pairs_per_item);
        FASTB_REQUIRE(!d_noise && !sh && !d_chi, "%s: a batch uses the device RNG and has no sub-harmonic term", who);
    }
    if (stats) {
        FASTB_REQUIRE(stats->d_sums && stats->d_minmax && stats->d_hist, "%s: statistics buffers must not be NULL", who);
        FASTB_REQUIRE(stats->nbins >= 1 && stats->db_hi > stats->db_lo, "%s: bad histogram range", who);
    }
    if (p->n_pairs == 0) return FASTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int impl = pick_impl(p);
    const bool fast = (p->flags & FASTB_RUN_RNG_FAST) != 0;
    const int rng = d_noise ? kRngHost : (fast ? kRngFast : kRngPhilox);
    if (impl == kImplPair) {
        FASTB_REQUIRE(!batch && !fast, "%s: the line-pair kernel serves single configurations of the default stream", who);
    }
    const Layout l = layout_of(p, n_items, impl);
    FASTB_REQUIRE(workspace_bytes >= (int64_t)(l.scratch_off() + l.slot_bytes), "%s: workspace too small (%lld B)", who,
                  (long long)workspace_bytes);
    long long max_grid = (long long)(((size_t)workspace_bytes - l.scratch_off()) / l.slot_bytes);
    if (max_grid > (1 << 20)) max_grid = 1 << 20;

    RunArgs a = {};
    a.n = p->n;
    a.n_pup = p->n_pup;
    a.lo = p->lo;
    a.coherent = p->coherent;
    a.n_pairs = p->n_pairs;
    a.first_pair = p->first_pair;
    a.ppc = p->pairs_per_chunk;
    a.seed = p->seed;
    a.inv_usum = (float)(1.0 / p->u_sum);
    a.sigma_chi = p->sigma_chi;
    a.weight = d_weight;
    a.weight_s = (float*)((char*)d_workspace + l.u_bytes);
    a.u_t = (const float*)d_workspace;
    a.u_p = (const float2*)d_workspace;
    a.chi = d_chi;
    a.noise = (const float2*)d_noise;
    a.out_a = d_out_a;
    a.out_b = d_out_b;
    a.scratch = (float2*)((char*)d_workspace + l.scratch_off());
    a.rows_per_block = direct_rows(p->n);
    a.noise_stride = noise_stride(p->n, p->n_pup);
    a.n_items = n_items;
    a.ppi = batch ? batch->pairs_per_item : 0;
    a.item_sigma = batch ? batch->d_sigma_chi : nullptr;
    a.item_seed = batch ? (const unsigned long long*)batch->d_seeds : nullptr;
    if (stats) {
        a.st_sums = stats->d_sums;
        a.st_minmax = stats->d_minmax;
        a.st_hist = stats->d_hist;
        a.st_lo = stats->db_lo;
        a.st_hi = stats->db_hi;
        a.st_nbins = stats->nbins;
    }
#ifdef FASTB_TUNE_DBG
    a.dbg = getenv("FASTB_DBG") ? atoi(getenv("FASTB_DBG")) : 0;
#endif
    if (sh) {
        FASTB_REQUIRE(sh->d_weight && sh->d_ex && sh->d_ey && sh->d_mean, "%s: sub-harmonic tables must not be NULL", who);
        FASTB_REQUIRE((sh->d_noise == nullptr) == (d_noise == nullptr),
                      "%s: d_noise and sh->d_noise must both be given or both be NULL", who);
        a.sh_weight = sh->d_weight;
        a.sh_noise = (const float2*)sh->d_noise;
        a.sh_ex = (const float2*)sh->d_ex;
        a.sh_ey = (const float2*)sh->d_ey;
        a.sh_mean = (const float2*)sh->d_mean;
    }

    const bool prepared = (p->flags & FASTB_RUN_PREPARED) != 0 && impl != kImplPair;
    if (!prepared) {
        // the radix launcher fills weight_s itself only when the device RNG needs it
        rc = prepare_tables(p, n_items, impl,
                            (impl == kImplBluestein || (impl == kImplRadix && rng != kRngHost)) ? d_weight : nullptr, d_U,
                            d_workspace, l, st);
        if (rc) return rc;
    }
    RadixRequest rq;
    rq.rng = rng;
    rq.max_grid = (int)max_grid;
    rq.prepared = true;          // prepare_tables has just run, or the caller vouches for the tables
    switch (impl) {
        case kImplPair: return launch_pair_n(ilog2(p->n), a, rq, st);
        case kImplRadix: return launch_radix_n(ilog2(p->n), a, rq, st);
        case kImplBluestein: {
            const size_t w_bytes = align_up(bluestein_weight_bytes(p->n, p->n_pup, n_items), 256);
            return launch_bluestein(a, rq, (const char*)d_workspace + l.u_bytes + w_bytes, st);
        }
        default: break;
    }
    return launch_direct(a, rng, sh != nullptr, max_grid, st, "screen_detect_direct");
}

extern "C" int fastb_screen_detect(const FastbRunParams* p, const float* d_weight, const float* d_U,
                                   const float* d_chi, const float* d_noise, const FastbSubharm* sh,
                                   float* d_out_a, float* d_out_b, void* d_workspace,
                                   int64_t workspace_bytes, void* stream) {
    return run_impl("fastb_screen_detect", p, nullptr, nullptr, d_weight, d_U, d_chi, d_noise, sh, d_out_a, d_out_b,
                    d_workspace, workspace_bytes, stream);
}

extern "C" int fastb_screen_detect_batch(const FastbRunParams* p, const FastbRunBatch* batch,
                                         const FastbRunStats* stats, const float* d_weight, const float* d_U,
                                         const float* d_chi, float* d_out_a, float* d_out_b, void* d_workspace,
                                         int64_t workspace_bytes, void* stream) {
    return run_impl("fastb_screen_detect_batch", p, batch, stats, d_weight, d_U, d_chi, nullptr, nullptr, d_out_a,
                    d_out_b, d_workspace, workspace_bytes, stream);
}

extern "C" int fastb_rng_dump(uint64_t seed, int64_t pair, int32_t n, float* d_noise_tile,
                              int64_t chi_first, int64_t chi_count, float* d_chi_normals, void* stream) {
    return fastb_rng_dump_mode(seed, pair, n, 0, d_noise_tile, chi_first, chi_count, d_chi_normals, stream);
}

extern "C" int fastb_rng_dump_mode(uint64_t seed, int64_t pair, int32_t n, int32_t fast, float* d_noise_tile,
                                     int64_t chi_first, int64_t chi_count, float* d_chi_normals, void* stream) {
    return fastb_rng_dump_stride(seed, pair, n, (n + 15) / 16, fast, d_noise_tile, chi_first, chi_count, d_chi_normals,
                                 stream);
}

extern "C" int32_t fastb_noise_stride(int32_t n, int32_t n_pup) {
    if (n < 2 || (n % 2) || n_pup < 1 || n_pup > n) return -1;
    return noise_stride(n, n_pup);
}

extern "C" int fastb_rng_dump_stride(uint64_t seed, int64_t pair, int32_t n, int32_t stride, int32_t fast,
                                     float* d_noise_tile, int64_t chi_first, int64_t chi_count, float* d_chi_normals,
                                     void* stream) {
    FASTB_REQUIRE(n >= 2 && (n % 2) == 0, "fastb_rng_dump: n must be even");
    FASTB_REQUIRE(stride >= 1 && (long long)stride * 16 >= n, "fastb_rng_dump: 16 stride must cover n");
    FASTB_REQUIRE(pair >= 0 && chi_first >= 0 && chi_count >= 0, "fastb_rng_dump: negative index");
    long long work = d_noise_tile ? (long long)n * stride : 0;
    if (d_chi_normals && chi_count > work) work = chi_count;
    if (work == 0) return FASTB_OK;
    rng_dump_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        seed, (unsigned long long)pair, n, stride, fast ? 1 : 0, (float2*)d_noise_tile, chi_first, chi_count,
        d_chi_normals);
    return check_launch("rng_dump_kernel");
}

extern "C" int fastb_screens_crop(const FastbRunParams* p, const float* d_weight, const float* d_noise,
                                  const FastbSubharm* sh, float* d_phs, void* d_workspace,
                                  int64_t workspace_bytes, void* stream) {
    int rc = validate_run(p, "fastb_screens_crop");
    if (rc) return rc;
    FASTB_REQUIRE(d_weight && d_phs && d_workspace, "fastb_screens_crop: NULL pointer");
    if (p->n_pairs == 0) return FASTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const Layout l = layout_of(p, 1, kImplDirect);
    FASTB_REQUIRE(workspace_bytes >= (int64_t)(l.scratch_off() + l.slot_bytes), "fastb_screens_crop: workspace too small");
    long long max_grid = (long long)(((size_t)workspace_bytes - l.scratch_off()) / l.slot_bytes);
    RunArgs a = {};
    a.n = p->n;
    a.n_pup = p->n_pup;
    a.lo = p->lo;
    a.n_pairs = p->n_pairs;
    a.first_pair = p->first_pair;
    a.ppc = p->pairs_per_chunk;
    a.seed = p->seed;
    a.weight = d_weight;
    a.noise = (const float2*)d_noise;
    a.scratch = (float2*)((char*)d_workspace + l.scratch_off());
    a.rows_per_block = direct_rows(p->n);
    a.noise_stride = noise_stride(p->n, p->n_pup);
    a.n_items = 1;
    a.phs = d_phs;
    if (sh) {
        FASTB_REQUIRE(sh->d_weight && sh->d_ex && sh->d_ey && sh->d_mean,
                      "fastb_screens_crop: sub-harmonic tables must not be NULL");
        FASTB_REQUIRE((sh->d_noise == nullptr) == (d_noise == nullptr),
                      "fastb_screens_crop: d_noise and sh->d_noise must both be given or both be NULL");
        a.sh_weight = sh->d_weight;
        a.sh_noise = (const float2*)sh->d_noise;
        a.sh_ex = (const float2*)sh->d_ex;
        a.sh_ey = (const float2*)sh->d_ey;
        a.sh_mean = (const float2*)sh->d_mean;
    }
    const int rng = d_noise ? kRngHost : ((p->flags & FASTB_RUN_RNG_FAST) ? kRngFast : kRngPhilox);
    return launch_direct(a, rng, sh != nullptr, max_grid, st, "screen_detect_direct(screens)");
}
