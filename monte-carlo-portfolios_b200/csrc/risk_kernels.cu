// risk_kernels.cu -- kernels (1)-(3) of the hot path: portfolio x scenario valuation,
// per-portfolio moments + exact tail selection (VaR order statistic with its trial index,
// CVaR), and the sorted breach-trial list.  DESIGN.md section 3 is the specification
// (the reference has no code for these stages); the oracle is oracle/risk_oracle.c.
//
// Bit-exactness contract: V[p][n] is ONE fma chain over k ascending starting from +0.0,
// on both sides, so every loss is bit-identical to the oracle's and the selected trial
// indices are exact.
#include "common.cuh"

namespace mcp {

constexpr unsigned kFullMask = 0xffffffffu;

// ============================================================ Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&o)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t lo0 = 0xD2511F53u * c0, hi0 = __umulhi(0xD2511F53u, c0);
        const uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = __umulhi(0xCD9E8D57u, c2);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// Irwin-Hall(12) variate from three Philox blocks; see oracle/risk_oracle.c orc_gen_normal.
__global__ void k_gen_matrix(double *out, int64_t rows, int32_t cols, uint32_t k0, uint32_t k1, uint32_t tensor_id,
                             double scale, double shift, int64_t row0)
{
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t r = (uint64_t)(row0 + i / cols);
        const uint32_t c = (uint32_t)(i % cols);
        uint64_t sum = 0;
#pragma unroll
        for (uint32_t j = 0; j < 3; ++j) {
            uint32_t o[4];
            philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), c, 4u * tensor_id + j, k0, k1, o);
            sum += (uint64_t)o[0] + o[1] + o[2] + o[3];
        }
        const double x = (double)((long long)sum - (long long)(6ull << 32)) * 0x1p-32;
        out[i] = fma(x, scale, shift);
    }
}

int gen_matrix(mcp_ctx *ctx, double *d_out, int64_t rows, int32_t cols, uint64_t seed, uint32_t tensor_id, double scale,
               double shift, int64_t row0)
{
    if (rows * cols == 0) return MCP_OK;
    KScope ks(ctx, MCP_K_GEN);
    k_gen_matrix<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(d_out, rows, cols, (uint32_t)seed,
                                                                            (uint32_t)(seed >> 32), tensor_id, scale, shift, row0);
    return check_launch(ctx, "k_gen_matrix");
}

// Bernoulli(rate) breach lists: one warp per list, 128 trials per round (4 per lane, one Philox block each)
template <bool WRITE>
__global__ void k_gen_breach(int64_t nlists, int32_t n_trials, uint32_t thresh, uint32_t k0, uint32_t k1,
                             int64_t first_list, int64_t *counts, const int64_t *offs, int32_t *vals)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t li = warp; li < nlists; li += nwarps) {
        const uint64_t gl = (uint64_t)(first_list + li);
        int64_t cnt = 0;
        int32_t *dst = WRITE ? vals + offs[li] : nullptr;
        for (int32_t base = 0; base < n_trials; base += 128) {
            const int32_t n0 = base + 4 * lane;
            uint32_t m = 0;
            if (n0 < n_trials) {
                uint32_t o[4];
                philox4x32_10((uint32_t)(n0 >> 2), (uint32_t)gl, (uint32_t)(gl >> 32), 0xB4EAC400u, k0, k1, o);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + j < n_trials && o[j] < thresh) m |= 1u << j;
            }
            const int c = __popc(m);
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFullMask, incl, d);
                if (lane >= d) incl += t;
            }
            if (WRITE) {
                int64_t o = cnt + incl - c;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m & (1u << j)) dst[o++] = n0 + j;
            }
            cnt += __shfl_sync(kFullMask, incl, 31);
        }
        if (!WRITE && lane == 0) counts[li] = cnt;
    }
}

int gen_breach_lists(mcp_ctx *ctx, int64_t nlists, int32_t n_trials, double rate, uint64_t seed, int64_t first_list,
                     int64_t *total)
{
    const uint32_t thresh = (uint32_t)(rate * 4294967296.0);
    MCP_CUDA(ctx, ctx->dSizes.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    MCP_CUDA(ctx, ctx->dGenOff.reserve((size_t)(nlists + 1) * sizeof(int64_t)));
    const int grid = ctx->prop.multiProcessorCount * 8;
    {
        KScope ks(ctx, MCP_K_GEN);
        k_gen_breach<false><<<grid, 256, 0, ctx->stream>>>(nlists, n_trials, thresh, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                          first_list, ctx->dSizes.as<int64_t>(), nullptr, nullptr);
        MCP_TRY(check_launch(ctx, "k_gen_breach<count>"));
    }
    MCP_TRY(exclusive_scan_i64(ctx, ctx->dSizes.as<int64_t>(), nlists, ctx->dGenOff.as<int64_t>()));
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dGenOff.as<int64_t>() + nlists, sizeof(int64_t), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *total = *reinterpret_cast<int64_t *>(ctx->hPinned);
    MCP_CUDA(ctx, ctx->dGenVals.reserve((size_t)(*total + 1) * sizeof(int32_t)));
    {
        KScope ks(ctx, MCP_K_GEN);
        k_gen_breach<true><<<grid, 256, 0, ctx->stream>>>(nlists, n_trials, thresh, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                         first_list, nullptr, ctx->dGenOff.as<int64_t>(), ctx->dGenVals.as<int32_t>());
        MCP_TRY(check_launch(ctx, "k_gen_breach<fill>"));
    }
    return MCP_OK;
}

// ============================================================ FP64 peak probe
// 8 independent DFMA chains per thread, no memory traffic: measures the FP64 FMA issue
// rate that the valuation kernel is bounded by (MEASURED_PEAKS.json has no FP64 entry).
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 1.2345e-300) out[0] = s; // never true; keeps the chains alive
}

int fp64_peak(mcp_ctx *ctx, double *tflops)
{
    MCP_CUDA(ctx, ctx->dErr.reserve(64));
    const int grid = ctx->prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t a, b;
    MCP_CUDA(ctx, cudaEventCreate(&a));
    MCP_CUDA(ctx, cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        KScope ks(ctx, MCP_K_GEN);
        cudaEventRecord(a, ctx->stream);
        k_fp64_peak<<<grid, 256, 0, ctx->stream>>>(ctx->dErr.as<double>(), iters, 0.999999, 1e-9);
        cudaEventRecord(b, ctx->stream);
        MCP_CUDA(ctx, cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return check_launch(ctx, "k_fp64_peak");
}

// ============================================================ valuation (generic F)
// V tile = 64 portfolios x 64 trials per CTA, 4x4 register micro-tile per thread, k staged
// through shared memory in slabs of 32.  Every (p,n) accumulator is a single fma chain in
// ascending k.  Writes either V or loss = 0.0 - V, row-major [p][n].
constexpr int VT = 64, VK = 32;

template <bool LOSS>
__global__ void __launch_bounds__(256) k_value_generic(const double *__restrict__ E, const double *__restrict__ S,
                                                       int64_t p0, int64_t P1, int64_t N, int32_t F,
                                                       double *__restrict__ out /* [(P1-p0)][N] */)
{
    __shared__ double Es[VK][VT + 1];
    __shared__ double Ss[VK][VT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t pb = p0 + (int64_t)blockIdx.y * VT, nb = (int64_t)blockIdx.x * VT;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int32_t k0 = 0; k0 < F; k0 += VK) {
        for (int idx = threadIdx.x; idx < VT * VK; idx += 256) {
            const int r = idx / VK, k = idx % VK;
            Es[k][r] = (pb + r < P1 && k0 + k < F) ? E[(pb + r) * F + k0 + k] : 0.0;
            Ss[k][r] = (nb + r < N && k0 + k < F) ? S[(nb + r) * F + k0 + k] : 0.0;
        }
        __syncthreads();
        const int kend = min(VK, F - k0);
        for (int k = 0; k < kend; ++k) {
            double e[4], s[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = Es[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] = Ss[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(e[i], s[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t p = pb + ty + 16 * i;
        if (p >= P1) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = nb + tx + 16 * j;
            if (n < N) out[(p - p0) * N + n] = LOSS ? 0.0 - acc[i][j] : acc[i][j];
        }
    }
}

// ============================================================ selection
// order-preserving image of a double (larger value -> larger key)
__device__ __forceinline__ uint64_t fkey(double x)
{
    const uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double fkey_inv(uint64_t k)
{
    const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

typedef unsigned __int128 u128;
__device__ __forceinline__ u128 comp_of(uint64_t key, uint32_t idx) { return ((u128)key << 32) | idx; }

constexpr int SEL_THREADS = 512;
constexpr int SEL_BINS = 4096;   // 12-bit digits over the 96-bit composite (key64, trial32)
constexpr int SEL_CAP = 8192;    // candidates resolved in shared memory
constexpr size_t SEL_SMEM = (size_t)SEL_CAP * 12 + (size_t)SEL_BINS * 4;

struct SelArgs {
    const double *loss;      // dense: [rows][row_stride]; candidate i of row r = loss[r*row_stride + i]
    const int32_t *idx;      // null -> trial = idx_base + i
    int64_t row_stride;
    const int32_t *count;    // null -> `n` candidates per row
    int64_t n;
    int32_t idx_base;
    int64_t tail;            // t
    int64_t moments_n;       // N for the moments (dense rows only; 0 = skip)
    uint64_t *kept_key;      // [rows][tail]
    int32_t *kept_idx;       // [rows][tail]
    double *mean, *variance, *var_loss;
    int32_t *var_trial;
};

__device__ __forceinline__ double block_sum(double v, double *red /* >= 32 */)
{
    for (int d = 16; d; d >>= 1) v += __shfl_down_sync(kFullMask, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int d = 16; d; d >>= 1) t += __shfl_down_sync(kFullMask, t, d);
    }
    __syncthreads();
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    t = red[0];
    __syncthreads();
    return t;
}

// One CTA per portfolio row: exact top-`tail` of (loss, trial) by MSB-first radix select.
__global__ void __launch_bounds__(SEL_THREADS) k_select(const SelArgs a)
{
    extern __shared__ __align__(16) unsigned char sel_smem[];
    uint64_t *ckey = reinterpret_cast<uint64_t *>(sel_smem);
    uint32_t *cidx = reinterpret_cast<uint32_t *>(ckey + SEL_CAP);
    int *hist = reinterpret_cast<int *>(cidx + SEL_CAP);
    __shared__ double red[32];
    __shared__ int s_sel, s_above, s_bucket, s_ncand, s_nkept;
    const int64_t row = blockIdx.x;
    const double *L = a.loss + row * a.row_stride;
    const int32_t *I = a.idx ? a.idx + row * a.row_stride : nullptr;
    const int64_t n = a.count ? a.count[row] : a.n;
    const int64_t t = a.tail;
    uint64_t *kk = a.kept_key + row * t;
    int32_t *ki = a.kept_idx + row * t;
    const int tid = threadIdx.x;

    // ---- moments of V = -loss with shift c = V[0]
    if (a.moments_n > 0) {
        const double c = -L[0];
        double s1 = 0.0, s2 = 0.0;
        for (int64_t i = tid; i < n; i += SEL_THREADS) {
            const double d = -L[i] - c;
            s1 += d;
            s2 = fma(d, d, s2);
        }
        s1 = block_sum(s1, red);
        s2 = block_sum(s2, red);
        if (tid == 0) {
            const double invn = 1.0 / (double)a.moments_n;
            const double m1 = s1 * invn;
            a.mean[row] = c + m1;
            a.variance[row] = fma(-m1, m1, s2 * invn);
        }
    }

    u128 prefix = 0;     // selected high digits
    int shift = 96;      // bits below the prefix
    int64_t need = t;    // rank from the top still to resolve inside the prefix bucket
    if (tid == 0) s_nkept = 0;
    int bucket = 0;
    for (int pass = 0; pass < 8; ++pass) {
        shift -= 12;
        for (int b = tid; b < SEL_BINS; b += SEL_THREADS) hist[b] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n; i += SEL_THREADS) {
            const u128 c = comp_of(fkey(L[i]), (uint32_t)(I ? I[i] : a.idx_base + (int32_t)i));
            if (pass == 0 || (c >> (shift + 12)) == prefix) atomicAdd(&hist[(uint32_t)(c >> shift) & (SEL_BINS - 1)], 1);
        }
        __syncthreads();
        // warp 0: find the digit where the count from the top reaches `need`
        if (tid < 32) {
            constexpr int PER = SEL_BINS / 32;
            const int hi = SEL_BINS - 1 - tid * PER; // lane 0 owns the highest bins
            int sum = 0;
            for (int j = 0; j < PER; ++j) sum += hist[hi - j];
            int incl = sum;
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(kFullMask, incl, d);
                if (tid >= d) incl += v;
            }
            const int excl = incl - sum;
            if ((int64_t)excl < need && need <= (int64_t)incl) {
                int above = excl;
                for (int j = 0; j < PER; ++j) {
                    const int h = hist[hi - j];
                    if (need <= (int64_t)above + h) { s_sel = hi - j; s_above = above; s_bucket = h; break; }
                    above += h;
                }
            }
        }
        __syncthreads();
        prefix = (prefix << 12) | (u128)(uint32_t)s_sel;
        need -= s_above;
        bucket = s_bucket;
        __syncthreads();
        if (bucket <= SEL_CAP) break;
    }
    // ---- gather: strictly-above-prefix -> kept; equal-prefix -> shared candidates
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += SEL_THREADS) {
        const uint64_t key = fkey(L[i]);
        const uint32_t id = (uint32_t)(I ? I[i] : a.idx_base + (int32_t)i);
        const u128 hp = comp_of(key, id) >> shift;
        if (hp > prefix) {
            const int o = atomicAdd(&s_nkept, 1);
            kk[o] = key;
            ki[o] = (int32_t)id;
        } else if (hp == prefix) {
            const int o = atomicAdd(&s_ncand, 1);
            ckey[o] = key;
            cidx[o] = id;
        }
    }
    __syncthreads();
    // ---- bitonic sort of the candidates, descending composite; pads (0,0) sink to the end
    int m = 1;
    while (m < bucket) m <<= 1;
    for (int i = bucket + tid; i < m; i += SEL_THREADS) { ckey[i] = 0; cidx[i] = 0; }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += SEL_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const u128 ci = comp_of(ckey[i], cidx[i]), cp = comp_of(ckey[p], cidx[p]);
                    const bool desc = (i & k) == 0;
                    if (desc ? (ci < cp) : (ci > cp)) {
                        const uint64_t tk = ckey[i]; ckey[i] = ckey[p]; ckey[p] = tk;
                        const uint32_t ti = cidx[i]; cidx[i] = cidx[p]; cidx[p] = ti;
                    }
                }
            }
            __syncthreads();
        }
    }
    const int base = s_nkept; // == t - need
    for (int i = tid; i < (int)need; i += SEL_THREADS) {
        kk[base + i] = ckey[i];
        ki[base + i] = (int32_t)cidx[i];
    }
    if (tid == 0) { // the need-th largest candidate is the VaR order statistic
        a.var_loss[row] = fkey_inv(ckey[need - 1]);
        a.var_trial[row] = (int32_t)cidx[need - 1];
    }
}

// ============================================================ finalize
// Sort the kept tail by trial index (ascending) -> breach list; CVaR = mean tail loss.
constexpr int FIN_THREADS = 256;

__global__ void __launch_bounds__(FIN_THREADS) k_finalize(const uint64_t *kept_key, const int32_t *kept_idx, int64_t t,
                                                          int m /* pow2 >= t */, int32_t *breach, double *cvar,
                                                          uint64_t *gkey /* global scratch or null */, int32_t *gidx)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    const int64_t row = blockIdx.x;
    uint64_t *sk;
    int32_t *si;
    if (gkey) { sk = gkey + row * m; si = gidx + row * m; }
    else { sk = reinterpret_cast<uint64_t *>(smem_raw); si = reinterpret_cast<int32_t *>(sk + m); }
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += FIN_THREADS) {
        if (i < t) { sk[i] = kept_key[row * t + i]; si[i] = kept_idx[row * t + i]; }
        else { sk[i] = 0; si[i] = 0x7fffffff; }
    }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += FIN_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const bool asc = (i & k) == 0;
                    const int32_t a = si[i], b = si[p];
                    if (asc ? (a > b) : (a < b)) {
                        si[i] = b; si[p] = a;
                        const uint64_t tk = sk[i]; sk[i] = sk[p]; sk[p] = tk;
                    }
                }
            }
            __syncthreads();
        }
    }
    double s = 0.0;
    for (int i = tid; i < t; i += FIN_THREADS) {
        breach[row * t + i] = si[i];
        s += fkey_inv(sk[i]);
    }
    s = block_sum(s, red);
    if (tid == 0) cvar[row] = s / (double)t;
}

// ============================================================ host orchestration
int risk_value(mcp_ctx *ctx, int64_t p0, int64_t p1, double *d_V)
{
    if (p1 <= p0) return MCP_OK;
    dim3 grid((unsigned)((ctx->N + VT - 1) / VT), (unsigned)((p1 - p0 + VT - 1) / VT));
    KScope ks(ctx, MCP_K_VALUE);
    k_value_generic<false><<<grid, 256, 0, ctx->stream>>>(ctx->dE.as<double>(), ctx->dS.as<double>(), p0, p1, ctx->N,
                                                         ctx->F, d_V);
    return check_launch(ctx, "k_value_generic");
}

int risk_run(mcp_ctx *ctx, double alpha, int codec_id, int framing)
{
    const int64_t P = ctx->P, N = ctx->N;
    const int64_t t = mcp_tail_size(N, alpha);
    if (P <= 0 || N <= 0) return fail(ctx, MCP_E_STATE, "run: empty problem");
    if (N > 0x7fffffff) return fail(ctx, MCP_E_ARG, "run: more than 2^31-1 trials per context");
    ctx->haveRun = false;
    MCP_CUDA(ctx, ctx->dMean.reserve(P * sizeof(double)));
    MCP_CUDA(ctx, ctx->dVar.reserve(P * sizeof(double)));
    MCP_CUDA(ctx, ctx->dVarLoss.reserve(P * sizeof(double)));
    MCP_CUDA(ctx, ctx->dVarTrial.reserve(P * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dCvar.reserve(P * sizeof(double)));
    MCP_CUDA(ctx, ctx->dBreach.reserve((size_t)P * t * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dKeptKey.reserve((size_t)P * t * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dKeptIdx.reserve((size_t)P * t * sizeof(int32_t)));

    // dense loss tile budget: a quarter of HBM, at most 32 GiB
    size_t budget = ctx->prop.totalGlobalMem / 4;
    if (budget > ((size_t)32 << 30)) budget = (size_t)32 << 30;
    int64_t Pt = (int64_t)(budget / ((size_t)N * sizeof(double)));
    if (Pt < 1) return fail(ctx, MCP_E_NOMEM, "run: one loss row does not fit the tile budget");
    if (Pt > P) Pt = P;
    if (Pt > 65535 * (int64_t)VT) Pt = 65535 * (int64_t)VT;
    MCP_CUDA(ctx, ctx->dLoss.reserve((size_t)Pt * N * sizeof(double)));

    int m = 1;
    while (m < t) m <<= 1;
    const size_t fin_smem = (size_t)m * 12;
    const bool fin_global = fin_smem > 200 * 1024;
    if (fin_global) {
        MCP_CUDA(ctx, ctx->dCandKey.reserve((size_t)Pt * m * sizeof(uint64_t)));
        MCP_CUDA(ctx, ctx->dCandIdx.reserve((size_t)Pt * m * sizeof(int32_t)));
    } else if (fin_smem > 48 * 1024) {
        MCP_CUDA(ctx, cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    }

    MCP_CUDA(ctx, cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEL_SMEM));

    for (int64_t p0 = 0; p0 < P; p0 += Pt) {
        const int64_t p1 = (p0 + Pt < P) ? p0 + Pt : P;
        {
            dim3 grid((unsigned)((N + VT - 1) / VT), (unsigned)((p1 - p0 + VT - 1) / VT));
            KScope ks(ctx, MCP_K_VALUE);
            k_value_generic<true><<<grid, 256, 0, ctx->stream>>>(ctx->dE.as<double>(), ctx->dS.as<double>(), p0, p1, N,
                                                                ctx->F, ctx->dLoss.as<double>());
            MCP_TRY(check_launch(ctx, "k_value_generic"));
        }
        SelArgs a{};
        a.loss = ctx->dLoss.as<double>();
        a.idx = nullptr;
        a.row_stride = N;
        a.count = nullptr;
        a.n = N;
        a.idx_base = 0;
        a.tail = t;
        a.moments_n = N;
        a.kept_key = ctx->dKeptKey.as<uint64_t>() + p0 * t;
        a.kept_idx = ctx->dKeptIdx.as<int32_t>() + p0 * t;
        a.mean = ctx->dMean.as<double>() + p0;
        a.variance = ctx->dVar.as<double>() + p0;
        a.var_loss = ctx->dVarLoss.as<double>() + p0;
        a.var_trial = ctx->dVarTrial.as<int32_t>() + p0;
        {
            KScope ks(ctx, MCP_K_SELECT);
            k_select<<<(unsigned)(p1 - p0), SEL_THREADS, SEL_SMEM, ctx->stream>>>(a);
            MCP_TRY(check_launch(ctx, "k_select"));
        }
        {
            KScope ks(ctx, MCP_K_SELECT);
            k_finalize<<<(unsigned)(p1 - p0), FIN_THREADS, fin_global ? 0 : fin_smem, ctx->stream>>>(
                a.kept_key, a.kept_idx, t, m, ctx->dBreach.as<int32_t>() + p0 * t, ctx->dCvar.as<double>() + p0,
                fin_global ? ctx->dCandKey.as<uint64_t>() : nullptr, fin_global ? ctx->dCandIdx.as<int32_t>() : nullptr);
            MCP_TRY(check_launch(ctx, "k_finalize"));
        }
    }
    // ---- encode the breach lists (uniform length t) into the JavaFastPFOR format
    MCP_CUDA(ctx, ctx->dEncOff.reserve((size_t)(P + 1) * sizeof(int64_t)));
    int64_t total = 0;
    MCP_TRY(codec_encode_plan(ctx, codec_id, framing, ctx->dBreach.as<int32_t>(), nullptr, t, P, ctx->dEncOff.as<int64_t>(), &total));
    MCP_CUDA(ctx, ctx->dEnc.reserve((size_t)total + 16));
    MCP_TRY(codec_encode_emit(ctx, codec_id, framing, ctx->dBreach.as<int32_t>(), nullptr, t, P, ctx->dEncOff.as<int64_t>(),
                              ctx->dEnc.as<uint8_t>()));
    ctx->runP = P;
    ctx->runN = N;
    ctx->runTail = t;
    ctx->runEncBytes = total;
    ctx->haveRun = true;
    return MCP_OK;
}

} // namespace mcp
