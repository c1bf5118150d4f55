// risk_kernels.cu -- kernels (1)-(3) of the hot path: portfolio x scenario valuation,
// per-portfolio moments + exact tail selection (VaR order statistic with its trial index,
// CVaR), and the sorted breach-trial list.  DESIGN.md section 3 is the specification
// (the reference has no code for these stages); the oracle is oracle/risk_oracle.c.
//
// Bit-exactness contract: V[p][n] is ONE fma chain over k ascending starting from +0.0,
// on both sides, so every loss is bit-identical to the oracle's and the selected trial
// indices are exact.
#include <algorithm>
#include <vector>

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

// ============================================================ ingest: positions -> exposure matrix
// E[p][f] = fma chain, in file order, of weight_j * B[instrument_j][f] over the portfolio's (instrument, weight)
// positions (LoadPortfolios.java:352-379 reads them; DESIGN.md 3.0 defines the product; oracle: orc_build_exposures).
// One warp per portfolio: 32 positions are fetched coalesced and handed round by shuffle, lane l owns factors l, l+32, ...
// The loading rows it gathers (F doubles each, 2.5 MB table for the shipped sample) stay in L2.
__global__ void __launch_bounds__(256) k_build_exposures(const int32_t *__restrict__ inst, const double *__restrict__ w,
                                                         const int64_t *__restrict__ off, int64_t P, const double *__restrict__ B,
                                                         int64_t n_inst, int F, double *__restrict__ E, int *err)
{
    const int lane = threadIdx.x & 31;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    const int64_t j0 = off[p], j1 = off[p + 1];
    for (int f0 = 0; f0 < F; f0 += 64) {
        const int fa = f0 + lane, fb = f0 + 32 + lane;
        double acc_a = 0.0, acc_b = 0.0;
        for (int64_t base = j0; base < j1; base += 32) {
            const int64_t j = base + lane;
            int32_t my_i = 0;
            double my_w = 0.0;
            if (j < j1) { my_i = inst[j]; my_w = w[j]; }
            if (j < j1 && (my_i < 0 || my_i >= n_inst)) { *err = 1; my_i = 0; my_w = 0.0; }
            const int cnt = (int)min((int64_t)32, j1 - base);
            for (int k = 0; k < cnt; ++k) {
                const int32_t ii = __shfl_sync(kFullMask, my_i, k);
                const double ww = __shfl_sync(kFullMask, my_w, k);
                const double *row = B + (int64_t)ii * F;
                if (fa < F) acc_a = fma(ww, __ldg(row + fa), acc_a);
                if (fb < F) acc_b = fma(ww, __ldg(row + fb), acc_b);
            }
        }
        if (fa < F) E[p * F + fa] = acc_a;
        if (fb < F) E[p * F + fb] = acc_b;
    }
}

int build_exposures(mcp_ctx *ctx, const int32_t *d_inst, const double *d_w, const int64_t *d_off, int64_t P, const double *d_B,
                    int64_t n_inst, int32_t F, double *d_E)
{
    if (P == 0) return MCP_OK;
    MCP_CUDA(ctx, ctx->dErr.reserve(64));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dErr.p, 0, sizeof(int), ctx->stream));
    {
        KScope ks(ctx, MCP_K_GEN);
        k_build_exposures<<<(unsigned)((P * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_inst, d_w, d_off, P, d_B, n_inst, F, d_E, ctx->dErr.as<int>());
        MCP_TRY(check_launch(ctx, "k_build_exposures"));
    }
    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*reinterpret_cast<int *>(ctx->hPinned) != 0) return fail(ctx, MCP_E_ARG, "build_exposures: instrument id outside the loading table");
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

// Same question for the FP64 TENSOR path the valuation kernel actually issues: 8 independent accumulator pairs per
// warp of mma.sync.m8n8k4.f64 (DMMA), no memory traffic.  512 FMA per instruction per warp.
__global__ void __launch_bounds__(1024) k_fp64_mma_peak(double *out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345e-300) out[0] = s; // never true; keeps the chains alive
}

int fp64_mma_peak(mcp_ctx *ctx, double *tflops)
{
    MCP_CUDA(ctx, ctx->dErr.reserve(64));
    const int grid = ctx->prop.multiProcessorCount, iters = 4096; // one CTA per SM; the best of 1, 2, 4 and 8 warps per sub-partition
    cudaEvent_t a, b;
    MCP_CUDA(ctx, cudaEventCreate(&a));
    MCP_CUDA(ctx, cudaEventCreate(&b));
    double best = 0;
    const int threads[4] = {128, 256, 512, 1024};
    for (int cfg = 0; cfg < 4; ++cfg)
        for (int rep = 0; rep < 3; ++rep) {
            KScope ks(ctx, MCP_K_GEN);
            cudaEventRecord(a, ctx->stream);
            k_fp64_mma_peak<<<grid, threads[cfg], 0, ctx->stream>>>(ctx->dErr.as<double>(), iters, 0.999999, 1e-9);
            cudaEventRecord(b, ctx->stream);
            MCP_CUDA(ctx, cudaEventSynchronize(b));
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            // per warp and instruction: 8 x 8 outputs x 4 k = 256 FMA = 512 flop; 32 instructions per iteration
            const double flops = 512.0 * 32.0 * iters * (threads[cfg] / 32.0) * grid;
            const double tf = flops / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best) best = tf;
        }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return check_launch(ctx, "k_fp64_mma_peak");
}

// ============================================================ valuation (generic F)
// V tile = 64 portfolios x 64 trials per CTA, 4x4 register micro-tile per thread, k staged
// through shared memory in slabs of 32.  Every (p,n) accumulator is a single fma chain in
// ascending k.  Writes V row-major [p - p0][N].  Used for factor counts
// without a specialised streaming kernel and for mcp_value().
constexpr int VT = 64, VK = 32;

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
            if (n < N) out[(p - p0) * N + n] = acc[i][j];
        }
    }
}

// ============================================================ valuation, streaming, FP64 tensor cores
// Kernels (1)+(2a)+(3a) fused, for factor counts F = 4*NK in {16, 32, 48, 64}.
//
// The contraction V = E x S^T runs on the FP64 tensor-core path: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4),
// 8 portfolios x 8 trials x 4 factors per instruction.  On B200 its result is BIT-IDENTICAL to the
// k-ascending chain of IEEE fma that DESIGN.md 3.1 specifies (profiles/microbench/dmma_probe.cu: 0
// mismatches in 128 000 outputs on inputs spanning 2^-20..2^20; the parity tests re-check it on every run),
// it peaks at 37.1 TFLOP/s here against 34.0 for scalar DFMA, and -- the reason it wins -- it needs one
// shared-memory operand fetch per 256 fma instead of one per 64: the scalar DFMA version of this kernel
// (git history) was shared-memory-bandwidth bound at 48 % of the FP64 peak.  (tcgen05 has no FP64 kind.)
//
// "E-stationary": a warp keeps the A fragments of VM_G portfolio groups (two up to F = 48: 16 portfolios x all F factors; four
// at F = 64) in registers and streams trials past them; a CTA = 8 warps = 128 (256) portfolios x one trial sub-chunk.  The
// scenario rows of the sub-chunk arrive in B-FRAGMENT ORDER (k_repack_scenarios), so a stage of 128 trials is ONE contiguous
// block: a single cp.async.bulk (TMA bulk copy, SASS UBLKCP) by one elected thread, completing on mbarriers, 2 stages deep, and
// every B-fragment load reads 32 consecutive doubles -- bank-conflict free without padding.  Fused on each value while
// it is still in a register:
//   * moments (MOMENTS variant only): d = V - c_p (c_p = V[p][trial 0]), s1 += d, s2 = fma(d,d,s2)
//   * tail filter: DENSE steps store every V; SPARSE steps append (V, trial) only when V is at or beyond the
//     portfolio's current bound: an integer compare on the raw bit pattern (its high word, when every bound of the
//     warp is negative) and predicated stores -- no divergence, no FP64-pipe slots.
// Each lane appends to its own private segment (portfolio, sub-chunk, lane%4): no atomics.
constexpr int VM_WARPS = 8;
constexpr int VM_THREADS = VM_WARPS * 32;
// portfolio groups (of 8) per warp: two up to F = 48 (106 registers: two CTAs per SM), FOUR at F = 64 -- there one CTA per SM is all
// the 128 KB of stages allow anyway, and four accumulator chains per warp and half the B-fragment loads per DMMA took C3 from
// 5.05 to 4.84 ms per 125 000 rows (234 registers).  (Four groups at F = 32 means one 8-warp CTA per SM: 2.40 instead of 2.03 ms.)
template <int NK> constexpr int vm_groups() { return NK >= 16 ? 4 : 2; }
static inline int vm_pg(int F) { return VM_WARPS * (F >= 64 ? 4 : 2) * 8; } // portfolios per CTA: 128, or 256 at F = 64
// Scenario staging: TWO stages of 128 trials (64 KB per CTA at F = 32, 128 KB at F = 64).  Measured on C2 with 16-warp CTAs:
// 4 stages x 32 trials 2.157 ms, 4 x 64 2.102, 3 x 128 2.068, 2 x 256 2.056 -- what costs is the per-stage handshake (mbarrier
// wait / arrive, the producer thread's detour), not the prefetch depth.  Round 2: 8-warp CTAs with 2 x 128-trial stages (two
// such CTAs share an SM up to F = 48: the same 16 warps, but independent of each other) 2.052 ms, and C3 (F = 64, one 8-warp
// CTA per SM) 5.11 instead of 5.16 ms per 125 000 rows: eight warps already keep the DMMA pipe as busy as sixteen.
constexpr int VM_NSTAGE = 2;
template <int NK> constexpr int vm_stage_trials() { return 128; }

struct StreamArgs {
    const double *E, *S;
    const double *Sfrag;     // scenarios repacked in B-fragment order: [trial block of 8][k-step][lane] (k_repack_scenarios)
    int64_t p0, p1;          // portfolio tile
    int64_t n_begin, n_end;  // trial range of this step
    int nsub;                // sub-chunks of the step
    int sub0;                // first sub-chunk of this launch (gridDim.x of them): a step may be launched in parts
    int sub_len;             // trials per sub-chunk (multiple of 8)
    const uint64_t *tkey;    // [P] candidate iff fkey(V) <= tkey, i.e. V <= -(current VaR lower bound)  (SPARSE)
    void *cand;              // candidate buffer, one row of `pitch` bytes per portfolio.  DENSE steps use the row as a
                             // plain double array (slot = trial - n_begin); SPARSE steps append 16-byte records {V, trial}
    int64_t pitch;
    int32_t *cnt;            // [p - p0][4*nsub]   (SPARSE)
    double *part;            // [P][part_stride][2]; null = no moments (mcp_value)
    int part_stride, part_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// D(8x8) += A(8x4) * B(4x8), FP64.  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

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

// candidate record of a sparse step: 16 bytes so that one predicated STG.128 appends it
struct __align__(16) CandRec { double v; int32_t trial; int32_t pad; };

// branch-free conditional append of one candidate record to a lane-private segment
__device__ __forceinline__ void append_if(uint32_t hit, CandRec *dst, double v, int32_t n)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 t;\n\tsetp.ne.u32 p, %0, 0;\n\tcvt.u64.u32 t, %3;\n\t@p st.global.v2.b64 [%1], {%2, t};\n\t}" ::"r"(hit),
                 "l"(dst), "l"(__double_as_longlong(v)), "r"(n) : "memory");
}

template <int NK>
constexpr size_t vm_smem_bytes() { return sizeof(double) * VM_NSTAGE * vm_stage_trials<NK>() * (NK * 4); }

// S[N][F] -> Sfrag[ceil(N/8)][NK][32]: element (b, s, lane) = S[8b + lane/4][4s + lane%4] (0 past the last trial).
// In this order a stage of 32 trials is one contiguous block (a single TMA bulk copy) and every B-fragment load is
// 32 consecutive doubles: bank-conflict free without padding.
__global__ void k_repack_scenarios(const double *__restrict__ S, int64_t N, int NK, double *__restrict__ Sfrag, int64_t blk0, int64_t blk1)
{
    const int64_t total = blk1 * NK * 32; // trial blocks [blk0, blk1)
    const int F = NK * 4;
    for (int64_t i = blk0 * NK * 32 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int lane = (int)(i & 31);
        const int64_t bs = i >> 5;
        const int sidx = (int)(bs % NK);
        const int64_t b = bs / NK;
        const int64_t n = 8 * b + (lane >> 2);
        Sfrag[i] = n < N ? S[n * F + 4 * sidx + (lane & 3)] : 0.0;
    }
}

template <int NK, bool DENSE, bool MOMENTS>
__global__ void __launch_bounds__(VM_THREADS, 1) k_value_mma(const StreamArgs a)
{
    constexpr int F = NK * 4;
    constexpr int VM_G = vm_groups<NK>();
    constexpr int VM_PG = VM_WARPS * VM_G * 8;
    constexpr int BLK = NK * 32;            // doubles per 8-trial block in fragment order
    constexpr int VM_ST = vm_stage_trials<NK>(); // trials per stage
    constexpr int STG = (VM_ST / 8) * BLK;       // doubles per stage
    extern __shared__ __align__(128) unsigned char vm_smem[];
    double *s_S = reinterpret_cast<double *>(vm_smem); // [NSTAGE][ST/8][NK][32]
    __shared__ __align__(8) uint64_t s_full[VM_NSTAGE], s_empty[VM_NSTAGE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 2, c = lane & 3;
    const int sub = blockIdx.x + a.sub0;
    const int64_t nb = a.n_begin + (int64_t)sub * a.sub_len;
    const int64_t ne = min(a.n_end, nb + a.sub_len);
    const int ntr = (int)(ne - nb);
    const int nst = (ntr + VM_ST - 1) / VM_ST;
    if (threadIdx.x == 0) {
        for (int s = 0; s < VM_NSTAGE; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], VM_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Producer duty: thread 0 keeps the stage ring full (prefetch distance NSTAGE-1 stages) with ONE bulk copy per
    // stage: in fragment order the stage's trial blocks are contiguous (nb is a multiple of 8).
    auto produce = [&](int it) {
        const int st = it % VM_NSTAGE;
        if (it >= VM_NSTAGE) mbar_wait(&s_empty[st], ((it / VM_NSTAGE) - 1) & 1);
        const int rows = min(VM_ST, ntr - it * VM_ST);
        const uint32_t bytes = (uint32_t)((rows + 7) >> 3) * BLK * sizeof(double);
        mbar_expect_tx(&s_full[st], bytes);
        bulk_g2s(s_S + (size_t)st * STG, a.Sfrag + ((nb >> 3) + (int64_t)it * (VM_ST / 8)) * BLK, bytes, &s_full[st]);
    };
    if (threadIdx.x == 0)
        for (int it = 0; it < VM_NSTAGE - 1 && it < nst; ++it) produce(it);

    // ---- A fragments: this warp's 4 groups x 8 portfolios, all factors, resident in registers
    const int64_t pw = a.p0 + (int64_t)blockIdx.y * VM_PG + warp * (VM_G * 8);
    double ef[VM_G][NK];
    bool ok[VM_G];
    int64_t prow[VM_G];
#pragma unroll
    for (int g = 0; g < VM_G; ++g) {
        const int64_t p = pw + g * 8 + r;
        ok[g] = p < a.p1;
        prow[g] = p;
        const double *er = a.E + (ok[g] ? p : a.p0) * F + c;
#pragma unroll
        for (int s = 0; s < NK; ++s) ef[g][s] = __ldg(er + 4 * s);
    }
    // ---- shift c_p = V[p][trial 0]: one DMMA pass over scenario rows 0..7, column 0 broadcast to the row's 4 lanes
    double cs[VM_G];
    if (MOMENTS) {
        double z0[VM_G], z1[VM_G];
#pragma unroll
        for (int g = 0; g < VM_G; ++g) { z0[g] = 0.0; z1[g] = 0.0; }
        const int64_t Ntot = a.n_end; // rows 0..7 exist whenever N >= 8; clamp otherwise
#pragma unroll
        for (int s = 0; s < NK; ++s) {
            const double b = __ldg(a.S + (int64_t)min((int64_t)r, Ntot - 1) * F + 4 * s + c);
#pragma unroll
            for (int g = 0; g < VM_G; ++g) dmma884(z0[g], z1[g], ef[g][s], b);
        }
#pragma unroll
        for (int g = 0; g < VM_G; ++g) cs[g] = __shfl_sync(0xffffffffu, z0[g], lane & ~3);
    } else {
#pragma unroll
        for (int g = 0; g < VM_G; ++g) cs[g] = 0.0;
    }
    uint64_t tk[VM_G];   // generic compare: candidate iff fkey(V) <= tk
    uint64_t traw[VM_G]; // fast compare (negative bound, the normal case): candidate iff bits(V) >= traw, unsigned
    uint32_t thi[VM_G];  // ... tested on the high words only
    int cnt[VM_G];
    double s1[VM_G], s2[VM_G];
    double *drow[VM_G];   // DENSE: the portfolio's row as doubles, at this sub-chunk
    CandRec *srec[VM_G];  // SPARSE: this lane's private segment of records
    bool neg = true;
#pragma unroll
    for (int g = 0; g < VM_G; ++g) {
        tk[g] = 0; // fkey(x) <= 0 never holds for a finite x: inactive rows never append
        traw[g] = ~0ull;
        thi[g] = 0xffffffffu;
        if (!DENSE && ok[g]) {
            tk[g] = a.tkey[prow[g]];
            const double bound = fkey_inv(tk[g]);
            traw[g] = (uint64_t)__double_as_longlong(bound);
            thi[g] = (uint32_t)(traw[g] >> 32);
            neg = neg && (bound < 0.0);
        }
        cnt[g] = 0;
        s1[g] = 0.0;
        s2[g] = 0.0;
        unsigned char *rowp = reinterpret_cast<unsigned char *>(a.cand) + (prow[g] - a.p0) * a.pitch;
        drow[g] = reinterpret_cast<double *>(rowp) + (int64_t)sub * a.sub_len;
        srec[g] = reinterpret_cast<CandRec *>(rowp) + (int64_t)sub * a.sub_len + (int64_t)c * (a.sub_len >> 2);
    }
    // all 32 portfolios of the warp have a negative bound on V (a positive VaR): one unsigned compare per value
    const bool fastcmp = __all_sync(0xffffffffu, neg);
    const bool vec_ok = (a.sub_len & 1) == 0 && (a.pitch & 15) == 0;
    constexpr bool moments = MOMENTS; // compile-time: even predicated-off scalar FP64 instructions cost DMMA-pipe slots
    // Scheduling.  The FP64 tensor pipe is the bound resource and scalar FP64 ops (the moments) share it, while
    // integer work co-issues with DMMA (profiles/microbench/dmma_coissue.cu).  Four warps per SM sub-partition give
    // the scheduler another warp's integer epilogue to run while one warp's DMMAs occupy the pipe (an explicit
    // two-warp ping-pong through named barriers was measured slower: the barrier round trips idled the pipe).  The
    // epilogue is software-pipelined by one block: the moments of block i-1 are interleaved between the k-steps of
    // block i's DMMAs (no dependency bubble), its filter runs after them.
    double pd0[VM_G], pd1[VM_G]; // previous block's values; initialised to the shift so that they add exactly 0
#pragma unroll
    for (int g = 0; g < VM_G; ++g) { pd0[g] = cs[g]; pd1[g] = cs[g]; }
    int ptl = 0;          // previous block: offset of this lane's first trial inside the sub-chunk
    uint32_t pv = 0;      // previous block: bit0 / bit1 = first / second trial is real

    // integer half of the epilogue for the block held in pd0/pd1
    auto int_epilogue = [&]() {
        const uint32_t v0 = pv & 1u, v1 = pv >> 1;
        const int32_t n0 = (int32_t)(nb + ptl);
#pragma unroll
        for (int g = 0; g < VM_G; ++g) {
            if (DENSE) {
                if (ok[g]) {
                    double *dst = drow[g] + ptl;
                    if (v1 && vec_ok) *reinterpret_cast<double2 *>(dst) = make_double2(pd0[g], pd1[g]);
                    else {
                        if (v0) dst[0] = pd0[g];
                        if (v1) dst[1] = pd1[g];
                    }
                }
            } else {
                uint32_t h0, h1;
                if (fastcmp) {
                    // high words only: one 32-bit compare per value.  A superset of the exact test (values whose high
                    // word ties with the bound's pass regardless of the low word, a 2^-20 relative band): the filter
                    // only has to be conservative, the selection is exact.
                    h0 = ((uint32_t)__double2hiint(pd0[g]) >= thi[g]) & v0;
                    h1 = ((uint32_t)__double2hiint(pd1[g]) >= thi[g]) & v1;
                } else {
                    h0 = (fkey(pd0[g]) <= tk[g]) & v0;
                    h1 = (fkey(pd1[g]) <= tk[g]) & v1;
                }
                append_if(h0, srec[g] + cnt[g], pd0[g], n0);
                cnt[g] += h0;
                append_if(h1, srec[g] + cnt[g], pd1[g], n0 + 1);
                cnt[g] += h1;
            }
        }
    };
    // FP64 half for value v (0 .. 2 VM_G - 1) of the block held in pd0/pd1
    auto moment_of = [&](int v) {
        const int g = v >> 1;
        const double d = ((v & 1) ? pd1[g] : pd0[g]) - cs[g];
        s1[g] += d;
        s2[g] = fma(d, d, s2[g]);
    };

    for (int it = 0; it < nst; ++it) {
        const int st = it % VM_NSTAGE;
        if (threadIdx.x == 0 && it + VM_NSTAGE - 1 < nst) produce(it + VM_NSTAGE - 1);
        __syncwarp();
        mbar_wait(&s_full[st], (it / VM_NSTAGE) & 1);
        const int rows = min(VM_ST, ntr - it * VM_ST);
        const double *sbase = s_S + (size_t)st * STG;
#pragma unroll 1 // (unrolling by 2 or 4 was measured within noise: the loop is pipe-bound, not issue-bound)
        for (int blk = 0; blk * 8 < rows; ++blk) {
            const double *bp = sbase + blk * BLK + lane;
            double sf[NK];
#pragma unroll
            for (int s = 0; s < NK; ++s) sf[s] = bp[32 * s];
            double d0[VM_G], d1[VM_G];
#pragma unroll
            for (int g = 0; g < VM_G; ++g) { d0[g] = 0.0; d1[g] = 0.0; }
#pragma unroll
            for (int s = 0; s < NK; ++s) {
#pragma unroll
                for (int g = 0; g < VM_G; ++g) dmma884(d0[g], d1[g], ef[g][s], sf[s]);
                if (moments) {
#pragma unroll
                    for (int v = 0; v < 2 * VM_G; ++v)
                        if ((v * NK) / (2 * VM_G) == s) moment_of(v);
                }
            }
            int_epilogue();                    // previous block, integer pipe only
            // rotate: this block becomes "previous"
            ptl = it * VM_ST + blk * 8 + 2 * c;
            if (it * VM_ST + blk * 8 + 8 <= ntr) { // whole block is real (warp-uniform)
                pv = 3u;
#pragma unroll
                for (int g = 0; g < VM_G; ++g) { pd0[g] = d0[g]; pd1[g] = d1[g]; }
            } else {
                pv = (ptl < ntr ? 1u : 0u) | (ptl + 1 < ntr ? 2u : 0u);
#pragma unroll
                for (int g = 0; g < VM_G; ++g) {
                    pd0[g] = (pv & 1u) ? d0[g] : cs[g]; // trials past the end of the sub-chunk contribute exactly 0
                    pd1[g] = (pv & 2u) ? d1[g] : cs[g];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);
    }
    // drain the software pipeline: the last block
    if (moments) {
#pragma unroll
        for (int v = 0; v < 2 * VM_G; ++v) moment_of(v);
    }
    int_epilogue();
    // ---- epilogue: per-portfolio partial moments (fixed 4-lane tree: deterministic) and segment counts
#pragma unroll
    for (int g = 0; g < VM_G; ++g) {
        if (MOMENTS) {
            double x = s1[g], y = s2[g];
            x += __shfl_xor_sync(0xffffffffu, x, 1); y += __shfl_xor_sync(0xffffffffu, y, 1);
            x += __shfl_xor_sync(0xffffffffu, x, 2); y += __shfl_xor_sync(0xffffffffu, y, 2);
            if (c == 0 && ok[g]) {
                double *q = a.part + ((int64_t)prow[g] * a.part_stride + a.part_base + sub) * 2;
                q[0] = x; q[1] = y;
            }
        }
        if (!DENSE && ok[g]) a.cnt[(prow[g] - a.p0) * (4 * (int64_t)a.nsub) + 4 * sub + c] = cnt[g];
    }
}

// mean / population variance from the per-sub-chunk partial sums: one warp per portfolio, lane-strided
// partials then a fixed shuffle tree, so the result does not depend on scheduling (deterministic)
__global__ void __launch_bounds__(256) k_moments_finalize(const double *E, const double *S, int64_t P, int32_t F, int64_t N,
                                                          const double *part, int part_stride, int nparts, double *mean,
                                                          double *variance)
{
    const int lane = threadIdx.x & 31;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    double s1 = 0.0, s2 = 0.0;
    const double2 *q = reinterpret_cast<const double2 *>(part) + p * part_stride;
    for (int j = lane; j < nparts; j += 32) {
        const double2 v = q[j];
        s1 += v.x;
        s2 += v.y;
    }
    for (int d = 16; d; d >>= 1) {
        s1 += __shfl_down_sync(kFullMask, s1, d);
        s2 += __shfl_down_sync(kFullMask, s2, d);
    }
    if (lane == 0) {
        double c = 0.0;
        for (int k = 0; k < F; ++k) c = fma(E[p * F + k], S[k], c);
        const double invn = 1.0 / (double)N, m1 = s1 * invn;
        mean[p] = c + m1;
        variance[p] = fma(-m1, m1, s2 * invn);
    }
}

// ============================================================ moments by linearity
// mean_p = E_p . mean(S) and var_p = E_p^T Cov(S) E_p are algebraically the mean and population variance of the
// portfolio's N valuations (V = E S^T is linear), and cost one Gram-matrix pass over S plus F^2 fma per portfolio
// instead of 3 scalar FP64 instructions per valuation.  That matters here because scalar FP64 instructions share
// the DMMA pipe and cost a full tensor-op slot each when interleaved with DMMAs (ncu: the fused per-value moment
// updates took 30 % of the valuation kernel).  The fused per-trial reduction stays available (option
// "fused_moments") and both are held to 1e-9 against the oracle's long-double per-trial reduction.
constexpr int GM_THREADS = 256;
constexpr int GM_ROWS = 256;  // scenario rows per CTA

// column sums of S over this CTA's rows -> partial[b][F]; row-major reads stay coalesced
__global__ void __launch_bounds__(GM_THREADS) k_col_sums(const double *__restrict__ S, int64_t N, int F, double *__restrict__ partial)
{
    __shared__ double red[GM_THREADS];
    const int64_t r0 = (int64_t)blockIdx.x * GM_ROWS, r1 = min(N, r0 + GM_ROWS);
    const int per = GM_THREADS / F;      // row lanes
    const int k = threadIdx.x % F, j = threadIdx.x / F;
    double acc = 0.0;
    if (j < per)
        for (int64_t r = r0 + j; r < r1; r += per) acc += S[r * F + k];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < F) {
        double t = 0.0;
        for (int q = 0; q < per; ++q) t += red[q * F + threadIdx.x];
        partial[(int64_t)blockIdx.x * F + threadIdx.x] = t;
    }
}

// mu[k] = sum of partials / N (fixed order)
__global__ void k_col_mean(const double *partial, int nb, int F, int64_t N, double *mu)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F) return;
    double acc = 0.0;
    for (int b = 0; b < nb; ++b) acc += partial[(int64_t)b * F + k];
    mu[k] = acc / (double)N;
}

// centered Gram matrix of this CTA's rows: partial[b][F][F] = sum_r (S_r - mu)(S_r - mu)^T
__global__ void __launch_bounds__(GM_THREADS) k_gram_centered(const double *__restrict__ S, int64_t N, int F, const double *__restrict__ mu,
                                                              double *__restrict__ partial)
{
    extern __shared__ double gm_s[]; // [32][F] rows being processed
    const int64_t r0 = (int64_t)blockIdx.x * GM_ROWS, r1 = min(N, r0 + GM_ROWS);
    const int FF = F * F;
    // each thread owns entries e = tid, tid + 256, ... (at most 16 for F = 64)
    double acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0;
    for (int64_t rb = r0; rb < r1; rb += 32) {
        const int nr = (int)min((int64_t)32, r1 - rb);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * F; i += GM_THREADS) gm_s[i] = S[rb * F + i] - mu[i % F];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int e = threadIdx.x + q * GM_THREADS;
            if (e < FF) {
                const int k = e / F, l = e % F;
                double a = acc[q];
                for (int r = 0; r < nr; ++r) a = fma(gm_s[r * F + k], gm_s[r * F + l], a);
                acc[q] = a;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int e = threadIdx.x + q * GM_THREADS;
        if (e < FF) partial[(int64_t)blockIdx.x * FF + e] = acc[q];
    }
}

// cov[e] = sum of partials / N (fixed order)
__global__ void k_gram_reduce(const double *partial, int nb, int FF, int64_t N, double *cov)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= FF) return;
    double acc = 0.0;
    for (int b = 0; b < nb; ++b) acc += partial[(int64_t)b * FF + e];
    cov[e] = acc / (double)N;
}

// one warp per portfolio: mean = E.mu, variance = E^T Cov E  (Cov and mu staged in shared memory)
__global__ void __launch_bounds__(256) k_moments_linear(const double *__restrict__ E, int64_t P, int F, const double *__restrict__ mu,
                                                        const double *__restrict__ cov, double *mean, double *variance)
{
    extern __shared__ double ml_s[]; // [F*F + F]
    for (int i = threadIdx.x; i < F * F; i += 256) ml_s[i] = cov[i];
    for (int i = threadIdx.x; i < F; i += 256) ml_s[F * F + i] = mu[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < P; p += (int64_t)gridDim.x * 8) {
        const double *e = E + p * F;
        double m = 0.0, v = 0.0;
        for (int k = lane; k < F; k += 32) {
            const double ek = e[k];
            m = fma(ek, ml_s[F * F + k], m);
            double row = 0.0;
            for (int l = 0; l < F; ++l) row = fma(ml_s[l * F + k], e[l], row); // Cov is symmetric: column read = conflict-free
            v = fma(ek, row, v);
        }
        for (int d = 16; d; d >>= 1) {
            m += __shfl_down_sync(kFullMask, m, d);
            v += __shfl_down_sync(kFullMask, v, d);
        }
        if (lane == 0) { mean[p] = m; variance[p] = v; }
    }
}

// ============================================================ selection
typedef unsigned __int128 u128;
__device__ __forceinline__ u128 comp_of(uint64_t key, uint32_t idx) { return ((u128)key << 32) | idx; }

constexpr int SEL_THREADS = 256;
constexpr int SEL_BITS_MAX = 11;
constexpr int SEL_BINS_MAX = 1 << SEL_BITS_MAX; // widest digit of the general path (k_select_t<256>)
constexpr int SEL_SORT = 256;           // keep refining digits until the boundary bucket is at most this big
constexpr int FS_BITS = 10, FS_BINS = 1 << FS_BITS; // staged fast path: bins per range-adaptive pass
constexpr int FS_SORT = 96;                        // ... and the bucket size its rank sort takes
constexpr int kSelMaxSeg = 4096;         // most segments (4 per sub-chunk) a sparse step may have

// Candidates of one portfolio row = the previously kept tail (key, trial) plus this step's
// segments of raw values V (loss = 0.0 - V), one segment per trial sub-chunk.
struct SelArgs {
    const void *cand;          // one row of `pitch` bytes per portfolio: dense rows are double arrays (trial = n_begin +
                               // seg*sub_len + i), sparse rows are CandRec arrays in lane-private segments
    int64_t pitch;
    const int32_t *cnt;        // [rows][nsub]; null -> dense (segment full)
    int nsub;
    int64_t sub_len;
    int64_t step_len;          // dense: number of valid trials in the step
    int64_t n_begin;
    const uint64_t *kept_key_in;
    const int32_t *kept_idx_in;
    int64_t kept_n;            // entries (= row stride) of the kept_*_in rows; 0 on the first step
    int kept_segs;             // > 1: the kept rows come in `kept_segs` blocks (one per rank of a trial-sharded run,
    int64_t kept_seg_bytes;    //   `kept_seg_bytes` apart), each [rows][kept_n / kept_segs]: row r = the blocks' rows kept_row0 + r
    int64_t kept_row0;
    uint64_t *kept_key_out;    // [rows][tail]
    int32_t *kept_idx_out;
    int64_t tail;              // t
    int64_t moments_n;         // > 0: also compute mean/variance from the dense row (generic path)
    double *mean, *variance;   // [rows] (moments_n > 0)
    double *var_loss;          // [rows]
    int32_t *var_trial;
    uint64_t *tkey;            // [rows] fkey(0.0 - var_loss): the next step's filter bound on V
    int cap;                   // shared-memory candidate capacity (entries)
    int *fail_cnt;             // optimistic bound (see risk_run_stream): rows with fewer than `tail` candidates are
    int32_t *fail_rows;        // appended here (row index) and redone by the guaranteed schedule; null otherwise
    int64_t row_base;          // global index of row 0 (for fail_rows)
    const int32_t *rows;       // k_select_t_list: the CTAs walk rows[0 .. *nrows) (the rows a warp-per-row kernel handed back)
    const int *nrows;
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

// entry i of row `row` of the kept tail handed to the selection (one array, or one block per rank)
__device__ __forceinline__ void kept_at(const SelArgs &a, int64_t row, int64_t i, uint64_t &key, uint32_t &idx)
{
    if (a.kept_segs <= 1) {
        key = a.kept_key_in[row * a.kept_n + i];
        idx = (uint32_t)a.kept_idx_in[row * a.kept_n + i];
        return;
    }
    const int64_t m = a.kept_n / a.kept_segs, g = i / m, j = i - g * m, r = a.kept_row0 + row;
    key = *reinterpret_cast<const uint64_t *>(reinterpret_cast<const unsigned char *>(a.kept_key_in) + g * a.kept_seg_bytes + (r * m + j) * 8);
    idx = (uint32_t)*reinterpret_cast<const int32_t *>(reinterpret_cast<const unsigned char *>(a.kept_idx_in) + g * a.kept_seg_bytes + (r * m + j) * 4);
}

// visit every candidate of the row straight from global memory
template <int THREADS, class Fn>
__device__ __forceinline__ void for_each_global(const SelArgs &a, int64_t row, const int *segoff, Fn f)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t i = tid; i < a.kept_n; i += THREADS) {
        uint64_t k; uint32_t id;
        kept_at(a, row, i, k, id);
        f(k, id);
    }
    const unsigned char *rowp = reinterpret_cast<const unsigned char *>(a.cand) + row * a.pitch;
    if (a.cnt) {
        // sparse step: many short lane-private segments of 16-byte records.  Flat index over all their entries
        // (segoff = exclusive prefix of the segment counts, in shared memory): independent loads in flight.
        const int4 *R = reinterpret_cast<const int4 *>(rowp);
        const int total = segoff[a.nsub];
        for (int i = tid; i < total; i += THREADS) {
            int lo = 0, hi = a.nsub; // largest seg with segoff[seg] <= i
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (segoff[mid] <= i) lo = mid; else hi = mid;
            }
            const int4 rec = R[(int64_t)lo * a.sub_len + (i - segoff[lo])];
            const double v = __hiloint2double(rec.y, rec.x);
            f(fkey(0.0 - v), (uint32_t)rec.z);
        }
        return;
    }
    const double *V = reinterpret_cast<const double *>(rowp);
    for (int seg = warp; seg < a.nsub; seg += THREADS / 32) {
        const int64_t base = (int64_t)seg * a.sub_len;
        int64_t c = a.step_len - base;
        c = c < 0 ? 0 : (c > a.sub_len ? a.sub_len : c);
        for (int64_t i = lane; i < c; i += 32) f(fkey(0.0 - V[base + i]), (uint32_t)(a.n_begin + base + i));
    }
}

// One CTA per portfolio row: exact top-`tail` of (loss, trial) by MSB-first radix select
// (kernel (2b)).  The row's candidates are staged ONCE into shared memory when they fit
// (the normal case: a few tail sizes), so the digit passes never go back to HBM; rows with
// more candidates than `cap` (adversarial trial orders, very large tails) run the same
// passes straight from global memory.
// THREADS = 256 for tails of a thousand trials, 64 for small tails (t ~ 100: a 256-thread CTA per row would spend its
// time in barriers); the digit widths follow so that the histogram scans stay one pass per thread.
template <int THREADS>
__device__ void select_t_row(const SelArgs &a, int64_t row);

// one CTA per row ...
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_select_t(const SelArgs a)
{
    select_t_row<THREADS>(a, blockIdx.x);
}
// ... or a grid-stride walk over the listed rows (what a warp-per-row kernel handed back: usually none)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_select_t_list(const SelArgs a)
{
    const int n = *a.nrows;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        select_t_row<THREADS>(a, (int64_t)a.rows[b]);
        __syncthreads();
    }
}

template <int THREADS>
__device__ void select_t_row(const SelArgs &a, const int64_t row)
{
    constexpr int SEL_BITS = THREADS >= 256 ? 11 : 8, SEL_BINS = 1 << SEL_BITS; // general path: digits of the 96-bit composite
    constexpr int FS_BITS = THREADS >= 256 ? 10 : 8, FS_BINS = 1 << FS_BITS;    // staged path: bins per range-adaptive pass
    constexpr int FS_SORT = THREADS >= 256 ? 96 : 64;                           // ... and its rank-sorted bucket (one thread per element)
    extern __shared__ __align__(16) unsigned char sel_smem[];
    // layout: [segoff: nsub + 2 ints, rounded up to 16 bytes] then EITHER the staging area and the staged path's histogram
    // ([cap] keys, [cap] trials, FS_BINS bins) OR the general path's wider histogram (SEL_BINS bins; it stages nothing)
    int *segoff = reinterpret_cast<int *>(sel_smem); // [nsub + 1] exclusive prefix of the segment counts (sparse steps)
    unsigned char *area = sel_smem + (((size_t)(a.nsub + 2) * 4 + 15) & ~(size_t)15);
    uint64_t *skey = reinterpret_cast<uint64_t *>(area);
    uint32_t *sidx = reinterpret_cast<uint32_t *>(skey + a.cap);
    int *hist = reinterpret_cast<int *>(sidx + a.cap);
    __shared__ uint64_t ckey[SEL_SORT];
    __shared__ uint32_t cidx[SEL_SORT];
    __shared__ double red[32];
    __shared__ int s_wsum[THREADS / 32];
    __shared__ int s_sel, s_above, s_bucket, s_ncand, s_nkept;
    const int64_t t = a.tail;
    uint64_t *kk = a.kept_key_out + row * t;
    int32_t *ki = a.kept_idx_out + row * t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // what the row needs first -- the "lost" mark, this thread's entry of the kept tail, its segment count -- is requested
    // before anything is waited for (one global round trip instead of three)
    const bool kflat = a.kept_n > 0 && a.kept_segs <= 1;
    int32_t first = 0, pcnt = 0;
    uint64_t pk = 0;
    uint32_t pi = 0;
    if (kflat) first = a.kept_idx_in[row * a.kept_n];
    const bool have_pk = tid < a.kept_n;
    if (have_pk) kept_at(a, row, tid, pk, pi);
    if (a.cnt && tid < a.nsub) pcnt = a.cnt[row * a.nsub + tid];
    // a row that already lost its optimistic bet at an earlier step stays lost (it is redone separately)
    if (kflat && first < 0) {
        if (tid == 0) ki[0] = -1;
        return;
    }
    // ---- moments of the dense row with shift c = V[0] (generic path only)
    if (a.moments_n > 0) {
        const double *V = reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(a.cand) + row * a.pitch);
        const double c = V[0];
        double s1 = 0.0, s2 = 0.0;
        for (int64_t i = tid; i < a.step_len; i += THREADS) {
            const double d = V[i] - c;
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

    // ---- how many candidates does this row have?  stage them if they fit
    int64_t M;
    if (a.cnt) {
        // exclusive prefix of the segment counts: coalesced load, every thread scans a contiguous run, block scan
        if (tid < a.nsub) segoff[tid + 1] = pcnt;
        for (int sgi = tid + THREADS; sgi < a.nsub; sgi += THREADS) segoff[sgi + 1] = a.cnt[row * a.nsub + sgi];
        if (tid == 0) segoff[0] = 0;
        __syncthreads();
        const int per = (a.nsub + THREADS - 1) / THREADS;
        const int g0 = min(a.nsub, tid * per), g1 = min(a.nsub, g0 + per);
        int c = 0;
        for (int g = g0; g < g1; ++g) c += segoff[g + 1];
        int incl = c;
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int run = incl - c;
        for (int w = 0; w < warp; ++w) run += s_wsum[w];
        for (int g = g0; g < g1; ++g) { run += segoff[g + 1]; segoff[g + 1] = run; } // inclusive at g+1 == exclusive of g+1
        __syncthreads();
        M = a.kept_n + segoff[a.nsub];
    } else
        M = a.kept_n + a.step_len;
    if (M < t) {
        // Only possible under an optimistic filter bound that turned out too tight for this row: report the row
        // (it is redone with the guaranteed schedule) and mark its kept list so that the finalize kernel skips it.
        if (tid == 0) {
            ki[0] = -1;
            if (a.fail_cnt) a.fail_rows[atomicAdd(a.fail_cnt, 1)] = (int32_t)(a.row_base + row);
        }
        return;
    }
    const bool staged = M <= a.cap;
    if (tid == 0) { s_ncand = 0; s_nkept = 0; }
    __syncthreads();
    const int Ms = (int)M;
    if (staged) {
        // ===== fast path: stage the row's candidates ONCE in shared memory (deterministic slots, independent
        // loads), tracking the key range on the way.
        __shared__ uint64_t s_lo, s_hi;
        __shared__ uint64_t s_red[2 * (THREADS / 32)];
        uint64_t mn = ~0ull, mx = 0;
        auto put = [&](int slot, uint64_t k, uint32_t id) {
            skey[slot] = k; sidx[slot] = id;
            mn = k < mn ? k : mn; mx = k > mx ? k : mx;
        };
        const int kn = (int)a.kept_n;
        if (have_pk) put(tid, pk, pi);
        for (int i = tid + THREADS; i < kn; i += THREADS) {
            uint64_t k; uint32_t id;
            kept_at(a, row, i, k, id);
            put(i, k, id);
        }
        const unsigned char *rowp = reinterpret_cast<const unsigned char *>(a.cand) + row * a.pitch;
        if (a.cnt) {
            // sparse step: ~3 records in each of several hundred lane-private segments.  Phase 1: every thread labels
            // the slots of its own segments with the segment id (shared memory only); phase 2: a flat, unrolled loop
            // over the slots issues the 16-byte record loads independently of one another.
            const int per = (a.nsub + THREADS - 1) / THREADS;
            const int g0 = min(a.nsub, tid * per), g1 = min(a.nsub, g0 + per);
            for (int g = g0; g < g1; ++g)
                for (int j = segoff[g]; j < segoff[g + 1]; ++j) sidx[kn + j] = (uint32_t)g;
            __syncthreads();
            const int4 *R = reinterpret_cast<const int4 *>(rowp);
            const int total = segoff[a.nsub];
            // four records per thread and round: all four labels are read and all four loads issued before the first slot is
            // rewritten (labels and results share `sidx`: left to the compiler, every load waited for the store before it)
            for (int i0 = tid; i0 < total; i0 += 4 * THREADS) {
                int4 rec[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * THREADS;
                    if (i < total) {
                        const int g = (int)sidx[kn + i];
                        rec[u] = __ldg(R + (int64_t)g * a.sub_len + (i - segoff[g]));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * THREADS;
                    if (i < total) put(kn + i, fkey(0.0 - __hiloint2double(rec[u].y, rec[u].x)), (uint32_t)rec[u].z);
                }
            }
        } else {
            const double *V = reinterpret_cast<const double *>(rowp);
            const int total = (int)a.step_len; // dense rows are contiguous: slot = trial - n_begin
#pragma unroll 4
            for (int i = tid; i < total; i += THREADS) put(kn + i, fkey(0.0 - __ldg(V + i)), (uint32_t)(a.n_begin + i));
        }
        for (int d = 16; d; d >>= 1) {
            const uint64_t a1 = __shfl_xor_sync(kFullMask, mn, d), b1 = __shfl_xor_sync(kFullMask, mx, d);
            mn = a1 < mn ? a1 : mn; mx = b1 > mx ? b1 : mx;
        }
        if (lane == 0) { s_red[warp] = mn; s_red[THREADS / 32 + warp] = mx; }
        __syncthreads();
        if (tid == 0) {
            uint64_t lo = ~0ull, hi = 0;
            for (int w = 0; w < THREADS / 32; ++w) { lo = s_red[w] < lo ? s_red[w] : lo; hi = s_red[THREADS / 32 + w] > hi ? s_red[THREADS / 32 + w] : hi; }
            s_lo = lo; s_hi = hi;
        }
        __syncthreads();
        // Range-adaptive radix select: histogram the keys on FS_BINS equal-width bins of the CURRENT [lo, hi] range
        // (not on raw bit fields), so one pass usually isolates a boundary bin of a handful of elements; repeat on
        // that bin's range until it holds <= FS_SORT.
        int64_t need = t;
        int bucket = 0;
        bool on_idx = false;      // ties: more than FS_SORT candidates share one loss value -> continue on the trial index
        uint64_t kstar = 0;       // that loss key
        uint64_t lo = s_lo, hi = s_hi; // current range (of keys, or of trial indices when on_idx)
        uint64_t blo = lo, bhi = hi;   // boundary bin
        for (int pass = 0; pass < 16; ++pass) {
            const uint64_t width = hi - lo;
            const int shift = width < (uint64_t)FS_BINS ? 0 : (64 - __clzll((long long)width)) - FS_BITS;
            for (int b = tid; b < FS_BINS; b += THREADS) hist[b] = 0;
            __syncthreads();
            for (int i = tid; i < Ms; i += THREADS) {
                const uint64_t k = skey[i];
                if (on_idx) { if (k == kstar) { const uint64_t v = sidx[i]; if (v >= lo && v <= hi) atomicAdd(&hist[(int)((v - lo) >> shift)], 1); } }
                else if (k >= lo && k <= hi) atomicAdd(&hist[(int)((k - lo) >> shift)], 1);
            }
            __syncthreads();
            {
                constexpr int PER = FS_BINS / THREADS;
                const int hb = FS_BINS - 1 - tid * PER; // thread 0 owns the highest bins
                int sum = 0;
#pragma unroll
                for (int j = 0; j < PER; ++j) sum += hist[hb - j];
                int incl = sum;
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(kFullMask, incl, d);
                    if (lane >= d) incl += v;
                }
                if (lane == 31) s_wsum[warp] = incl;
                __syncthreads();
                int woff = 0;
                for (int w = 0; w < warp; ++w) woff += s_wsum[w];
                incl += woff;
                const int excl = incl - sum;
                if ((int64_t)excl < need && need <= (int64_t)incl) {
                    int above = excl;
                    for (int j = 0; j < PER; ++j) {
                        const int h = hist[hb - j];
                        if (need <= (int64_t)above + h) { s_sel = hb - j; s_above = above; s_bucket = h; break; }
                        above += h;
                    }
                }
            }
            __syncthreads();
            need -= s_above;
            bucket = s_bucket;
            blo = lo + ((uint64_t)s_sel << shift);
            bhi = shift == 0 ? blo : blo + ((1ull << shift) - 1);
            if (bhi > hi) bhi = hi;
            __syncthreads();
            if (bucket <= FS_SORT) break;
            if (shift == 0) {
                // a single value holds more than FS_SORT candidates
                if (on_idx) break; // cannot happen: trial indices are distinct
                on_idx = true;
                kstar = blo;
                lo = 0; hi = 0xffffffffull;
            } else { lo = blo; hi = bhi; }
        }
        // --- gather: above the boundary bin -> kept; inside it -> the small bucket.  Warp-aggregated slot claims
        // (one shared-memory atomic per warp and round, not one per element)
        for (int base = 0; base < Ms; base += THREADS) {
            const int i = base + tid;
            int rel = -1; // +1 above, 0 inside, -1 below / out of range
            uint64_t k = 0;
            uint32_t id = 0;
            if (i < Ms) {
                k = skey[i];
                id = sidx[i];
                if (!on_idx) rel = k > bhi ? 1 : (k >= blo ? 0 : -1);
                else if (k != kstar) rel = k > kstar ? 1 : -1;
                else rel = (uint64_t)id > bhi ? 1 : ((uint64_t)id >= blo ? 0 : -1);
            }
            const unsigned mk = __ballot_sync(kFullMask, rel > 0), mc = __ballot_sync(kFullMask, rel == 0);
            int bk = 0, bc = 0;
            if (lane == 0) {
                if (mk) bk = atomicAdd(&s_nkept, __popc(mk));
                if (mc) bc = atomicAdd(&s_ncand, __popc(mc));
            }
            bk = __shfl_sync(kFullMask, bk, 0);
            bc = __shfl_sync(kFullMask, bc, 0);
            const unsigned lt = (1u << lane) - 1u;
            if (rel > 0) {
                const int o = bk + __popc(mk & lt);
                kk[o] = k;
                ki[o] = (int32_t)id;
            } else if (rel == 0) {
                const int o = bc + __popc(mc & lt);
                ckey[o] = k;
                cidx[o] = id;
            }
        }
        __syncthreads();
        // --- rank the bucket by (key, trial) descending: rank = number of larger elements (no barriers in the loop)
        const int base = s_nkept; // == t - need
        if (tid < bucket) {
            const u128 me = comp_of(ckey[tid], cidx[tid]);
            int rank = 0;
            for (int j = 0; j < bucket; ++j) rank += comp_of(ckey[j], cidx[j]) > me;
            if (rank < (int)need) {
                kk[base + rank] = ckey[tid];
                ki[base + rank] = (int32_t)cidx[tid];
            }
            if (rank == (int)need - 1) { // the need-th largest candidate is the VaR order statistic (so far)
                const double vl = fkey_inv(ckey[tid]);
                a.var_loss[row] = vl;
                a.var_trial[row] = (int32_t)cidx[tid];
                if (a.tkey) a.tkey[row] = fkey(0.0 - vl);
            }
        }
        return;
    }
    // ===== general path: more candidates than shared memory holds; digit passes straight from global memory
    hist = reinterpret_cast<int *>(area); // (nothing is staged: the wider histogram takes the staging area's place)
    auto for_each = [&](auto f) { for_each_global<THREADS>(a, row, segoff, f); };

    // Radix select on the 64-bit loss key first (cheap 64-bit digit extraction); the trial index only breaks ties,
    // so its digits are visited only if more than SEL_SORT candidates share one exact loss value.
    // `kpre`/`ipre` = selected high bits of key / trial; `kshift`/`ishift` = bits not yet resolved.
    uint64_t kpre = 0;
    uint32_t ipre = 0;
    int kshift = 64, ishift = 32;
    int64_t need = t;    // rank from the top still to resolve inside the prefix bucket
    int bucket = 0;
    auto in_prefix = [&](uint64_t key, uint32_t id) -> bool {
        if (kshift < 64 && (key >> kshift) != kpre) return false;
        if (kshift == 0 && ishift < 32 && (id >> ishift) != ipre) return false;
        return true;
    };
    for (int pass = 0; pass < 12; ++pass) {
        const bool on_key = kshift > 0;
        const int rem = on_key ? kshift : ishift;
        const int width = rem >= SEL_BITS ? SEL_BITS : rem;
        const int sh = rem - width;
        const uint32_t dmask = (1u << width) - 1u;
        for (int b = tid; b < SEL_BINS; b += THREADS) hist[b] = 0;
        __syncthreads();
        for_each([&](uint64_t key, uint32_t id) {
            if (in_prefix(key, id)) atomicAdd(&hist[on_key ? (uint32_t)(key >> sh) & dmask : (id >> sh) & dmask], 1);
        });
        __syncthreads();
        // find the digit where the count from the top reaches `need`: thread 0 owns the highest bins
        {
            constexpr int PER = SEL_BINS / THREADS;
            const int hi = SEL_BINS - 1 - tid * PER;
            int sum = 0;
#pragma unroll
            for (int j = 0; j < PER; ++j) sum += hist[hi - j];
            int incl = sum;
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(kFullMask, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < warp; ++w) woff += s_wsum[w];
            incl += woff;
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
        if (on_key) { kpre = (width == 64 ? 0 : (kpre << width)) | (uint64_t)(uint32_t)s_sel; kshift = sh; }
        else { ipre = (ipre << width) | (uint32_t)s_sel; ishift = sh; }
        need -= s_above;
        bucket = s_bucket;
        __syncthreads();
        if (bucket <= SEL_SORT) break;
    }
    // ---- gather: strictly-above-prefix -> kept; equal-prefix -> the small boundary bucket
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    for_each([&](uint64_t key, uint32_t id) {
        // compare the resolved prefix of the composite (key, trial)
        int rel; // +1 above, 0 inside, -1 below
        const uint64_t kp = kshift < 64 ? (key >> kshift) : 0;
        if (kp != kpre) rel = kp > kpre ? 1 : -1;
        else if (kshift == 0 && ishift < 32) {
            const uint32_t ip = id >> ishift;
            rel = ip == ipre ? 0 : (ip > ipre ? 1 : -1);
        } else rel = 0;
        if (rel > 0) {
            const int o = atomicAdd(&s_nkept, 1);
            kk[o] = key;
            ki[o] = (int32_t)id;
        } else if (rel == 0) {
            const int o = atomicAdd(&s_ncand, 1);
            ckey[o] = key;
            cidx[o] = id;
        }
    });
    __syncthreads();
    // ---- bitonic sort of the boundary bucket, descending composite; pads (0,0) sink to the end
    int m = 1;
    while (m < bucket) m <<= 1;
    for (int i = bucket + tid; i < m; i += THREADS) { ckey[i] = 0; cidx[i] = 0; }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += THREADS) {
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
    for (int i = tid; i < (int)need; i += THREADS) {
        kk[base + i] = ckey[i];
        ki[base + i] = (int32_t)cidx[i];
    }
    if (tid == 0) { // the need-th largest candidate is the VaR order statistic (so far)
        const double vl = fkey_inv(ckey[need - 1]);
        a.var_loss[row] = vl;
        a.var_trial[row] = (int32_t)cidx[need - 1];
        if (a.tkey) a.tkey[row] = fkey(0.0 - vl); // canonical +0.0 for a zero loss, so V = -0.0 and +0.0 both pass
    }
}

// ============================================================ selection, dense step, register-resident
// The dense first step of the streaming path hands every row n0 = max(2048, 4t) contiguous values and wants only the
// top `tail` of them (the whole tail t on the guaranteed schedule, the ~t n0/N + 7 sigma best under the optimistic
// bound).  When n0 <= 256 * ITEMS the row fits the CTA's REGISTERS: every thread loads ITEMS values with independent
// coalesced loads (all in flight at once), the trial index is implicit in the slot, and the range-adaptive radix
// passes, the gather and the boundary-bucket sort run without staging anything in shared memory -- 5 KB of shared
// memory per CTA instead of ~60 KB, so the SM holds as many rows as its registers allow.
// from the histogram (bins descending from FS_BINS-1): the bin where the count from the top reaches `need`
template <int THREADS, int BINS = FS_BINS>
__device__ __forceinline__ void find_boundary_bin(const int *hist, int64_t need, int *s_wsum, int *s_sel, int *s_above, int *s_bucket)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int PER = BINS / THREADS;
    const int hb = BINS - 1 - tid * PER; // thread 0 owns the highest bins
    int sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) sum += hist[hb - j];
    int incl = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_wsum[w];
    incl += woff;
    const int excl = incl - sum;
    if ((int64_t)excl < need && need <= (int64_t)incl) {
        int above = excl;
        for (int j = 0; j < PER; ++j) {
            const int h = hist[hb - j];
            if (need <= (int64_t)above + h) { *s_sel = hb - j; *s_above = above; *s_bucket = h; break; }
            above += h;
        }
    }
    __syncthreads();
}

template <int ITEMS, int THREADS>
__device__ void select_dense_row(const SelArgs &a, int64_t row);

template <int ITEMS, int THREADS>
__global__ void __launch_bounds__(THREADS, (ITEMS > 8 ? 3 : 4) * (256 / THREADS)) k_select_dense(const SelArgs a)
{
    select_dense_row<ITEMS, THREADS>(a, blockIdx.x);
}

template <int ITEMS, int THREADS>
__device__ void select_dense_row(const SelArgs &a, const int64_t row)
{
    // bins per pass: 1 024 for the 2 048 - 4 096-value samples, 256 for samples of up to 1 024 values (C3: fewer bins to zero
    // and scan per row; the boundary bin still holds a handful of elements)
    constexpr int DB_BITS = ITEMS * THREADS <= 1024 ? 8 : FS_BITS, DB_BINS = 1 << DB_BITS;
    __shared__ int hist[DB_BINS];
    __shared__ uint64_t ckey[FS_SORT];
    __shared__ uint32_t cidx[FS_SORT];
    __shared__ double s_red[2 * (THREADS / 32)];
    __shared__ int s_wsum[THREADS / 32];
    __shared__ int s_sel, s_above, s_bucket;
    const int64_t t = a.tail;
    const int n = (int)a.step_len;
    uint64_t *kk = a.kept_key_out + row * t;
    int32_t *ki = a.kept_idx_out + row * t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *V = reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(a.cand) + row * a.pitch);
    for (int b = tid; b < DB_BINS; b += THREADS) hist[b] = 0; // (visible after the barrier of the range reduction)
    // ---- load: slot i = tid + 256 j holds the loss of trial n_begin + i (slots past n repeat the thread's first
    // loss: harmless for min / max, masked everywhere else)
    double x[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int i = tid + THREADS * j;
        x[j] = 0.0 - __ldg(V + (i < n ? i : (tid < n ? tid : 0)));
    }
    // range of the row from the HIGH WORDS of the order-preserving keys (one integer min / max per value instead of
    // FP64 compares): [lo, hi] is a slightly generous cover of the losses, which is all the binning needs.  A
    // non-finite loss shows up as a key beyond +-infinity and sends the row to the key-space passes.
    uint32_t hmn = 0xffffffffu, hmx = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t hw = (uint32_t)__double2hiint(x[j]);
        const uint32_t hk = (hw >> 31) ? ~hw : (hw | 0x80000000u);
        hmn = min(hmn, hk); hmx = max(hmx, hk);
    }
    hmn = __reduce_min_sync(kFullMask, hmn);
    hmx = __reduce_max_sync(kFullMask, hmx);
    uint32_t *s_redw = reinterpret_cast<uint32_t *>(s_red);
    if (lane == 0) { s_redw[warp] = hmn; s_redw[THREADS / 32 + warp] = hmx; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) { hmn = min(hmn, s_redw[w]); hmx = max(hmx, s_redw[THREADS / 32 + w]); }
    const bool finite = hmx < 0xfff00000u && hmn > 0x000fffffu; // inside (-inf, +inf)
    const double mn = fkey_inv((uint64_t)hmn << 32), mx = fkey_inv(((uint64_t)hmx << 32) | 0xffffffffull);
    // ---- pass 0 in VALUE space: DB_BINS equal-width bins of [min, max].  The bin index is a monotone function of
    // the loss (the same three roundings for every element), so "higher bin" implies "not smaller": the partition is
    // exact.  Losses of a Monte Carlo run are smooth, so one pass leaves a boundary bin of a few elements; ties and
    // pathological distributions continue below with the key-space passes, non-finite values skip straight to them.
    int64_t need = t;
    int bucket = n;
    const double width = mx - mn;
    const bool byval = finite && width > 1.0e-290 && width < 1.0e300; // scale = DB_BINS / width must stay finite (0 * inf = NaN)
    const double scale = byval ? (double)DB_BINS / width : 0.0;
    auto bin_of = [&](double v) -> int {
        const int b = __double2int_rd((v - mn) * scale);
        return min(max(b, 0), DB_BINS - 1);
    };
    int sel = 0;
    if (byval) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
            if (tid + THREADS * j < n) atomicAdd(&hist[bin_of(x[j])], 1);
        __syncthreads();
        find_boundary_bin<THREADS, DB_BINS>(hist, need, s_wsum, &s_sel, &s_above, &s_bucket);
        sel = s_sel;
        need -= s_above;
        bucket = s_bucket;
    }
    // ---- key-space refinement of the boundary bucket (rare): range-adaptive radix passes on the 64-bit keys
    bool on_idx = false, bykey = false;
    uint64_t kstar = 0, blo = 0, bhi = ~0ull;
    if (bucket > FS_SORT) {
        bykey = true;
        // key range of the bucket (monotonicity: every element whose key lies inside it is in the bucket)
        uint64_t kmn = ~0ull, kmx = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (tid + THREADS * j < n && (!byval || bin_of(x[j]) == sel)) {
                const uint64_t k = fkey(x[j]);
                kmn = k < kmn ? k : kmn; kmx = k > kmx ? k : kmx;
            }
        }
        for (int d = 16; d; d >>= 1) {
            const uint64_t a1 = __shfl_xor_sync(kFullMask, kmn, d), b1 = __shfl_xor_sync(kFullMask, kmx, d);
            kmn = a1 < kmn ? a1 : kmn; kmx = b1 > kmx ? b1 : kmx;
        }
        uint64_t *s_redu = reinterpret_cast<uint64_t *>(s_red);
        __syncthreads();
        if (lane == 0) { s_redu[warp] = kmn; s_redu[THREADS / 32 + warp] = kmx; }
        __syncthreads();
        uint64_t lo = ~0ull, hi = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) { lo = s_redu[w] < lo ? s_redu[w] : lo; hi = s_redu[THREADS / 32 + w] > hi ? s_redu[THREADS / 32 + w] : hi; }
        blo = lo; bhi = hi;
        for (int pass = 0; pass < 16; ++pass) {
            const uint64_t w64 = hi - lo;
            const int shift = w64 < (uint64_t)DB_BINS ? 0 : (64 - __clzll((long long)w64)) - DB_BITS;
            for (int b = tid; b < DB_BINS; b += THREADS) hist[b] = 0;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const int i = tid + THREADS * j;
                if (i < n) {
                    const uint64_t k = fkey(x[j]);
                    if (on_idx) { if (k == kstar) { const uint64_t v = (uint32_t)(a.n_begin + i); if (v >= lo && v <= hi) atomicAdd(&hist[(int)((v - lo) >> shift)], 1); } }
                    else if (k >= lo && k <= hi) atomicAdd(&hist[(int)((k - lo) >> shift)], 1);
                }
            }
            __syncthreads();
            find_boundary_bin<THREADS, DB_BINS>(hist, need, s_wsum, &s_sel, &s_above, &s_bucket);
            need -= s_above;
            bucket = s_bucket;
            blo = lo + ((uint64_t)s_sel << shift);
            bhi = shift == 0 ? blo : blo + ((1ull << shift) - 1);
            if (bhi > hi) bhi = hi;
            __syncthreads();
            if (bucket <= FS_SORT) break;
            if (shift == 0) {
                if (on_idx) break; // cannot happen: trial indices are distinct
                on_idx = true;
                kstar = blo;
                lo = 0; hi = 0xffffffffull;
            } else { lo = blo; hi = bhi; }
        }
    }
    // ---- gather without atomics: per-thread counts (kept | bucket << 16), one block scan, deterministic slots
    int cnt = 0;
    uint32_t relk = 0, relb = 0; // bit j: my item j is above / inside the boundary bin
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int i = tid + THREADS * j;
        if (i < n) {
            int rel; // +1 above, 0 inside, -1 below
            if (!bykey) { const int b = bin_of(x[j]); rel = b > sel ? 1 : (b == sel ? 0 : -1); }
            else {
                const uint64_t k = fkey(x[j]);
                const uint64_t id = (uint32_t)(a.n_begin + i);
                if (!on_idx) rel = k > bhi ? 1 : (k >= blo ? 0 : -1);
                else if (k != kstar) rel = k > kstar ? 1 : -1;
                else rel = id > bhi ? 1 : (id >= blo ? 0 : -1);
            }
            if (rel > 0) { relk |= 1u << j; cnt += 1; }
            else if (rel == 0) { relb |= 1u << j; cnt += 1 << 16; }
        }
    }
    int incl = cnt;
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += v;
    }
    __syncthreads();
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int off = incl - cnt;
    for (int w = 0; w < warp; ++w) off += s_wsum[w];
    int ok = off & 0xffff, ob = off >> 16;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int32_t id = (int32_t)(a.n_begin + tid + THREADS * j);
        if (relk & (1u << j)) { kk[ok] = fkey(x[j]); ki[ok] = id; ++ok; }
        else if (relb & (1u << j)) { ckey[ob] = fkey(x[j]); cidx[ob] = (uint32_t)id; ++ob; }
    }
    __syncthreads();
    // ---- rank the bucket by (key, trial) descending
    const int base = (int)(t - need); // elements strictly above the boundary bin
    if (tid < bucket) {
        const u128 me = comp_of(ckey[tid], cidx[tid]);
        int rank = 0;
        for (int j = 0; j < bucket; ++j) rank += comp_of(ckey[j], cidx[j]) > me;
        if (rank < (int)need) {
            kk[base + rank] = ckey[tid];
            ki[base + rank] = (int32_t)cidx[tid];
        }
        if (rank == (int)need - 1) {
            const double vl = fkey_inv(ckey[tid]);
            a.var_loss[row] = vl;
            a.var_trial[row] = (int32_t)cidx[tid];
            if (a.tkey) a.tkey[row] = fkey(0.0 - vl);
        }
    }
}

// ============================================================ selection, one WARP per row (sparse steps of small rows)
// The CTA-per-row kernel spends a small row's time on its fixed work: C3 (10 000 trials, tail 101) hands k_select_t<64> ~340
// candidates per row, and ncu counts 280 thread instructions per CANDIDATE for it (a 256-bin histogram zeroed and scanned, a
// dozen CTA barriers, per-segment books by 64 threads): 0.73 ms for 125 000 rows.  Here a warp owns a row: 128 bins (four per
// lane, found by one shuffle scan), a 32-element boundary bucket ranked by one lane per element, warp barriers only, four
// independent rows per CTA: 0.48 ms.  Same results: the top `tail` of (loss key, trial) exactly, in any order, and the tail-th
// largest as the VaR order statistic.  A row that does not fit (more staged candidates than WS_CAP: hostile trial orders) is
// appended to a list and redone by the CTA-per-row kernel right after (k_select_t_list): exactness never depends on the row
// being small.  (The same for the dense first step -- 1 024 values per row in a warp's registers -- was measured at 0.57 ms
// against 0.55 for k_select_dense<8,128> and dropped: per-slot ballots and 167 registers.)
constexpr int WS_WARPS = 4;     // rows per CTA
constexpr int WS_CAP = 512;     // staged candidates per row
constexpr int WS_BITS = 7, WS_BINS = 1 << WS_BITS; // bins per range-adaptive pass: four per lane
constexpr int WS_SORT = 32;     // rank-sorted boundary bucket: one lane per element
constexpr int WS_MAXSEG = 64;   // segments of a sparse step the warp keeps books for (two per lane)

struct WsRow { // one warp's shared memory
    uint64_t key[WS_CAP];
    uint32_t idx[WS_CAP];
    int hist[WS_BINS];
    uint64_t ckey[WS_SORT];
    uint32_t cidx[WS_SORT];
    int segoff[WS_MAXSEG + 2];
};

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) { const uint64_t o = __shfl_xor_sync(kFullMask, v, d); v = o < v ? o : v; }
    return v;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) { const uint64_t o = __shfl_xor_sync(kFullMask, v, d); v = o > v ? o : v; }
    return v;
}

// From the warp's histogram (bins descending from WS_BINS-1, lane 0 owns the four highest): the bin where the count from the
// top reaches `need`, the count above it and its own population.
__device__ __forceinline__ void ws_boundary_bin(const int *hist, int need, int lane, int &sel, int &above, int &bucket)
{
    const int hb = WS_BINS - 1 - 4 * lane;
    const int h0 = hist[hb], h1 = hist[hb - 1], h2 = hist[hb - 2], h3 = hist[hb - 3];
    const int sum = h0 + h1 + h2 + h3;
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += v;
    }
    const int excl = incl - sum;
    const bool mine = excl < need && need <= incl;
    int s = 0, ab = 0, bu = 0;
    if (mine) {
        ab = excl;
        if (need <= ab + h0) { s = hb; bu = h0; }
        else if (need <= ab + h0 + h1) { s = hb - 1; ab += h0; bu = h1; }
        else if (need <= ab + h0 + h1 + h2) { s = hb - 2; ab += h0 + h1; bu = h2; }
        else { s = hb - 3; ab += h0 + h1 + h2; bu = h3; }
    }
    const unsigned who = __ballot_sync(kFullMask, mine);
    const int src = who ? __ffs(who) - 1 : 0;
    sel = __shfl_sync(kFullMask, s, src);
    above = __shfl_sync(kFullMask, ab, src);
    bucket = __shfl_sync(kFullMask, bu, src);
}

// The top `need` of the M staged candidates of W (need <= M <= WS_CAP) by (key, trial), written to kk / ki slots
// [out0, out0 + need) in any order; the need-th largest is the VaR order statistic (so far).
__device__ __forceinline__ void ws_select(WsRow &W, int M, int need, int out0, uint64_t mn, uint64_t mx, uint64_t *kk, int32_t *ki, const SelArgs &a,
                                          int64_t row, int lane)
{
    uint64_t lo = warp_min_u64(mn), hi = warp_max_u64(mx); // (mn / mx: this lane's share of the staged keys' range)
    bool on_idx = false;      // ties: more than WS_SORT candidates share one loss value -> continue on the trial index
    uint64_t kstar = 0, blo = lo, bhi = hi;
    int bucket = M, rem = need; // rem: rank from the top still to resolve inside the boundary bucket
    for (int pass = 0; pass < 16; ++pass) {
        const uint64_t width = hi - lo;
        const int shift = width < (uint64_t)WS_BINS ? 0 : (64 - __clzll((long long)width)) - WS_BITS;
#pragma unroll
        for (int b = 0; b < 4; ++b) W.hist[lane + 32 * b] = 0;
        __syncwarp();
        for (int i = lane; i < M; i += 32) {
            const uint64_t k = W.key[i];
            if (on_idx) { if (k == kstar) { const uint64_t v = W.idx[i]; if (v >= lo && v <= hi) atomicAdd(&W.hist[(int)((v - lo) >> shift)], 1); } }
            else if (k >= lo && k <= hi) atomicAdd(&W.hist[(int)((k - lo) >> shift)], 1);
        }
        __syncwarp();
        int sel, above;
        ws_boundary_bin(W.hist, rem, lane, sel, above, bucket);
        rem -= above;
        blo = lo + ((uint64_t)sel << shift);
        bhi = shift == 0 ? blo : blo + ((1ull << shift) - 1);
        if (bhi > hi) bhi = hi;
        __syncwarp();
        if (bucket <= WS_SORT) break;
        if (shift == 0) {
            if (on_idx) break; // cannot happen: trial indices are distinct
            on_idx = true;
            kstar = blo;
            lo = 0; hi = 0xffffffffull;
        } else { lo = blo; hi = bhi; }
    }
    // --- gather: strictly above the boundary bin (need - rem of them) -> kept; inside it -> the bucket (ballot-ranked slots)
    int nk = 0, nb = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int b0 = 0; b0 < M; b0 += 32) {
        const int i = b0 + lane;
        int rel = -1; // +1 above, 0 inside, -1 below / out of range
        uint64_t k = 0;
        uint32_t id = 0;
        if (i < M) {
            k = W.key[i]; id = W.idx[i];
            if (!on_idx) rel = k > bhi ? 1 : (k >= blo ? 0 : -1);
            else if (k != kstar) rel = k > kstar ? 1 : -1;
            else rel = (uint64_t)id > bhi ? 1 : ((uint64_t)id >= blo ? 0 : -1);
        }
        const unsigned mk = __ballot_sync(kFullMask, rel > 0), mc = __ballot_sync(kFullMask, rel == 0);
        if (rel > 0) { const int o = out0 + nk + __popc(mk & lt); kk[o] = k; ki[o] = (int32_t)id; }
        else if (rel == 0) { const int o = nb + __popc(mc & lt); if (o < WS_SORT) { W.ckey[o] = k; W.cidx[o] = id; } }
        nk += __popc(mk); nb += __popc(mc);
    }
    __syncwarp();
    // --- rank the bucket by (key, trial) descending: rank = number of larger elements
    const int base = out0 + need - rem; // == out0 + nk
    if (lane < bucket && lane < WS_SORT) {
        const uint64_t mk = W.ckey[lane];
        const uint32_t mi = W.cidx[lane];
        int rank = 0;
        for (int j = 0; j < bucket; ++j) {
            const uint64_t kj = W.ckey[j];
            rank += (kj > mk || (kj == mk && W.cidx[j] > mi)) ? 1 : 0;
        }
        if (rank < rem) { kk[base + rank] = W.ckey[lane]; ki[base + rank] = (int32_t)W.cidx[lane]; }
        if (rank == rem - 1) { // the need-th largest candidate is the VaR order statistic (so far)
            const double vl = fkey_inv(W.ckey[lane]);
            a.var_loss[row] = vl;
            a.var_trial[row] = (int32_t)W.cidx[lane];
            if (a.tkey) a.tkey[row] = fkey(0.0 - vl); // canonical +0.0 for a zero loss, so V = -0.0 and +0.0 both pass
        }
    }
    __syncwarp();
}

// sparse step (and any step whose candidates are the kept tail + lane-private segments of records): M <= WS_CAP staged
__global__ void __launch_bounds__(WS_WARPS * 32) k_wsel_sparse(const SelArgs a, int64_t rows, int *ovf_cnt, int32_t *ovf_rows)
{
    __shared__ WsRow s_rows[WS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * WS_WARPS + warp;
    if (row >= rows) return;
    WsRow &W = s_rows[warp];
    const int t = (int)a.tail;
    uint64_t *kk = a.kept_key_out + row * a.tail;
    int32_t *ki = a.kept_idx_out + row * a.tail;
    const int kn = (int)a.kept_n;
    // everything the row needs first -- the kept tail (two entries per lane), the segment counts -- is requested before
    // anything is waited for: one global round trip instead of three
    const uint64_t *kin = a.kept_key_in + row * a.kept_n;
    const int32_t *iin = a.kept_idx_in + row * a.kept_n;
    uint64_t pk0 = 0, pk1 = 0;
    int32_t pi0 = 0, pi1 = 0;
    if (lane < kn) { pk0 = kin[lane]; pi0 = iin[lane]; }
    if (lane + 32 < kn) { pk1 = kin[lane + 32]; pi1 = iin[lane + 32]; }
    int c0 = lane < a.nsub ? a.cnt[row * a.nsub + lane] : 0;
    int c1 = lane + 32 < a.nsub ? a.cnt[row * a.nsub + lane + 32] : 0;
    // a row that already lost its optimistic bet at an earlier step stays lost (it is redone separately)
    if (kn > 0 && __shfl_sync(kFullMask, pi0, 0) < 0) {
        if (lane == 0) ki[0] = -1;
        return;
    }
    // ---- segment books: exclusive prefix of the counts (two segments per lane)
    int i0 = c0, i1 = c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u0 = __shfl_up_sync(kFullMask, i0, d), u1 = __shfl_up_sync(kFullMask, i1, d);
        if (lane >= d) { i0 += u0; i1 += u1; }
    }
    const int tot0 = __shfl_sync(kFullMask, i0, 31), total = tot0 + __shfl_sync(kFullMask, i1, 31);
    const int M = kn + total;
    if (M < t) {
        // only possible under an optimistic filter bound that turned out too tight for this row: report the row (it is
        // redone with the guaranteed schedule) and mark its kept list so that the finalize kernel skips it
        if (lane == 0) {
            ki[0] = -1;
            if (a.fail_cnt) a.fail_rows[atomicAdd(a.fail_cnt, 1)] = (int32_t)(a.row_base + row);
        }
        return;
    }
    if (M > WS_CAP) { // does not fit the warp's staging area: the CTA-per-row kernel redoes the row
        if (lane == 0) ovf_rows[atomicAdd(ovf_cnt, 1)] = (int32_t)row;
        return;
    }
    W.segoff[lane] = i0 - c0;
    W.segoff[lane + 32] = tot0 + i1 - c1;
    if (lane == 0) W.segoff[WS_MAXSEG] = total;
    // ---- stage: the kept tail, then the records (every lane labels the slots of its segments, a flat loop loads them)
    uint64_t mn = ~0ull, mx = 0;
    if (lane < kn) { W.key[lane] = pk0; W.idx[lane] = (uint32_t)pi0; mn = pk0; mx = pk0; }
    if (lane + 32 < kn) { W.key[lane + 32] = pk1; W.idx[lane + 32] = (uint32_t)pi1; mn = pk1 < mn ? pk1 : mn; mx = pk1 > mx ? pk1 : mx; }
    for (int i = lane + 64; i < kn; i += 32) {
        const uint64_t k = kin[i];
        W.key[i] = k; W.idx[i] = (uint32_t)iin[i];
        mn = k < mn ? k : mn; mx = k > mx ? k : mx;
    }
    for (int j = 0; j < c0; ++j) W.idx[kn + i0 - c0 + j] = (uint32_t)lane;
    for (int j = 0; j < c1; ++j) W.idx[kn + tot0 + i1 - c1 + j] = (uint32_t)(lane + 32);
    __syncwarp();
    const int4 *R = reinterpret_cast<const int4 *>(reinterpret_cast<const unsigned char *>(a.cand) + row * a.pitch);
    for (int i0 = lane; i0 < total; i0 += 128) { // four records per lane and round: four loads in flight (see k_select_t)
        int4 rec[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            if (i < total) {
                const int g = (int)W.idx[kn + i];
                rec[u] = __ldg(R + (int64_t)g * a.sub_len + (i - W.segoff[g]));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            if (i < total) {
                const uint64_t k = fkey(0.0 - __hiloint2double(rec[u].y, rec[u].x));
                W.key[kn + i] = k;
                W.idx[kn + i] = (uint32_t)rec[u].z;
                mn = k < mn ? k : mn; mx = k > mx ? k : mx;
            }
        }
    }
    __syncwarp();
    ws_select(W, M, t, 0, mn, mx, kk, ki, a, row, lane);
}

static size_t sel_smem_bytes(int cap, int nseg, int threads = SEL_THREADS)
{
    const size_t seg = ((size_t)(nseg + 2) * 4 + 15) & ~(size_t)15;
    const size_t staged = (size_t)cap * 12 + (size_t)(threads >= 256 ? 1024 : 256) * 4;  // keys, trials, the staged path's bins
    const size_t general = (size_t)(threads >= 256 ? SEL_BINS_MAX : 256) * 4;            // the general path's bins
    return seg + (staged > general ? staged : general);
}

// opt-in to the large dynamic shared memory of both instantiations
static cudaError_t sel_configure(int bytes)
{
    cudaError_t e = cudaFuncSetAttribute(k_select_t<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_select_t<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_select_t_list<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_select_t_list<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_select_t_list<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_select_t<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// ============================================================ finalize
// Kernel (3): the kept tail -> ascending breach-trial list, and CVaR = mean tail loss.
// Stream compaction over a per-portfolio trial bitmap in shared memory: set one bit per kept
// trial, prefix-sum the word popcounts, then every set bit knows its rank.  (Falls back to a
// bitonic sort by trial index when the trial range does not fit the bitmap.)
constexpr int FIN_THREADS = 256;
// shared memory of k_finalize_bitmap: bitmap + per-word prefix (N/4 bytes) + the tail's losses in list order (8 t bytes)
static size_t fin_bitmap_smem(int64_t n_trials, int64_t t) { return (size_t)((n_trials + 31) / 32) * 8 + (size_t)t * 8; }
constexpr size_t FIN_BITMAP_MAX_SMEM = 200 * 1024;

template <int FIN_THREADS>
__global__ void __launch_bounds__(FIN_THREADS) k_finalize_bitmap_t(const uint64_t *kept_key, const int32_t *kept_idx, int64_t t,
                                                                 int64_t n_trials, int32_t trial_base, int32_t *breach,
                                                                 double *cvar)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ int s_wsum[FIN_THREADS / 32];
    const int nwords = (int)((n_trials + 31) >> 5);
    double *sl = reinterpret_cast<double *>(smem_raw);               // [t] losses in ascending-trial order
    uint32_t *bm = reinterpret_cast<uint32_t *>(sl + t);             // [nwords] one bit per kept trial
    int *pre = reinterpret_cast<int *>(bm + nwords);                 // [nwords] kept trials before this word
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the kept tail is read ONCE, up front (up to FIN_PER entries per thread in registers; all loads in flight while the
    // bitmap is zeroed): read where it is used, the kernel paid three dependent global round trips per row
    constexpr int FIN_PER = 4;
    const bool inreg = t <= (int64_t)FIN_PER * FIN_THREADS;
    int32_t rid[FIN_PER];
    uint64_t rkey[FIN_PER];
    if (inreg) {
#pragma unroll
        for (int u = 0; u < FIN_PER; ++u) {
            const int i = tid + u * FIN_THREADS;
            rid[u] = i < t ? __ldg(kept_idx + row * t + i) : 0;
            rkey[u] = i < t ? __ldg(kept_key + row * t + i) : 0ull;
        }
    }
    for (int i = tid; i < nwords; i += FIN_THREADS) bm[i] = 0;
    const int failed = tid == 0 && (inreg ? rid[0] : kept_idx[row * t]) < 0;
    if (__syncthreads_or(failed)) return; // row failed its optimistic bound: redone separately (k_select)
    if (inreg) {
#pragma unroll
        for (int u = 0; u < FIN_PER; ++u)
            if (tid + u * FIN_THREADS < t) { const int32_t id = rid[u] - trial_base; atomicOr(&bm[id >> 5], 1u << (id & 31)); }
    } else {
        for (int i = tid; i < t; i += FIN_THREADS) {
            const int32_t id = kept_idx[row * t + i] - trial_base;
            atomicOr(&bm[id >> 5], 1u << (id & 31));
        }
    }
    __syncthreads();
    // exclusive prefix of the word popcounts: each thread owns a contiguous run of words
    const int per = (nwords + FIN_THREADS - 1) / FIN_THREADS;
    const int w0 = min(nwords, tid * per), w1 = min(nwords, w0 + per);
    int c = 0;
    for (int w = w0; w < w1; ++w) c += __popc(bm[w]);
    int incl = c;
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int off = incl - c;
    for (int w = 0; w < warp; ++w) off += s_wsum[w];
    for (int w = w0; w < w1; ++w) { pre[w] = off; off += __popc(bm[w]); }
    __syncthreads();
    // every kept trial knows its rank in the ascending list: emit the trial and park its loss at that rank
    int32_t *dst = breach + row * t;
    if (inreg) {
#pragma unroll
        for (int u = 0; u < FIN_PER; ++u)
            if (tid + u * FIN_THREADS < t) {
                const int32_t id = rid[u] - trial_base;
                const int rank = pre[id >> 5] + __popc(bm[id >> 5] & ((1u << (id & 31)) - 1u));
                dst[rank] = rid[u];
                sl[rank] = fkey_inv(rkey[u]);
            }
    } else {
        for (int i = tid; i < t; i += FIN_THREADS) {
            const int32_t gid = kept_idx[row * t + i];
            const int32_t id = gid - trial_base;
            const int rank = pre[id >> 5] + __popc(bm[id >> 5] & ((1u << (id & 31)) - 1u));
            dst[rank] = gid;
            sl[rank] = fkey_inv(kept_key[row * t + i]);
        }
    }
    __syncthreads();
    // CVaR = mean tail loss, summed in list order with a fixed tree: bit-reproducible whatever order the
    // selection kernel's atomics left the kept tail in
    double s = 0.0;
    for (int i = tid; i < t; i += FIN_THREADS) s += sl[i];
    s = block_sum(s, red);
    if (tid == 0) cvar[row] = s / (double)t;
}

template <int FIN_THREADS>
__global__ void __launch_bounds__(FIN_THREADS) k_finalize_t(const uint64_t *kept_key, const int32_t *kept_idx, int64_t t,
                                                          int m /* pow2 >= t */, int32_t *breach, double *cvar,
                                                          uint64_t *gkey /* global scratch or null */, int32_t *gidx)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    const int64_t row = blockIdx.x;
    if (kept_idx[row * t] < 0) return; // row failed its optimistic bound: redone separately (k_select)
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

// Finalize for LARGE tails over a trial range too long for the bitmap (C5: 10 001 of 10 M trials per row): the bitonic sort above
// is 105 barrier-separated stages over 16 384 padded entries (350 us per row: 29 of C5's 299 ms).  Trial indices are distinct, so
// an ascending list is two counting steps: (1) bucket = trial >> shift (<= 4 096 buckets of <= 4 096 trials): histogram,
// exclusive scan, scatter -- the buckets are then in place, their contents in any order; (2) a warp per bucket ranks its
// entries: up to 32 by comparing one another through shuffles (the common case: ~2.4 entries per bucket), more through a
// warp-private 4 096-bit bitmap of the bucket's trial range.  The losses land in list order in shared memory and are summed
// exactly like k_finalize_t does (same thread count, same tree): bit-identical CVaR.
constexpr int FR_THREADS = 1024, FR_MAXB = 4096, FR_BMW = 128; // buckets; bitmap words per warp (4 096 bits)
static size_t fin_radix_smem(int64_t t) { return (size_t)(FR_MAXB + 4 + FR_MAXB) * 4 + (size_t)t * 4 + (((size_t)t * 2 + 15) & ~(size_t)15) + (size_t)t * 8 + (size_t)(FR_THREADS / 32) * FR_BMW * 4 + 64; }

__global__ void __launch_bounds__(FR_THREADS, 1) k_finalize_radix(const uint64_t *kept_key, const int32_t *kept_idx, int64_t t64, int shift, int nb,
                                                                  int32_t *breach, double *cvar)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ int s_wsum[FR_THREADS / 32];
    const int t = (int)t64;
    double *sl = reinterpret_cast<double *>(smem_raw);                       // [t] losses in ascending-trial order
    int *off = reinterpret_cast<int *>(sl + t);                              // [FR_MAXB + 4] counts, then exclusive offsets
    int *cur = off + FR_MAXB + 4;                                            // [FR_MAXB] scatter cursors
    int32_t *bidx = cur + FR_MAXB;                                           // [t] trial indices, bucket by bucket
    uint16_t *bsrc = reinterpret_cast<uint16_t *>(bidx + t);                 // [t] where the entry came from
    uint32_t *bmall = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(bsrc) + (((size_t)t * 2 + 15) & ~(size_t)15));
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *ki = kept_idx + row * t64;
    const uint64_t *kk = kept_key + row * t64;
    if (ki[0] < 0) return; // row failed its optimistic bound: redone separately (k_select)
    for (int b = tid; b < FR_MAXB + 4; b += FR_THREADS) off[b] = 0;
    __syncthreads();
    // (trial indices lie in [0, n_trials) by construction; the clamp keeps anything else inside the bucket array)
    auto bucket_of = [&](int32_t id) -> int { const unsigned b = (unsigned)id >> shift; return (int)(b < (unsigned)nb ? b : (unsigned)nb - 1u); };
    for (int i = tid; i < t; i += FR_THREADS) atomicAdd(&off[bucket_of(ki[i])], 1);
    __syncthreads();
    { // exclusive scan of the bucket counts: four buckets per thread
        const int b0 = 4 * tid;
        const int c0 = off[b0], c1 = off[b0 + 1], c2 = off[b0 + 2], c3 = off[b0 + 3];
        const int c = c0 + c1 + c2 + c3;
        int incl = c;
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int run = incl - c;
        for (int w = 0; w < warp; ++w) run += s_wsum[w];
        off[b0] = run; cur[b0] = run; run += c0;
        off[b0 + 1] = run; cur[b0 + 1] = run; run += c1;
        off[b0 + 2] = run; cur[b0 + 2] = run; run += c2;
        off[b0 + 3] = run; cur[b0 + 3] = run;
        if (tid == 0) off[FR_MAXB] = t;
    }
    __syncthreads();
    for (int i = tid; i < t; i += FR_THREADS) {
        const int32_t id = ki[i];
        const int pos = atomicAdd(&cur[bucket_of(id)], 1);
        bidx[pos] = id;
        bsrc[pos] = (uint16_t)i;
    }
    __syncthreads();
    int32_t *dst = breach + row * t64;
    uint32_t *bm = bmall + warp * FR_BMW;
    const uint32_t low = (1u << shift) - 1u;
    for (int b = warp; b < nb; b += FR_THREADS / 32) {
        const int o = off[b], c = off[b + 1] - o;
        if (c == 0) continue;
        if (c <= 32) {
            const int32_t id = lane < c ? bidx[o + lane] : 0x7fffffff;
            int rank = 0;
            for (int j = 0; j < c; ++j) rank += __shfl_sync(kFullMask, id, j) < id;
            if (lane < c) {
                dst[o + rank] = id;
                sl[o + rank] = fkey_inv(kk[bsrc[o + lane]]);
            }
        } else { // a crowded bucket (clustered tail trials): rank through a bitmap of the bucket's trial range
            for (int w = lane; w < FR_BMW; w += 32) bm[w] = 0;
            __syncwarp();
            for (int e = lane; e < c; e += 32) { const uint32_t r = (uint32_t)bidx[o + e] & low; atomicOr(&bm[r >> 5], 1u << (r & 31)); }
            __syncwarp();
            const uint32_t w0 = bm[4 * lane], w1 = bm[4 * lane + 1], w2 = bm[4 * lane + 2], w3 = bm[4 * lane + 3];
            const int mine = __popc(w0) + __popc(w1) + __popc(w2) + __popc(w3);
            int incl = mine;
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(kFullMask, incl, d);
                if (lane >= d) incl += v;
            }
            const int excl = incl - mine;
            for (int e0 = 0; e0 < c; e0 += 32) {
                const int e = e0 + lane;
                const bool have = e < c;
                const int32_t id = have ? bidx[o + e] : 0;
                const uint32_t r = (uint32_t)id & low, wi = r >> 5;
                int rank = __shfl_sync(kFullMask, excl, (int)(wi >> 2));
                for (uint32_t q = wi & ~3u; q < wi; ++q) rank += __popc(bm[q]);
                rank += __popc(bm[wi] & ((1u << (r & 31)) - 1u));
                if (have) {
                    dst[o + rank] = id;
                    sl[o + rank] = fkey_inv(kk[bsrc[o + e]]);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    double s = 0.0;
    for (int i = tid; i < t; i += FR_THREADS) s += sl[i];
    s = block_sum(s, red);
    if (tid == 0) cvar[row] = s / (double)t;
}

// ============================================================ host orchestration
struct Step { int64_t begin, end; int nsub; int sub_len; };
constexpr int64_t kStepRatio = 3;   // trials seen grow 3x per step: ~2 tail sizes of new candidates per selection pass

// length of the dense first step: ~4 tail sizes, at least kDenseFloor trials, a multiple of 8 (tensor-core trial blocks)
static int64_t g_dense_floor = 1024; // process-wide tuning knob (mcp_set_option "dense_floor"); 1 024: C3 6.93 ms vs 7.17 at 2 048
void risk_set_dense_floor(int64_t v) { g_dense_floor = v < 64 ? 64 : v; }
static int64_t g_dense_mult = 4;     // tail sizes in the dense first step (mcp_set_option "dense_mult")
void risk_set_dense_mult(int64_t v) { g_dense_mult = v < 1 ? 1 : (v > 16 ? 16 : v); }
// `two_step` (optimistic schedule): the sample only has to pin down a quantile, so it is capped at the 4 096 values the
// register-resident selection handles -- for large tails the bound gets relatively tighter, not looser (r ~ t n0 / N).
static int64_t first_step_len(int64_t N, int64_t t, bool two_step = false)
{
    int64_t c0 = (two_step ? g_dense_mult : 4) * t;
    if (c0 < g_dense_floor) c0 = g_dense_floor;
    if (two_step && c0 > 16 * SEL_THREADS) c0 = 16 * SEL_THREADS;
    c0 = (c0 + 7) & ~(int64_t)7;
    return c0 < N ? c0 : N;
}

int64_t risk_first_step_len(int64_t N, int64_t t) { return first_step_len(N, t, true); }

// Trial schedule of the streaming path.  Guaranteed: a dense first step of ~4 tail sizes establishes each
// portfolio's threshold, then steps grow 3x; every later step inserts ~2 tail sizes of candidates.  Optimistic
// (`two_step`): the dense first step, then ONE sparse step over all remaining trials (see optimistic_rank).
static inline bool steps_hint_two(bool, int64_t N, int64_t e) { return e < N; } // a dense step that is followed by sparse ones

static std::vector<Step> make_schedule(int64_t N, int64_t t, int64_t nsg, int sms, bool two_step = false, int64_t kItemTrials = 2048,
                                       int ctas_per_sm = 1)
{
    std::vector<Step> steps;
    const int64_t c0 = first_step_len(N, t, two_step);
    int64_t b = 0, len = c0;
    while (b < N) {
        int64_t e = b + len;
        if (two_step && b > 0) {
            // optimistic schedule: one sparse step over the rest -- unless the sample's share of the tail is so thin
            // (t n0 / N < 16: trial-sharded runs, very high confidence levels) that its bound would let several tail
            // sizes through; then a middle step up to ~64 N / t trials tightens it first (see optimistic_rank)
            e = N;
            const int64_t mid = (64 * N / (t > 0 ? t : 1) + 7) & ~(int64_t)7;
            if (steps.size() == 1 && t * b < 16 * N && mid > 2 * b && mid <= N / 8) e = mid;
        } else if (e > N || (b > 0 && N - e < len / 4)) e = N; // fold a short remainder into the last (sparse) step
        const int64_t L = e - b;
        // the short dense step wants small work items (two or three full waves of CTAs instead of a part-filled one)
        const int64_t item = (b == 0 && steps_hint_two(two_step, N, e)) ? std::min<int64_t>(kItemTrials, 512) : kItemTrials;
        int64_t waves = (L * nsg + (int64_t)sms * item / 2) / ((int64_t)sms * item);
        if (waves < 1) waves = 1;
        int64_t nsub = (sms * waves) / nsg;
        if (nsub < 1) nsub = 1;
        if (nsub > L / 8) nsub = L / 8 > 0 ? L / 8 : 1;
        if (nsub > kSelMaxSeg / 4) nsub = kSelMaxSeg / 4;
        // Wave quantisation: a launch is nsub x nsg CTAs on sms x ctas_per_sm resident slots, every CTA works ~sl trials.  Among
        // the sub-chunk counts near the target, take the one with the smallest (waves, rounded up) x (trials per CTA): at a
        // dozen waves a last wave that is a quarter full costs 6 % (measured: C3 5.105 -> 5.036 ms per 125 000 rows).
        {
            const int64_t slots = (int64_t)sms * ctas_per_sm;
            int64_t best = nsub, best_cost = INT64_MAX;
            const int64_t lo = std::max<int64_t>(1, (2 * nsub) / 3), hi = std::min<int64_t>({(3 * nsub + 1) / 2, L / 8 > 0 ? L / 8 : 1, (int64_t)kSelMaxSeg / 4});
            for (int64_t c = lo; c <= hi; ++c) {
                const int64_t csl = (((L + c - 1) / c) + 7) & ~(int64_t)7, cn = (L + csl - 1) / csl;
                const int64_t cost = ((cn * nsg + slots - 1) / slots) * csl;
                if (cost < best_cost || (cost == best_cost && std::llabs(cn - nsub) < std::llabs(best - nsub))) { best_cost = cost; best = cn; }
            }
            nsub = best;
        }
        int64_t sl = (L + nsub - 1) / nsub;
        sl = (sl + 7) & ~(int64_t)7;
        nsub = (L + sl - 1) / sl;
        steps.push_back({b, e, (int)nsub, (int)sl});
        b = e;
        len = (kStepRatio - 1) * e; // geometric: after this step the trials seen grow by kStepRatio
    }
    return steps;
}

static int run_finalize(mcp_ctx *ctx, int64_t rows, int64_t t, const uint64_t *kk, const int32_t *ki, int32_t *breach, double *cvar,
                        int64_t n_trials = -1)
{
    if (n_trials < 0) n_trials = ctx->N;
    if (fin_bitmap_smem(n_trials, t) <= FIN_BITMAP_MAX_SMEM && !ctx->forceSortFinalize) {
        const size_t bm_smem = fin_bitmap_smem(n_trials, t);
        const bool small = ctx->smallSelect && t <= 256; // small tails: 64-thread CTAs, more rows resident per SM
        if (bm_smem > 40 * 1024) {
            MCP_CUDA(ctx, cudaFuncSetAttribute(k_finalize_bitmap_t<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_smem));
            MCP_CUDA(ctx, cudaFuncSetAttribute(k_finalize_bitmap_t<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_smem));
            MCP_CUDA(ctx, cudaFuncSetAttribute(k_finalize_bitmap_t<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_smem));
        }
        KScope ks(ctx, MCP_K_SELECT);
        if (small) k_finalize_bitmap_t<64><<<(unsigned)rows, 64, bm_smem, cur_stream(ctx)>>>(kk, ki, t, n_trials, 0, breach, cvar);
        else if (bm_smem > 72 * 1024) k_finalize_bitmap_t<1024><<<(unsigned)rows, 1024, bm_smem, cur_stream(ctx)>>>(kk, ki, t, n_trials, 0, breach, cvar);
        else k_finalize_bitmap_t<256><<<(unsigned)rows, 256, bm_smem, cur_stream(ctx)>>>(kk, ki, t, n_trials, 0, breach, cvar);
        return check_launch(ctx, "k_finalize_bitmap");
    }
    { // large tails over a trial range of up to 2^24: two counting steps instead of a bitonic sort (k_finalize_radix)
        int bits = 1;
        while (bits < 31 && ((int64_t)1 << bits) < n_trials) ++bits;
        const int shift = bits > 12 ? bits - 12 : 0;
        if (t >= 1024 && t <= 65535 && shift <= 12 && fin_radix_smem(t) <= 200 * 1024 && ctx->radixFinalize) {
            const int nb = (int)(((n_trials - 1) >> shift) + 1);
            MCP_CUDA(ctx, cudaFuncSetAttribute(k_finalize_radix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_radix_smem(t)));
            KScope ks(ctx, MCP_K_SELECT);
            k_finalize_radix<<<(unsigned)rows, FR_THREADS, fin_radix_smem(t), cur_stream(ctx)>>>(kk, ki, t, shift, nb, breach, cvar);
            return check_launch(ctx, "k_finalize_radix");
        }
    }
    int m = 1;
    while (m < t) m <<= 1;
    const size_t fin_smem = (size_t)m * 12;
    const bool fin_global = fin_smem > 200 * 1024;
    if (fin_global) {
        MCP_CUDA(ctx, ctx->dFinKey.reserve((size_t)rows * m * sizeof(uint64_t)));
        MCP_CUDA(ctx, ctx->dFinIdx.reserve((size_t)rows * m * sizeof(int32_t)));
    } else if (fin_smem > 40 * 1024) { // static shared memory counts against the 48 KB default as well
        MCP_CUDA(ctx, cudaFuncSetAttribute(k_finalize_t<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
        MCP_CUDA(ctx, cudaFuncSetAttribute(k_finalize_t<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    }
    KScope ks(ctx, MCP_K_SELECT);
    uint64_t *gk = fin_global ? ctx->dFinKey.as<uint64_t>() : nullptr;
    int32_t *gi = fin_global ? ctx->dFinIdx.as<int32_t>() : nullptr;
    if (m >= 8192) k_finalize_t<1024><<<(unsigned)rows, 1024, fin_global ? 0 : fin_smem, cur_stream(ctx)>>>(kk, ki, t, m, breach, cvar, gk, gi); // >= 96 KB: one CTA per SM
    else k_finalize_t<256><<<(unsigned)rows, 256, fin_global ? 0 : fin_smem, cur_stream(ctx)>>>(kk, ki, t, m, breach, cvar, gk, gi);
    return check_launch(ctx, "k_finalize");
}

int risk_value(mcp_ctx *ctx, int64_t p0, int64_t p1, double *d_V);

// generic factor count: dense V tile + one selection pass per portfolio
static int risk_run_dense(mcp_ctx *ctx, int64_t t)
{
    const int64_t P = ctx->P, N = ctx->N;
    size_t budget = ctx->prop.totalGlobalMem / 4;
    if (budget > ((size_t)32 << 30)) budget = (size_t)32 << 30;
    int64_t Pt = (int64_t)(budget / ((size_t)N * sizeof(double)));
    if (Pt < 1) return fail(ctx, MCP_E_NOMEM, "run: one loss row does not fit the tile budget");
    if (Pt > P) Pt = P;
    if (Pt > 65535 * (int64_t)VT) Pt = 65535 * (int64_t)VT;
    MCP_CUDA(ctx, ctx->dLoss.reserve((size_t)Pt * N * sizeof(double)));
    for (int64_t p0 = 0; p0 < P; p0 += Pt) {
        const int64_t p1 = (p0 + Pt < P) ? p0 + Pt : P;
        MCP_TRY(risk_value(ctx, p0, p1, ctx->dLoss.as<double>()));
        SelArgs a{};
        a.cand = ctx->dLoss.as<double>();
        a.pitch = N * (int64_t)sizeof(double);
        a.nsub = 1;
        a.sub_len = N;
        a.step_len = N;
        a.n_begin = 0;
        a.kept_n = 0;
        a.kept_key_out = ctx->dKeptKey.as<uint64_t>() + p0 * t;
        a.kept_idx_out = ctx->dKeptIdx.as<int32_t>() + p0 * t;
        a.tail = t;
        a.moments_n = N;
        a.mean = ctx->dMean.as<double>() + p0;
        a.variance = ctx->dVar.as<double>() + p0;
        a.var_loss = ctx->dVarLoss.as<double>() + p0;
        a.var_trial = ctx->dVarTrial.as<int32_t>() + p0;
        {
            KScope ks(ctx, MCP_K_SELECT);
            a.cap = ctx->selCap;
            k_select_t<256><<<(unsigned)(p1 - p0), 256, sel_smem_bytes(a.cap, a.nsub), ctx->stream>>>(a);
            MCP_TRY(check_launch(ctx, "k_select"));
        }
        MCP_TRY(run_finalize(ctx, p1 - p0, t, a.kept_key_out, a.kept_idx_out, ctx->dBreach.as<int32_t>() + p0 * t,
                             ctx->dCvar.as<double>() + p0));
    }
    return MCP_OK;
}

template <int NK>
static int launch_value_mma(mcp_ctx *ctx, const StreamArgs &v, bool dense, dim3 grid)
{
    const bool configured = (ctx->mmaConfigured >> NK) & 1u; // per context: the attribute is per device
    if (!configured) {
        const int sm = (int)vm_smem_bytes<NK>();
        MCP_CUDA(ctx, cudaFuncSetAttribute(k_value_mma<NK, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        MCP_CUDA(ctx, cudaFuncSetAttribute(k_value_mma<NK, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        MCP_CUDA(ctx, cudaFuncSetAttribute(k_value_mma<NK, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        MCP_CUDA(ctx, cudaFuncSetAttribute(k_value_mma<NK, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        ctx->mmaConfigured |= 1u << NK;
    }
    KScope ks(ctx, MCP_K_VALUE);
    const size_t sm = vm_smem_bytes<NK>();
    const bool mom = v.part != nullptr;
    if (dense && mom) k_value_mma<NK, true, true><<<grid, VM_THREADS, sm, cur_stream(ctx)>>>(v);
    else if (dense) k_value_mma<NK, true, false><<<grid, VM_THREADS, sm, cur_stream(ctx)>>>(v);
    else if (mom) k_value_mma<NK, false, true><<<grid, VM_THREADS, sm, cur_stream(ctx)>>>(v);
    else k_value_mma<NK, false, false><<<grid, VM_THREADS, sm, cur_stream(ctx)>>>(v);
    return check_launch(ctx, "k_value_mma");
}

static bool mma_supported(int32_t F) { return F == 16 || F == 32 || F == 48 || F == 64; }

// (re)build the fragment-order copy of the scenario matrix when it is stale
// With a pipelined upload (mcp_simulate) only the first `sHeadRows` scenario rows are on the device when the dense
// step starts; the rest arrives on the side stream and is repacked by scenarios_tail_ready() before the first kernel
// that touches those trials.
static int ensure_sfrag(mcp_ctx *ctx)
{
    if (ctx->sfragValid) return MCP_OK;
    const int NK = ctx->F / 4;
    const int64_t nblk = (ctx->N + 7) / 8;
    MCP_CUDA(ctx, ctx->dSfrag.reserve((size_t)nblk * NK * 32 * sizeof(double) + 16));
    const int64_t b1 = ctx->sTailPending ? ctx->sHeadRows / 8 : nblk;
    KScope ks(ctx, MCP_K_GEN);
    k_repack_scenarios<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(ctx->dS.as<double>(), ctx->N, NK, ctx->dSfrag.as<double>(), 0, b1);
    MCP_TRY(check_launch(ctx, "k_repack_scenarios"));
    ctx->sfragValid = !ctx->sTailPending;
    return MCP_OK;
}

static int scenarios_tail_ready(mcp_ctx *ctx)
{
    if (!ctx->sTailPending) return MCP_OK;
    MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evTail, 0));
    ctx->sTailPending = false;
    const int NK = ctx->F / 4;
    KScope ks(ctx, MCP_K_GEN, 1, ctx->stream);
    k_repack_scenarios<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(ctx->dS.as<double>(), ctx->N, NK, ctx->dSfrag.as<double>(),
                                                                                  ctx->sHeadRows / 8, (ctx->N + 7) / 8);
    MCP_TRY(check_launch(ctx, "k_repack_scenarios (tail)"));
    ctx->sfragValid = true;
    return MCP_OK;
}

int64_t risk_first_step_len(int64_t N, int64_t t);

static int dispatch_value_mma(mcp_ctx *ctx, const StreamArgs &v, bool dense, dim3 grid)
{
    switch (ctx->F) {
    case 16: return launch_value_mma<4>(ctx, v, dense, grid);
    case 32: return launch_value_mma<8>(ctx, v, dense, grid);
    case 48: return launch_value_mma<12>(ctx, v, dense, grid);
    case 64: return launch_value_mma<16>(ctx, v, dense, grid);
    }
    return fail(ctx, MCP_E_UNSUPPORTED, "no tensor-core valuation kernel for this factor count");
}

int risk_value(mcp_ctx *ctx, int64_t p0, int64_t p1, double *d_V)
{
    if (p1 <= p0) return MCP_OK;
    if (ctx->valueMma && mma_supported(ctx->F)) {
        // the streaming kernel in DENSE mode writes V itself: [p - p0][N]
        const int VM_PG = vm_pg(ctx->F);
        const int64_t nsg = (p1 - p0 + VM_PG - 1) / VM_PG;
        std::vector<Step> st = make_schedule(ctx->N, ctx->N, nsg, ctx->prop.multiProcessorCount); // one step: [0, N)
        MCP_TRY(ensure_sfrag(ctx));
        StreamArgs v{};
        v.E = ctx->dE.as<double>(); v.S = ctx->dS.as<double>(); v.Sfrag = ctx->dSfrag.as<double>();
        v.p0 = p0; v.p1 = p1;
        v.n_begin = 0; v.n_end = ctx->N;
        v.nsub = st[0].nsub; v.sub_len = st[0].sub_len;
        v.cand = d_V; v.pitch = ctx->N * (int64_t)sizeof(double);
        dim3 grid((unsigned)v.nsub, (unsigned)nsg);
        return dispatch_value_mma(ctx, v, true, grid);
    }
    dim3 grid((unsigned)((ctx->N + VT - 1) / VT), (unsigned)((p1 - p0 + VT - 1) / VT));
    KScope ks(ctx, MCP_K_VALUE);
    k_value_generic<<<grid, 256, 0, ctx->stream>>>(ctx->dE.as<double>(), ctx->dS.as<double>(), p0, p1, ctx->N, ctx->F, d_V);
    return check_launch(ctx, "k_value_generic");
}

// mean / variance of every portfolio by linearity (see "moments by linearity" above)
static int run_moments_linear(mcp_ctx *ctx)
{
    const int64_t N = ctx->N, P = ctx->P;
    const int F = ctx->F;
    if (F > 64) return fail(ctx, MCP_E_UNSUPPORTED, "linear moments: more than 64 factors");
    const int nb = (int)((N + GM_ROWS - 1) / GM_ROWS);
    MCP_CUDA(ctx, ctx->dGram.reserve(((size_t)nb * F * F + (size_t)nb * F + (size_t)F * F + F) * sizeof(double)));
    double *pg = ctx->dGram.as<double>(), *ps = pg + (size_t)nb * F * F, *cov = ps + (size_t)nb * F, *mu = cov + (size_t)F * F;
    cudaStream_t st = ctx->stream2; // forked from / joined to ctx->stream by risk_run
    KScope ks(ctx, MCP_K_MOMENTS, 5, st);
    k_col_sums<<<nb, GM_THREADS, 0, st>>>(ctx->dS.as<double>(), N, F, ps);
    k_col_mean<<<(F + 63) / 64, 64, 0, st>>>(ps, nb, F, N, mu);
    k_gram_centered<<<nb, GM_THREADS, 32 * F * sizeof(double), st>>>(ctx->dS.as<double>(), N, F, mu, pg);
    k_gram_reduce<<<(F * F + 255) / 256, 256, 0, st>>>(pg, nb, F * F, N, cov);
    {
        int64_t g = (P + 7) / 8;
        const int64_t gmax = (int64_t)ctx->prop.multiProcessorCount * 4;
        if (g > gmax) g = gmax;
        const size_t sm = ((size_t)F * F + F) * sizeof(double);
        k_moments_linear<<<(unsigned)g, 256, sm, st>>>(ctx->dE.as<double>(), P, F, mu, cov, ctx->dMean.as<double>(),
                                                                ctx->dVar.as<double>());
    }
    return check_launch(ctx, "moments_linear");
}

// ---- optimistic filter bound
// The guaranteed schedule (make_schedule) filters step i+1 with the exact t-th largest loss of the trials seen so
// far: always sound, but it needs log_3(N / 4t) valuation launches and as many selection passes.  For trials in an
// exchangeable order (Monte Carlo draws are) a far better bound is available after the dense first step alone: the
// r-th largest loss of the n0-trial sample, r = ts + 7 sqrt(ts) + 2 with ts = t n0 / N the sample's expected share of
// the tail.  The number of the other N - n0 trials above it is negative-hypergeometric; fewer than t - r of them --
// the only way the bound can be too tight -- has probability < 1e-10 per portfolio for the configurations of
// BASELINE.json (C2: r = 87 of 4 004, about 2.2 t candidates).  The bound is a bet, never a correctness assumption:
// k_select counts every row's candidates, rows with fewer than t are collected (`fail_rows`), re-run through the
// guaranteed schedule on a gathered copy of their exposures and scattered back, so the result is exact for ANY trial
// order (descending-loss order makes every row fail; the tests do exactly that).
static int64_t optimistic_rank(int64_t N, int64_t n0, int64_t t)
{
    const double ts = (double)t * (double)n0 / (double)N;
    return (int64_t)ceil(ts + 7.0 * sqrt(ts) + 2.0);
}

// rows of E selected by `rows` -> a compact matrix (fallback of the optimistic bound)
__global__ void k_gather_rows(const double *__restrict__ E, const int32_t *__restrict__ rows, int64_t n, int F, double *__restrict__ out)
{
    const int64_t total = n * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = E[(int64_t)rows[i / F] * F + i % F];
}

// results of the compact fallback problem -> their rows in the full result arrays (one CTA per row)
__global__ void k_scatter_rows(const int32_t *__restrict__ rows, int64_t t, const double *vl, const int32_t *vt, const double *cv,
                               const int32_t *br, double *var_loss, int32_t *var_trial, double *cvar, int32_t *breach)
{
    const int64_t i = blockIdx.x, r = rows[i];
    if (threadIdx.x == 0) { var_loss[r] = vl[i]; var_trial[r] = vt[i]; cvar[r] = cv[i]; }
    for (int64_t j = threadIdx.x; j < t; j += blockDim.x) breach[r * t + j] = br[i * t + j];
}

// One streaming problem: exposures E[P][F] (device) against the context's scenarios, results into `o`.
struct StreamIO {
    const double *E;
    int64_t P;
    double *var_loss; int32_t *var_trial; double *cvar; int32_t *breach; // [P], [P], [P], [P][t]
    uint64_t *kkey[2]; int32_t *kidx[2];                                 // kept-tail ping-pong, [P][t] each
    uint64_t *tkey;                                                      // [P]
    bool moments;                                                        // fused per-trial moments into ctx->dPart
    int *fail_cnt; int32_t *fail_rows;                                   // non-null: optimistic schedule
    bool keep_only;                                                      // stop after the selection: the kept tail is the result
    int final_buf;                                                       // (out) which of kkey/kidx holds the kept tail
};

// streaming valuation (FP64 tensor cores) + threshold filter + radix-select; `optimistic` picks the schedule
static int risk_stream_core(mcp_ctx *ctx, StreamIO &io, int64_t t, bool optimistic, int *nparts_out)
{
    const int64_t P = io.P, N = ctx->N;
    const int sms = ctx->prop.multiProcessorCount;
    // portfolio tile: candidate buffers are Pt x (longest step) x 16 bytes; keep them under 1/4 of HBM
    size_t budget = ctx->prop.totalGlobalMem / 4;
    if (budget > ((size_t)40 << 30)) budget = (size_t)40 << 30;
    int64_t Pt = P;
    // overlap_tiles > 1 (off by default): portfolio tiles alternate between two streams, so that the selection / compaction
    // kernels of one tile (latency-bound, small CTAs) run under the valuation kernel of the next instead of after it.
    const int VM_PG = vm_pg(ctx->F);
    const int64_t groups = (P + VM_PG - 1) / VM_PG;
    const int64_t want_tiles = ctx->overlapTiles > 1 && groups >= 2 * ctx->overlapTiles ? ctx->overlapTiles : 1;
    if (want_tiles > 1) Pt = ((groups + want_tiles - 1) / want_tiles) * VM_PG;
    std::vector<Step> steps;
    int64_t stride = 0, rank0 = t;
    int max_nsub = 0, nparts = 0;
    for (;;) {
        const int64_t nsg = (Pt + VM_PG - 1) / VM_PG;
        steps = make_schedule(N, t, nsg, sms, optimistic, ctx->itemTrials, ctx->F <= 48 ? 2 : 1); // (two 8-warp valuation CTAs per SM up to F = 48)
        stride = 0; max_nsub = 0; nparts = 0;
        for (auto &s : steps) {
            const int64_t w = (int64_t)s.nsub * s.sub_len;
            if (w > stride) stride = w;
            if (s.nsub > max_nsub) max_nsub = s.nsub;
            nparts += s.nsub;
        }
        stride = (stride + 3) & ~(int64_t)3;
        if ((size_t)Pt * stride * 16 <= budget || Pt <= VM_PG) break;
        Pt = ((Pt / 2 + VM_PG - 1) / VM_PG) * VM_PG;
    }
    if (optimistic) rank0 = optimistic_rank(N, steps[0].end, t);
    if (nparts_out) *nparts_out = nparts;
    if (Pt > 65535 * (int64_t)VM_PG) Pt = 65535 * (int64_t)VM_PG;
    const int nstreams = (P > Pt && ctx->overlapTiles > 1) ? 2 : 1;
    mcp::DevBuf *candKey[2] = {&ctx->dCandKey, &ctx->dCandKey2}, *candCnt[2] = {&ctx->dCandCnt, &ctx->dCandCnt2};
    for (int i = 0; i < nstreams; ++i) {
        MCP_CUDA(ctx, candKey[i]->reserve((size_t)Pt * stride * 16));
        MCP_CUDA(ctx, candCnt[i]->reserve((size_t)Pt * max_nsub * 4 * sizeof(int32_t)));
    }
    if (io.moments) MCP_CUDA(ctx, ctx->dPart.reserve((size_t)P * nparts * 2 * sizeof(double)));
    MCP_CUDA(ctx, ctx->dOvf.reserve((size_t)2 * (Pt + 16) * sizeof(int32_t))); // rows handed back by the warp-per-row selection
    MCP_TRY(ensure_sfrag(ctx));
    cudaStream_t tile_stream[2] = {ctx->stream, ctx->stream3};
    if (nstreams == 2) {
        MCP_CUDA(ctx, cudaEventRecord(ctx->evTileFork, ctx->stream));
        MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->evTileFork, 0));
    }
    bool tail_event = false, tail_waited = false; // pipelined upload: the second stream waits for the tail repack too
    struct CurReset { mcp_ctx *c; ~CurReset() { c->cur = nullptr; } } cur_reset{ctx};

    int64_t tile = 0;
    for (int64_t p0 = 0; p0 < P; p0 += Pt, ++tile) {
        const int64_t p1 = (p0 + Pt < P) ? p0 + Pt : P;
        const int64_t rows = p1 - p0;
        const int sidx_ = (int)(tile % nstreams);
        ctx->cur = tile_stream[sidx_];
        uint64_t *kkey[2] = {io.kkey[0] + p0 * t, io.kkey[1] + p0 * t};
        int32_t *kidx[2] = {io.kidx[0] + p0 * t, io.kidx[1] + p0 * t};
        int cur = 0, part_base = 0;
        int64_t kept = 0; // entries per row of the kept tail so far
        for (size_t si = 0; si < steps.size(); ++si) {
            const Step &s = steps[si];
            const bool dense = si == 0;
            int split_a = 0; // > 0: this step's valuation is launched in two parts, sub-chunks [0, split_a) first
            if (ctx->sTailPending && s.end > ctx->sHeadRows && ctx->sMidRows > s.begin && nstreams == 1 && P <= Pt && !dense &&
                s.begin >= ctx->sHeadRows) {
                const int64_t na = (ctx->sMidRows - s.begin) / s.sub_len; // sub-chunks whose trials all lie below sMidRows
                if (na > 0 && na < s.nsub) {
                    split_a = (int)na;
                    MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evTail1, 0));
                    ctx->sTailPending = false;
                    KScope ks(ctx, MCP_K_GEN, 1, ctx->stream);
                    k_repack_scenarios<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(ctx->dS.as<double>(), ctx->N, ctx->F / 4, ctx->dSfrag.as<double>(),
                                                                                                  ctx->sHeadRows / 8, ctx->sMidRows / 8);
                    MCP_TRY(check_launch(ctx, "k_repack_scenarios (tail, first part)"));
                }
            }
            if (ctx->sTailPending && s.end > ctx->sHeadRows) {
                MCP_TRY(scenarios_tail_ready(ctx)); // on ctx->stream
                MCP_CUDA(ctx, cudaEventRecord(ctx->evSfrag, ctx->stream));
                tail_event = true;
            }
            if (sidx_ == 1 && tail_event && !tail_waited && s.end > ctx->sHeadRows) {
                MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->evSfrag, 0));
                tail_waited = true;
            }
            StreamArgs v{};
            v.E = io.E;
            v.S = ctx->dS.as<double>();
            v.Sfrag = ctx->dSfrag.as<double>();
            v.p0 = p0; v.p1 = p1;
            v.n_begin = s.begin; v.n_end = s.end;
            v.nsub = s.nsub; v.sub_len = s.sub_len;
            v.tkey = io.tkey;
            v.cand = candKey[sidx_]->p;
            v.pitch = stride * 16;
            v.cnt = candCnt[sidx_]->as<int32_t>();
            v.part = io.moments ? ctx->dPart.as<double>() : nullptr;
            v.part_stride = nparts; v.part_base = part_base;
            dim3 grid((unsigned)s.nsub, (unsigned)((rows + VM_PG - 1) / VM_PG));
            if (split_a > 0) {
                // pipelined upload in two parts: the sub-chunks whose trials arrived first now, the others behind the rest
                grid.x = (unsigned)split_a;
                MCP_TRY(dispatch_value_mma(ctx, v, dense, grid));
                MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evTail, 0));
                {
                    KScope ks(ctx, MCP_K_GEN, 1, ctx->stream);
                    k_repack_scenarios<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(ctx->dS.as<double>(), ctx->N, ctx->F / 4, ctx->dSfrag.as<double>(),
                                                                                                  ctx->sMidRows / 8, (ctx->N + 7) / 8);
                    MCP_TRY(check_launch(ctx, "k_repack_scenarios (tail, second part)"));
                }
                ctx->sfragValid = true;
                v.sub0 = split_a;
                grid.x = (unsigned)(s.nsub - split_a);
            }
            MCP_TRY(dispatch_value_mma(ctx, v, dense, grid));
            SelArgs a{};
            a.cand = v.cand;
            a.pitch = v.pitch;
            a.cnt = dense ? nullptr : v.cnt;
            // sparse steps: every sub-chunk holds four lane-private segments of sub_len/4 slots
            a.nsub = dense ? s.nsub : 4 * s.nsub;
            a.sub_len = dense ? s.sub_len : s.sub_len / 4;
            a.step_len = s.end - s.begin;
            a.n_begin = s.begin;
            a.kept_key_in = kkey[cur];
            a.kept_idx_in = kidx[cur];
            a.kept_n = kept;
            a.kept_key_out = kkey[cur ^ 1];
            a.kept_idx_out = kidx[cur ^ 1];
            // optimistic schedule: the dense step keeps only the top `rank0` of the sample (its rank0-th largest loss
            // is the bound of the one sparse step); every other selection keeps the full tail
            a.tail = (optimistic && si + 1 < steps.size()) ? (dense ? rank0 : optimistic_rank(N, s.end, t)) : t;
            a.moments_n = 0;
            a.var_loss = io.var_loss + p0;
            a.var_trial = io.var_trial + p0;
            a.tkey = io.tkey + p0;
            a.fail_cnt = io.fail_cnt;
            a.fail_rows = io.fail_rows;
            a.row_base = p0;
            {
                KScope ks(ctx, MCP_K_SELECT);
                // sparse steps stage kept + ~2.2 t candidates (optimistic) or ~2 t (guaranteed): a smaller staging area
                // than the dense step's 4 t buys a resident CTA per SM more; rows beyond it take the global path
                a.cap = dense ? ctx->selCap : (int)std::min<int64_t>(ctx->selCap, std::max<int64_t>(2304, 2 * t + 900 + ctx->selSlack));
                const bool reg = dense && ctx->regSelect && a.tail <= a.step_len; // the row fits the CTA's registers
                // sparse steps of small rows: one WARP per row; the rows that do not fit come back in a list and take the kernels below
                const bool wsparse = !dense && ctx->warpSelect && a.kept_segs <= 1 && a.nsub <= WS_MAXSEG && a.tail <= WS_CAP &&
                                     a.kept_n + 4 * t + 64 <= WS_CAP;
                if (wsparse) {
                    int *ovf_cnt = ctx->dOvf.as<int>() + (size_t)sidx_ * (Pt + 16);
                    int32_t *ovf_rows = ovf_cnt + 4;
                    MCP_CUDA(ctx, cudaMemsetAsync(ovf_cnt, 0, sizeof(int), cur_stream(ctx)));
                    k_wsel_sparse<<<(unsigned)((rows + WS_WARPS - 1) / WS_WARPS), WS_WARPS * 32, 0, cur_stream(ctx)>>>(a, rows, ovf_cnt, ovf_rows);
                    MCP_TRY(check_launch(ctx, "k_wsel_sparse"));
                    a.rows = ovf_rows;
                    a.nrows = ovf_cnt;
                }
                if (wsparse) {
                    // the rows handed back (usually none): a grid-stride walk of a small grid over the list
                    const unsigned sgrid = (unsigned)std::min<int64_t>(rows, 2 * ctx->prop.multiProcessorCount);
                    if (ctx->smallSelect && a.kept_n + 4 * t + 64 <= 1024) {
                        a.cap = 1024;
                        k_select_t_list<64><<<sgrid, 64, sel_smem_bytes(a.cap, a.nsub, 64), cur_stream(ctx)>>>(a);
                    } else if (sel_smem_bytes(a.cap, a.nsub) > 110 * 1024)
                        k_select_t_list<1024><<<sgrid, 1024, sel_smem_bytes(a.cap, a.nsub), cur_stream(ctx)>>>(a);
                    else
                        k_select_t_list<256><<<sgrid, 256, sel_smem_bytes(a.cap, a.nsub), cur_stream(ctx)>>>(a);
                } else if (reg && a.step_len <= 8 * 128) k_select_dense<8, 128><<<(unsigned)rows, 128, 0, cur_stream(ctx)>>>(a);
                else if (reg && a.step_len <= 16 * 128) k_select_dense<16, 128><<<(unsigned)rows, 128, 0, cur_stream(ctx)>>>(a);
                else if (reg && a.step_len <= 16 * SEL_THREADS) k_select_dense<16, 256><<<(unsigned)rows, 256, 0, cur_stream(ctx)>>>(a);
                else if (!dense && ctx->smallSelect && a.kept_n + 4 * t + 64 <= 1024) {
                    // small tails (t ~ 100): a 64-thread CTA per row, a staging area to match (rows beyond it take the
                    // global-memory path): a dozen rows resident per SM instead of four
                    a.cap = 1024;
                    k_select_t<64><<<(unsigned)rows, 64, sel_smem_bytes(a.cap, a.nsub, 64), cur_stream(ctx)>>>(a);
                } else if (sel_smem_bytes(a.cap, a.nsub) > 110 * 1024)
                    // large tails: the staging area leaves room for ONE CTA per SM, so make it 1 024 threads wide
                    k_select_t<1024><<<(unsigned)rows, 1024, sel_smem_bytes(a.cap, a.nsub), cur_stream(ctx)>>>(a);
                else
                    k_select_t<256><<<(unsigned)rows, 256, sel_smem_bytes(a.cap, a.nsub), cur_stream(ctx)>>>(a);
                MCP_TRY(check_launch(ctx, "k_select"));
            }
            kept = a.tail;
            cur ^= 1;
            part_base += s.nsub;
        }
        io.final_buf = cur;
        if (!io.keep_only) MCP_TRY(run_finalize(ctx, rows, t, kkey[cur], kidx[cur], io.breach + p0 * t, io.cvar + p0));
    }
    ctx->cur = nullptr;
    if (nstreams == 2) {
        MCP_CUDA(ctx, cudaEventRecord(ctx->evTileJoin, ctx->stream3));
        MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evTileJoin, 0));
    }
    return MCP_OK;
}

static int risk_run_stream(mcp_ctx *ctx, int64_t t)
{
    const int64_t P = ctx->P, N = ctx->N;
    MCP_CUDA(ctx, ctx->dTau.reserve((size_t)P * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dKept2Key.reserve((size_t)P * t * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dKept2Idx.reserve((size_t)P * t * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dFail.reserve((size_t)(P + 4) * sizeof(int32_t)));
    int *fail_cnt = ctx->dFail.as<int>();
    int32_t *fail_rows = ctx->dFail.as<int32_t>() + 4;
    // the optimistic bound needs a sample that is a small part of the trials and a rank below the tail size
    const int64_t n0 = first_step_len(N, t, true);
    const bool optimistic = ctx->optimistic && N >= 2 * n0 && optimistic_rank(N, n0, t) < t && 2 * optimistic_rank(N, n0, t) <= n0;
    StreamIO io{};
    io.E = ctx->dE.as<double>(); io.P = P;
    io.var_loss = ctx->dVarLoss.as<double>(); io.var_trial = ctx->dVarTrial.as<int32_t>();
    io.cvar = ctx->dCvar.as<double>(); io.breach = ctx->dBreach.as<int32_t>();
    io.kkey[0] = ctx->dKeptKey.as<uint64_t>(); io.kkey[1] = ctx->dKept2Key.as<uint64_t>();
    io.kidx[0] = ctx->dKeptIdx.as<int32_t>(); io.kidx[1] = ctx->dKept2Idx.as<int32_t>();
    io.tkey = ctx->dTau.as<uint64_t>();
    io.moments = ctx->fusedMoments;
    if (optimistic) {
        MCP_CUDA(ctx, cudaMemsetAsync(fail_cnt, 0, sizeof(int), ctx->stream));
        io.fail_cnt = fail_cnt; io.fail_rows = fail_rows;
    }
    int nparts = 0;
    MCP_TRY(risk_stream_core(ctx, io, t, optimistic, &nparts));
    if (ctx->fusedMoments) {
        KScope ks(ctx, MCP_K_MOMENTS);
        k_moments_finalize<<<(unsigned)((P + 7) / 8), 256, 0, ctx->stream>>>(ctx->dE.as<double>(), ctx->dS.as<double>(), P, ctx->F, N,
                                                                            ctx->dPart.as<double>(), nparts, nparts,
                                                                            ctx->dMean.as<double>(), ctx->dVar.as<double>());
        MCP_TRY(check_launch(ctx, "k_moments_finalize"));
    }
    ctx->runFallbackRows = 0;
    ctx->failPending = false;
    if (optimistic) {
        // The count of rows that lost the bet is read LATER, together with the encoded size (risk_run_inner): the
        // encode is enqueued speculatively, so the common case (no such row) has no host round trip in the middle.
        MCP_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(ctx->hPinned) + 64, fail_cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->failPending = true;
    }
    return MCP_OK;
}

// redo the rows whose optimistic bound was too tight with the guaranteed schedule on a compact copy of their exposures
static int risk_stream_fallback(mcp_ctx *ctx, int64_t t, int64_t nf)
{
    const int F = ctx->F;
    int32_t *fail_rows = ctx->dFail.as<int32_t>() + 4;
    MCP_CUDA(ctx, ctx->dFbE.reserve((size_t)nf * F * sizeof(double)));
    MCP_CUDA(ctx, ctx->dFbOut.reserve((size_t)nf * (3 * sizeof(double) + (size_t)(t + 1) * sizeof(int32_t)) + 64));
    MCP_CUDA(ctx, ctx->dFbKey.reserve((size_t)nf * t * 2 * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dFbIdx.reserve((size_t)nf * t * 2 * sizeof(int32_t)));
    {
        KScope ks(ctx, MCP_K_SELECT);
        k_gather_rows<<<(unsigned)std::min<int64_t>((nf * F + 255) / 256, 4096), 256, 0, ctx->stream>>>(ctx->dE.as<double>(), fail_rows, nf, F,
                                                                                                      ctx->dFbE.as<double>());
        MCP_TRY(check_launch(ctx, "k_gather_rows"));
    }
    StreamIO fb{};
    fb.E = ctx->dFbE.as<double>(); fb.P = nf;
    // dFbOut: [var_loss nf f64][cvar nf f64][tkey nf u64][breach nf*t i32][var_trial nf i32]
    double *o = ctx->dFbOut.as<double>();
    fb.var_loss = o; fb.cvar = o + nf; fb.tkey = reinterpret_cast<uint64_t *>(o + 2 * nf);
    fb.breach = reinterpret_cast<int32_t *>(o + 3 * nf); fb.var_trial = fb.breach + nf * t;
    fb.kkey[0] = ctx->dFbKey.as<uint64_t>(); fb.kkey[1] = fb.kkey[0] + nf * t;
    fb.kidx[0] = ctx->dFbIdx.as<int32_t>(); fb.kidx[1] = fb.kidx[0] + nf * t;
    MCP_TRY(risk_stream_core(ctx, fb, t, false, nullptr));
    KScope ks(ctx, MCP_K_SELECT);
    k_scatter_rows<<<(unsigned)nf, 128, 0, ctx->stream>>>(fail_rows, t, fb.var_loss, fb.var_trial, fb.cvar, fb.breach, ctx->dVarLoss.as<double>(),
                                                         ctx->dVarTrial.as<int32_t>(), ctx->dCvar.as<double>(), ctx->dBreach.as<int32_t>());
    return check_launch(ctx, "k_scatter_rows");
}

// breach lists of uniform length t -> CSR of encoded lists: the single-pass batched kernels when the codec and the list
// length allow, else size pass + scan + emit pass of the warp-per-list kernels
static int encode_uniform_lists(mcp_ctx *ctx, int codec_id, int framing, const int32_t *d_lists, int64_t t, int64_t rows, mcp::DevBuf &off,
                                mcp::DevBuf &enc, int64_t *total)
{
    if (ctx->fastCodec && codec_fast_applicable(codec_id, rows * t, rows, true)) {
        const int64_t bound = codec_fast_max_bytes(codec_id, rows * t, rows);
        MCP_CUDA(ctx, enc.reserve((size_t)bound + 16));
        const int rc = codec_fast_encode(ctx, codec_id, framing, d_lists, nullptr, t, rows, rows * t, off.as<int64_t>(), enc.as<uint8_t>(), bound, total);
        if (rc != MCP_E_UNSUPPORTED) return rc;
    }
    MCP_TRY(codec_encode_plan(ctx, codec_id, framing, d_lists, nullptr, t, rows, off.as<int64_t>(), total));
    MCP_CUDA(ctx, enc.reserve((size_t)*total + 16));
    return codec_encode_emit(ctx, codec_id, framing, d_lists, nullptr, t, rows, off.as<int64_t>(), enc.as<uint8_t>());
}

// ============================================================ trial-sharded mode (BASELINE config C5)
// Rank g holds ALL P exposure rows and the scenario rows [trial_base, trial_base + N_local) of an N_global-trial run.
// Per portfolio the global tail (t largest losses of N_global) is found from the ranks' LOCAL top-m candidates,
// m ~ t/G + 6 sigma, exchanged by ONE all-gather of fixed-stride blocks:
//     block = [header 64 B][keys P*m u64][mean P f64][variance P f64][global trial idx P*m i32]
// Every rank then merges all P rows from the gathered blocks (cheap next to the valuation), which makes the
// exactness test deterministic and identical on all ranks without another exchange: a row is exact unless some rank
// g had more candidates than it sent AND its smallest sent candidate still beats the merged threshold (then rank g may
// hold unsent trials of the tail).  Such rows -- only adversarial trial orders produce them -- go through a second
// round with the full local tails (m = t), so the result is exact for any input.  Finalisation (breach list, CVaR,
// moments by Chan's pooled-variance rule, encoding) is done per rank for its portfolio slice [P g/G, P (g+1)/G).
// The exchange itself is the caller's (mcp_ts_* phase calls: any transport) or NCCL's (mcp_run_trial_sharded).
struct TsLayout { int64_t P, m, off_keys, off_mean, off_var, off_idx, bytes; };
static TsLayout ts_layout(int64_t P, int64_t m)
{
    TsLayout L{};
    L.P = P; L.m = m;
    L.off_keys = 64;
    L.off_mean = L.off_keys + P * m * 8;
    L.off_var = L.off_mean + P * 8;
    L.off_idx = L.off_var + P * 8;
    L.bytes = (L.off_idx + P * m * 4 + 15) & ~(int64_t)15;
    return L;
}

int64_t ts_candidates_per_rank(int64_t n_global, int64_t t, int world)
{
    const int64_t nmax = (n_global + world - 1) / world; // largest shard of a balanced split
    const double share = (double)t * (double)nmax / (double)n_global;
    int64_t m = (int64_t)ceil(share + 6.0 * sqrt(share * (1.0 - 1.0 / world)) + 8.0);
    if (m > t) m = t;
    if (m > nmax) m = nmax;
    return m < 1 ? 1 : m;
}

// kept tail of the local selection -> this rank's exchange block (trial indices made global; rows with fewer than m
// local trials are padded with key 0, which sorts below every real loss)
__global__ void k_ts_pack(const uint64_t *kkey, const int32_t *kidx, int64_t P, int64_t m, int64_t m_have, int64_t trial_base,
                          int64_t n_local, const double *mean, const double *var, unsigned char *block, TsLayout L)
{
    const int64_t total = P * m;
    uint64_t *ok = reinterpret_cast<uint64_t *>(block + L.off_keys);
    int32_t *oi = reinterpret_cast<int32_t *>(block + L.off_idx);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / m, j = i - p * m;
        const bool have = j < m_have;
        ok[i] = have ? kkey[p * m_have + j] : 0ull;
        oi[i] = have ? (int32_t)(kidx[p * m_have + j] + trial_base) : -1;
    }
    if (mean) {
        double *om = reinterpret_cast<double *>(block + L.off_mean), *ov = reinterpret_cast<double *>(block + L.off_var);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) { om[i] = mean[i]; ov[i] = var[i]; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t *h = reinterpret_cast<int64_t *>(block);
        h[0] = n_local; h[1] = m; h[2] = P; h[3] = trial_base;
    }
}

// one warp per row: exact unless a rank that held back candidates has its smallest sent one above the threshold
__global__ void __launch_bounds__(256) k_ts_check(const unsigned char *gath, int world, TsLayout L, int64_t row0, int64_t rows,
                                                  const double *var_loss, const int32_t *var_trial, unsigned char *flags, int *nflag)
{
    const int lane = threadIdx.x & 31;
    const int64_t rl = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (rl >= rows) return;
    const int64_t r = row0 + rl; // global row: this rank checks its own portfolio slice
    const uint64_t tk = fkey(var_loss[r]);
    const uint32_t ti = (uint32_t)var_trial[r];
    bool bad = false;
    for (int g = 0; g < world; ++g) {
        const unsigned char *blk = gath + (int64_t)g * L.bytes;
        const int64_t n_local = reinterpret_cast<const int64_t *>(blk)[0];
        if (L.m >= n_local) continue; // the rank sent every trial it has
        const uint64_t *k = reinterpret_cast<const uint64_t *>(blk + L.off_keys) + r * L.m;
        const int32_t *ix = reinterpret_cast<const int32_t *>(blk + L.off_idx) + r * L.m;
        uint64_t mk = ~0ull;
        uint32_t mi = 0xffffffffu;
        for (int64_t j = lane; j < L.m; j += 32) {
            const uint64_t kk = k[j];
            const uint32_t ii = (uint32_t)ix[j];
            if (kk < mk || (kk == mk && ii < mi)) { mk = kk; mi = ii; }
        }
        for (int d = 16; d; d >>= 1) {
            const uint64_t ok = __shfl_xor_sync(kFullMask, mk, d);
            const uint32_t oi = __shfl_xor_sync(kFullMask, mi, d);
            if (ok < mk || (ok == mk && oi < mi)) { mk = ok; mi = oi; }
        }
        if (mk > tk || (mk == tk && mi > ti)) bad = true;
    }
    if (lane == 0) {
        flags[rl] = bad ? 1 : 0;
        if (bad) atomicAdd(nflag, 1);
    }
}

// Chan et al. pooled mean / population variance of the ranks' per-shard moments, rows [p0, p1)
__global__ void k_ts_moments_merge(const unsigned char *gath, int world, TsLayout L, int64_t p0, int64_t p1, double *mean, double *variance)
{
    const int64_t p = p0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    double n = 0.0, mu = 0.0, m2 = 0.0;
    for (int g = 0; g < world; ++g) {
        const unsigned char *blk = gath + (int64_t)g * L.bytes;
        const double ng = (double)reinterpret_cast<const int64_t *>(blk)[0];
        if (ng <= 0.0) continue;
        const double mg = reinterpret_cast<const double *>(blk + L.off_mean)[p], vg = reinterpret_cast<const double *>(blk + L.off_var)[p];
        const double d = mg - mu, nn = n + ng;
        m2 += vg * ng + d * d * (n * ng / nn);
        mu += d * (ng / nn);
        n = nn;
    }
    mean[p] = mu;
    variance[p] = m2 / n;
}

// rows of the [rows][t] kept arrays of a compact problem -> rows `ids` of the full arrays
__global__ void k_ts_scatter_kept(const int32_t *ids, int64_t t, const uint64_t *kk, const int32_t *ki, const double *vl, const int32_t *vt,
                                  uint64_t *okk, int32_t *oki, double *ovl, int32_t *ovt)
{
    const int64_t i = blockIdx.x, r = ids[i];
    if (threadIdx.x == 0) { ovl[r] = vl[i]; ovt[r] = vt[i]; }
    for (int64_t j = threadIdx.x; j < t; j += blockDim.x) { okk[r * t + j] = kk[i * t + j]; oki[r * t + j] = ki[i * t + j]; }
}

// local top-`m` (keys + local trial indices) of `rows` exposure rows at E, by the guaranteed schedule -> block.
// kk / ki: two scratch arrays of rows * min(m, N_local) entries each (the selection's ping-pong).
static int ts_local_block(mcp_ctx *ctx, const double *E, int64_t rows, int64_t m, bool with_moments, uint64_t *const kk[2], int32_t *const ki[2],
                          mcp::DevBuf &blockbuf, TsLayout &L)
{
    const int64_t m_have = m < ctx->N ? m : ctx->N;
    MCP_CUDA(ctx, ctx->dTau.reserve((size_t)rows * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dFbOut.reserve((size_t)rows * 16));
    {
        int64_t cap = 4 * m_have + 264;
        if (cap < 2304) cap = 2304;
        if (cap > 16384) cap = 16384;
        ctx->selCap = (int)cap;
        MCP_CUDA(ctx, sel_configure((int)sel_smem_bytes(16384, kSelMaxSeg)));
    }
    StreamIO io{};
    io.E = E; io.P = rows;
    io.var_loss = ctx->dFbOut.as<double>(); io.var_trial = reinterpret_cast<int32_t *>(ctx->dFbOut.as<double>() + rows); // scratch: local order statistics
    io.kkey[0] = kk[0]; io.kkey[1] = kk[1];
    io.kidx[0] = ki[0]; io.kidx[1] = ki[1];
    io.tkey = ctx->dTau.as<uint64_t>();
    io.keep_only = true;
    // first round: the two-step optimistic schedule (section 4.0) for the local top-m as well; rows where the bet fails
    // are redone by the guaranteed schedule and their kept lists scattered back
    const int64_t n0 = first_step_len(ctx->N, m_have, true);
    const bool optimistic = with_moments && ctx->optimistic && ctx->N >= 2 * n0 && optimistic_rank(ctx->N, n0, m_have) < m_have &&
                            2 * optimistic_rank(ctx->N, n0, m_have) <= n0;
    if (optimistic) {
        MCP_CUDA(ctx, ctx->dFail.reserve((size_t)(rows + 4) * sizeof(int32_t)));
        MCP_CUDA(ctx, cudaMemsetAsync(ctx->dFail.p, 0, sizeof(int), ctx->stream));
        io.fail_cnt = ctx->dFail.as<int>();
        io.fail_rows = ctx->dFail.as<int32_t>() + 4;
    }
    MCP_TRY(risk_stream_core(ctx, io, m_have, optimistic, nullptr));
    if (optimistic) {
        MCP_CUDA(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dFail.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const int64_t nf = *reinterpret_cast<int *>(ctx->hPinned);
        if (nf > 0) {
            const int F = ctx->F;
            MCP_CUDA(ctx, ctx->dFbE.reserve((size_t)nf * F * sizeof(double)));
            MCP_CUDA(ctx, ctx->dFbKey.reserve((size_t)nf * m_have * 2 * sizeof(uint64_t)));
            MCP_CUDA(ctx, ctx->dFbIdx.reserve((size_t)nf * m_have * 2 * sizeof(int32_t)));
            MCP_CUDA(ctx, ctx->dFinKey.reserve((size_t)nf * 16));
            {
                KScope ks(ctx, MCP_K_SELECT);
                k_gather_rows<<<(unsigned)std::min<int64_t>((nf * F + 255) / 256, 4096), 256, 0, ctx->stream>>>(E, io.fail_rows, nf, F, ctx->dFbE.as<double>());
                MCP_TRY(check_launch(ctx, "k_gather_rows"));
            }
            StreamIO fb{};
            fb.E = ctx->dFbE.as<double>(); fb.P = nf;
            fb.var_loss = ctx->dFinKey.as<double>(); fb.var_trial = reinterpret_cast<int32_t *>(ctx->dFinKey.as<double>() + nf);
            fb.kkey[0] = ctx->dFbKey.as<uint64_t>(); fb.kkey[1] = fb.kkey[0] + nf * m_have;
            fb.kidx[0] = ctx->dFbIdx.as<int32_t>(); fb.kidx[1] = fb.kidx[0] + nf * m_have;
            mcp::DevBuf &tk = ctx->dTsFlags;       // thresholds of the compact problem
            MCP_CUDA(ctx, tk.reserve((size_t)nf * sizeof(uint64_t) + 16));
            fb.tkey = tk.as<uint64_t>();
            fb.keep_only = true;
            MCP_TRY(risk_stream_core(ctx, fb, m_have, false, nullptr));
            KScope ks(ctx, MCP_K_SELECT);
            k_ts_scatter_kept<<<(unsigned)nf, 128, 0, ctx->stream>>>(io.fail_rows, m_have, fb.kkey[fb.final_buf], fb.kidx[fb.final_buf], fb.var_loss, fb.var_trial,
                                                                    io.kkey[io.final_buf], io.kidx[io.final_buf], io.var_loss, io.var_trial);
            MCP_TRY(check_launch(ctx, "k_ts_scatter_kept"));
        }
    }
    L = ts_layout(rows, m);
    MCP_CUDA(ctx, blockbuf.reserve((size_t)L.bytes));
    KScope ks(ctx, MCP_K_SELECT);
    k_ts_pack<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(io.kkey[io.final_buf], io.kidx[io.final_buf], rows, m, m_have, ctx->tsTrialBase,
                                                                         ctx->N, with_moments ? ctx->dMean.as<double>() : nullptr,
                                                                         with_moments ? ctx->dVar.as<double>() : nullptr, blockbuf.as<unsigned char>(), L);
    return check_launch(ctx, "k_ts_pack");
}

int risk_ts_local(mcp_ctx *ctx, double alpha, int64_t n_global, int64_t trial_base, int world, void **d_block, int64_t *block_bytes)
{
    const int64_t P = ctx->P, N = ctx->N;
    if (P <= 0 || N <= 0) return fail(ctx, MCP_E_STATE, "trial-sharded run: empty problem");
    if (world < 1 || n_global < N || trial_base < 0 || trial_base + N > n_global || n_global > 0x7fffffff)
        return fail(ctx, MCP_E_ARG, "trial-sharded run: bad trial range");
    if (N > (n_global + world - 1) / world) return fail(ctx, MCP_E_ARG, "trial-sharded run: shards must be balanced (N_local <= ceil(N_global / world))");
    if (!mma_supported(ctx->F)) return fail(ctx, MCP_E_UNSUPPORTED, "trial-sharded run: factor count must be 16, 32, 48 or 64");
    ctx->haveRun = false;
    ctx->tsT = mcp_tail_size(n_global, alpha);
    ctx->tsM = ts_candidates_per_rank(n_global, ctx->tsT, world);
    if (ctx->tsSlackOverride >= 0) { // test hook: force a tighter m so that the second round is exercised
        int64_t m = (ctx->tsT + world - 1) / world + ctx->tsSlackOverride;
        if (m > ctx->tsT) m = ctx->tsT;
        ctx->tsM = m < 1 ? 1 : m;
    }
    ctx->tsNGlobal = n_global; ctx->tsTrialBase = trial_base; ctx->tsWorld = world; ctx->tsNFlag = 0; ctx->tsGath = nullptr;
    MCP_CUDA(ctx, ctx->dMean.reserve(P * sizeof(double)));
    MCP_CUDA(ctx, ctx->dVar.reserve(P * sizeof(double)));
    // mean / variance of this shard's trials by linearity, on the side stream under the valuation
    MCP_CUDA(ctx, cudaEventRecord(ctx->evFork, ctx->stream));
    MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
    MCP_TRY(run_moments_linear(ctx));
    MCP_CUDA(ctx, cudaEventRecord(ctx->evJoin, ctx->stream2));
    MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));
    TsLayout L;
    {
        const int64_t mh = ctx->tsM < N ? ctx->tsM : N;
        MCP_CUDA(ctx, ctx->dKeptKey.reserve((size_t)P * mh * sizeof(uint64_t)));
        MCP_CUDA(ctx, ctx->dKeptIdx.reserve((size_t)P * mh * sizeof(int32_t)));
        MCP_CUDA(ctx, ctx->dKept2Key.reserve((size_t)P * mh * sizeof(uint64_t)));
        MCP_CUDA(ctx, ctx->dKept2Idx.reserve((size_t)P * mh * sizeof(int32_t)));
        uint64_t *const kk[2] = {ctx->dKeptKey.as<uint64_t>(), ctx->dKept2Key.as<uint64_t>()};
        int32_t *const ki[2] = {ctx->dKeptIdx.as<int32_t>(), ctx->dKept2Idx.as<int32_t>()};
        MCP_TRY(ts_local_block(ctx, ctx->dE.as<double>(), P, ctx->tsM, true, kk, ki, ctx->dTsSend, L));
    }
    if (d_block) *d_block = ctx->dTsSend.p;
    if (block_bytes) *block_bytes = L.bytes;
    return MCP_OK;
}

// merge `rows` rows of `world` gathered blocks (layout L): top-t into kkey/kidx[rows][t], thresholds into vl/vt
static int ts_merge_rows(mcp_ctx *ctx, const unsigned char *gath, int world, const TsLayout &L, int64_t row0, int64_t rows, int64_t t,
                         uint64_t *kkey, int32_t *kidx, double *vl, int32_t *vt)
{
    if (rows <= 0) return MCP_OK;
    SelArgs a{};
    a.kept_row0 = row0;
    a.kept_key_in = reinterpret_cast<const uint64_t *>(gath + L.off_keys);
    a.kept_idx_in = reinterpret_cast<const int32_t *>(gath + L.off_idx);
    a.kept_n = (int64_t)world * L.m;
    a.kept_segs = world;
    a.kept_seg_bytes = L.bytes;
    a.kept_key_out = kkey; a.kept_idx_out = kidx;
    a.tail = t;
    a.var_loss = vl; a.var_trial = vt;
    int64_t cap = a.kept_n + 8;
    if (cap < 2304) cap = 2304;
    if (cap > 16384) cap = 16384;
    a.cap = (int)cap;
    MCP_CUDA(ctx, sel_configure((int)sel_smem_bytes(16384, kSelMaxSeg)));
    KScope ks(ctx, MCP_K_SELECT);
    if (sel_smem_bytes(a.cap, 0) > 110 * 1024) k_select_t<1024><<<(unsigned)rows, 1024, sel_smem_bytes(a.cap, 0), ctx->stream>>>(a); // one CTA per SM: make it wide
    else k_select_t<256><<<(unsigned)rows, 256, sel_smem_bytes(a.cap, 0), ctx->stream>>>(a);
    return check_launch(ctx, "k_select (merge)");
}

static void ts_slice(int64_t P, int rank, int world, int64_t &p0, int64_t &p1)
{
    const int64_t base = P / world, extra = P % world;
    p0 = rank * base + (rank < extra ? rank : extra);
    p1 = p0 + base + (rank < extra ? 1 : 0);
}

// What the merge below (and the pooled moments) read of the ranks' first-round blocks: the header, and the keys / trial
// indices / mean / variance of the reader's own portfolio slice -- an eighth of each block on eight GPUs.
int risk_ts_exchange_regions(mcp_ctx *ctx, int world, ExchRegion regions[5], int64_t *p0)
{
    if (ctx->tsWorld != world || world < 1) return fail(ctx, MCP_E_STATE, "ts exchange: call ts_local first (same world)");
    const TsLayout L = ts_layout(ctx->P, ctx->tsM);
    regions[0] = ExchRegion{0, 0, 64};
    regions[1] = ExchRegion{(size_t)L.off_keys, (size_t)L.m * 8, 0};
    regions[2] = ExchRegion{(size_t)L.off_mean, 8, 0};
    regions[3] = ExchRegion{(size_t)L.off_var, 8, 0};
    regions[4] = ExchRegion{(size_t)L.off_idx, (size_t)L.m * 4, 0};
    for (int g = 0; g < world; ++g) {
        int64_t a, b;
        ts_slice(ctx->P, g, world, a, b);
        p0[g] = a; p0[g + 1] = b;
    }
    return MCP_OK;
}

// Merge THIS RANK'S portfolio slice from the gathered blocks and test it; the slice's flags (one byte per row, padded to
// ceil(P / world) bytes so that every rank's block has the same size) are what the ranks exchange next.
int risk_ts_merge(mcp_ctx *ctx, const void *d_gathered, int rank, void **d_flags, int64_t *flag_bytes)
{
    if (ctx->tsWorld < 1 || !d_gathered) return fail(ctx, MCP_E_STATE, "ts_merge: call ts_local first");
    if (rank < 0 || rank >= ctx->tsWorld) return fail(ctx, MCP_E_ARG, "ts_merge: bad rank");
    const int64_t P = ctx->P, t = ctx->tsT;
    const TsLayout L = ts_layout(P, ctx->tsM);
    const unsigned char *g = reinterpret_cast<const unsigned char *>(d_gathered);
    ctx->tsGath = g;
    ctx->tsRank = rank;
    int64_t p0, p1;
    ts_slice(P, rank, ctx->tsWorld, p0, p1);
    const int64_t fb = (P + ctx->tsWorld - 1) / ctx->tsWorld;
    MCP_CUDA(ctx, ctx->dVarLoss.reserve(P * sizeof(double)));
    MCP_CUDA(ctx, ctx->dVarTrial.reserve(P * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dKeptKey.reserve((size_t)P * t * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dKeptIdx.reserve((size_t)P * t * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dFail.reserve((size_t)(P + 4) * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dTsFlags.reserve((size_t)fb + 16));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dTsFlags.p, 0, (size_t)fb + 16, ctx->stream));
    MCP_TRY(ts_merge_rows(ctx, g, ctx->tsWorld, L, p0, p1 - p0, t, ctx->dKeptKey.as<uint64_t>() + p0 * t, ctx->dKeptIdx.as<int32_t>() + p0 * t,
                          ctx->dVarLoss.as<double>() + p0, ctx->dVarTrial.as<int32_t>() + p0));
    MCP_CUDA(ctx, cudaMemsetAsync(ctx->dFail.p, 0, sizeof(int), ctx->stream));
    if (p1 > p0) {
        KScope ks(ctx, MCP_K_SELECT);
        k_ts_check<<<(unsigned)(((p1 - p0) * 32 + 255) / 256), 256, 0, ctx->stream>>>(g, ctx->tsWorld, L, p0, p1 - p0, ctx->dVarLoss.as<double>(),
                                                                                     ctx->dVarTrial.as<int32_t>(), ctx->dTsFlags.as<unsigned char>(), ctx->dFail.as<int>());
        MCP_TRY(check_launch(ctx, "k_ts_check"));
    }
    if (d_flags) *d_flags = ctx->dTsFlags.p;
    if (flag_bytes) *flag_bytes = fb;
    return MCP_OK;
}

// every rank's flag block (world x ceil(P / world) bytes, rank order) -> the flagged rows, ascending, identical everywhere
int risk_ts_flags(mcp_ctx *ctx, const void *d_gathered_flags, int64_t *nflag_out)
{
    if (ctx->tsWorld < 1 || !ctx->tsGath || !d_gathered_flags) return fail(ctx, MCP_E_STATE, "ts_flags: merge first");
    const int64_t P = ctx->P, fb = (P + ctx->tsWorld - 1) / ctx->tsWorld;
    std::vector<unsigned char> f((size_t)fb * ctx->tsWorld);
    MCP_CUDA(ctx, cudaMemcpyAsync(f.data(), d_gathered_flags, f.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->tsFlagRows.clear();
    for (int g = 0; g < ctx->tsWorld; ++g) {
        int64_t p0, p1;
        ts_slice(P, g, ctx->tsWorld, p0, p1);
        for (int64_t p = p0; p < p1; ++p)
            if (f[(size_t)(g * fb + (p - p0))]) ctx->tsFlagRows.push_back((int32_t)p);
    }
    ctx->tsNFlag = (int64_t)ctx->tsFlagRows.size();
    if (ctx->tsNFlag > 0) {
        MCP_CUDA(ctx, cudaMemcpyAsync(ctx->dFail.as<int32_t>() + 4, ctx->tsFlagRows.data(), ctx->tsFlagRows.size() * sizeof(int32_t),
                                      cudaMemcpyHostToDevice, ctx->stream));
        MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (nflag_out) *nflag_out = ctx->tsNFlag;
    return MCP_OK;
}

// second round, flagged rows only: the FULL local tail (min(t, N_local) candidates) of every flagged row
int risk_ts_local2(mcp_ctx *ctx, void **d_block, int64_t *block_bytes)
{
    const int64_t nf = ctx->tsNFlag;
    if (nf <= 0) return fail(ctx, MCP_E_STATE, "ts_local2: no flagged rows");
    const int F = ctx->F;
    MCP_CUDA(ctx, ctx->dFbE.reserve((size_t)nf * F * sizeof(double)));
    {
        KScope ks(ctx, MCP_K_SELECT);
        k_gather_rows<<<(unsigned)std::min<int64_t>((nf * F + 255) / 256, 4096), 256, 0, ctx->stream>>>(ctx->dE.as<double>(), ctx->dFail.as<int32_t>() + 4, nf, F,
                                                                                                      ctx->dFbE.as<double>());
        MCP_TRY(check_launch(ctx, "k_gather_rows"));
    }
    // (the first round's merged tails stay in dKeptKey / dKeptIdx; this selection ping-pongs in the fallback buffers)
    const int64_t t = ctx->tsT, mh = t < ctx->N ? t : ctx->N;
    MCP_CUDA(ctx, ctx->dFbKey.reserve((size_t)nf * mh * 2 * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dFbIdx.reserve((size_t)nf * mh * 2 * sizeof(int32_t)));
    uint64_t *const kk[2] = {ctx->dFbKey.as<uint64_t>(), ctx->dFbKey.as<uint64_t>() + nf * mh};
    int32_t *const ki[2] = {ctx->dFbIdx.as<int32_t>(), ctx->dFbIdx.as<int32_t>() + nf * mh};
    TsLayout L;
    MCP_TRY(ts_local_block(ctx, ctx->dFbE.as<double>(), nf, t, false, kk, ki, ctx->dTsSend, L));
    if (d_block) *d_block = ctx->dTsSend.p;
    if (block_bytes) *block_bytes = L.bytes;
    return MCP_OK;
}

int risk_ts_merge2(mcp_ctx *ctx, const void *d_gathered2)
{
    const int64_t nf = ctx->tsNFlag, t = ctx->tsT;
    if (nf <= 0 || !d_gathered2) return fail(ctx, MCP_E_STATE, "ts_merge2: no flagged rows");
    const TsLayout L = ts_layout(nf, t);
    MCP_CUDA(ctx, ctx->dFinKey.reserve((size_t)nf * t * sizeof(uint64_t)));
    MCP_CUDA(ctx, ctx->dFinIdx.reserve((size_t)nf * t * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dFbOut.reserve((size_t)nf * 16));
    double *vl = ctx->dFbOut.as<double>();
    int32_t *vt = reinterpret_cast<int32_t *>(vl + nf);
    MCP_TRY(ts_merge_rows(ctx, reinterpret_cast<const unsigned char *>(d_gathered2), ctx->tsWorld, L, 0, nf, t, ctx->dFinKey.as<uint64_t>(),
                          ctx->dFinIdx.as<int32_t>(), vl, vt));
    KScope ks(ctx, MCP_K_SELECT);
    k_ts_scatter_kept<<<(unsigned)nf, 128, 0, ctx->stream>>>(ctx->dFail.as<int32_t>() + 4, t, ctx->dFinKey.as<uint64_t>(), ctx->dFinIdx.as<int32_t>(), vl, vt,
                                                            ctx->dKeptKey.as<uint64_t>(), ctx->dKeptIdx.as<int32_t>(), ctx->dVarLoss.as<double>(),
                                                            ctx->dVarTrial.as<int32_t>());
    return check_launch(ctx, "k_ts_scatter_kept");
}

// this rank's portfolio slice: breach lists, CVaR, pooled moments, encoding
int risk_ts_finish(mcp_ctx *ctx, int rank, int world, int codec_id, int framing)
{
    if (!ctx->tsGath || world != ctx->tsWorld || rank != ctx->tsRank) return fail(ctx, MCP_E_STATE, "ts_finish: merge first (same rank / world)");
    const int64_t P = ctx->P, t = ctx->tsT;
    int64_t p0, p1;
    ts_slice(P, rank, world, p0, p1);
    const int64_t rows = p1 - p0;
    MCP_CUDA(ctx, ctx->dCvar.reserve(P * sizeof(double)));
    MCP_CUDA(ctx, ctx->dBreach.reserve((size_t)P * t * sizeof(int32_t)));
    MCP_CUDA(ctx, ctx->dEncOff.reserve((size_t)(rows + 1) * sizeof(int64_t)));
    int64_t total = 0;
    if (rows > 0) {
        MCP_TRY(run_finalize(ctx, rows, t, ctx->dKeptKey.as<uint64_t>() + p0 * t, ctx->dKeptIdx.as<int32_t>() + p0 * t,
                             ctx->dBreach.as<int32_t>() + p0 * t, ctx->dCvar.as<double>() + p0, ctx->tsNGlobal));
        {
            KScope ks(ctx, MCP_K_MOMENTS);
            k_ts_moments_merge<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(ctx->tsGath, world, ts_layout(P, ctx->tsM), p0, p1,
                                                                                       ctx->dMean.as<double>(), ctx->dVar.as<double>());
            MCP_TRY(check_launch(ctx, "k_ts_moments_merge"));
        }
        MCP_TRY(encode_uniform_lists(ctx, codec_id, framing, ctx->dBreach.as<int32_t>() + p0 * t, t, rows, ctx->dEncOff, ctx->dEnc, &total));
    } else
        MCP_CUDA(ctx, cudaMemsetAsync(ctx->dEncOff.p, 0, sizeof(int64_t), ctx->stream));
    MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->runP = rows; ctx->runRow0 = p0; ctx->runN = ctx->tsNGlobal; ctx->runTail = t; ctx->runEncBytes = total; ctx->runFallbackRows = ctx->tsNFlag;
    ctx->haveRun = true;
    return MCP_OK;
}

// rows x F -> rows x Fp with zero columns appended
__global__ void k_pad_columns(const double *__restrict__ in, int64_t rows, int F, int Fp, double *__restrict__ out)
{
    const int64_t total = rows * Fp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / Fp;
        const int k = (int)(i - r * Fp);
        out[i] = k < F ? in[r * F + k] : 0.0;
    }
}

static int risk_run_inner(mcp_ctx *ctx, double alpha, int codec_id, int framing);

// Factor counts without a tensor-core kernel (F < 64, not 16/32/48): run the streaming path on copies of E and S
// padded with zero columns to the next supported count.  A padded step computes fma(0, 0, acc) = acc, so every loss,
// order statistic, breach list and moment is unchanged (the only observable difference would be V = -0.0 becoming
// +0.0, and the path consumes 0.0 - V and sums, never V's sign bit; mcp_value keeps the unpadded scalar kernel).
int risk_run(mcp_ctx *ctx, double alpha, int codec_id, int framing)
{
    const int F = ctx->F;
    if (mma_supported(F) || F > 64 || ctx->forceDense) return risk_run_inner(ctx, alpha, codec_id, framing);
    const int Fp = F <= 16 ? 16 : F <= 32 ? 32 : F <= 48 ? 48 : 64;
    MCP_CUDA(ctx, ctx->dEpad.reserve((size_t)ctx->P * Fp * sizeof(double) + 16));
    MCP_CUDA(ctx, ctx->dSpad.reserve((size_t)ctx->N * Fp * sizeof(double) + 16));
    {
        KScope ks(ctx, MCP_K_GEN, 2);
        const int grid = ctx->prop.multiProcessorCount * 8;
        k_pad_columns<<<grid, 256, 0, ctx->stream>>>(ctx->dE.as<double>(), ctx->P, F, Fp, ctx->dEpad.as<double>());
        k_pad_columns<<<grid, 256, 0, ctx->stream>>>(ctx->dS.as<double>(), ctx->N, F, Fp, ctx->dSpad.as<double>());
        MCP_TRY(check_launch(ctx, "k_pad_columns"));
    }
    std::swap(ctx->dE, ctx->dEpad);
    std::swap(ctx->dS, ctx->dSpad);
    ctx->F = Fp;
    ctx->sfragValid = false;
    const int rc = risk_run_inner(ctx, alpha, codec_id, framing);
    std::swap(ctx->dE, ctx->dEpad);
    std::swap(ctx->dS, ctx->dSpad);
    ctx->F = F;
    ctx->sfragValid = false;
    return rc;
}

static int risk_run_inner(mcp_ctx *ctx, double alpha, int codec_id, int framing)
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
    {
        // shared-memory staging capacity of the selection kernel: the dense first step has max(2048, 4t) candidates
        int64_t cap = 4 * t + 264; // dense first step: max(2048, 4t) rounded up to 8; later steps: t kept + ~2t new
        if (cap < 2304) cap = 2304;
        if (cap > 16384) cap = 16384; // 200 KB; larger rows select straight from global memory
        ctx->selCap = (int)cap;
        MCP_CUDA(ctx, sel_configure((int)sel_smem_bytes(ctx->selCap, kSelMaxSeg)));
    }

    const bool stream_path = mma_supported(ctx->F) && !ctx->forceDense;
    const bool side_moments = stream_path && !ctx->fusedMoments;
    if (side_moments) {
        // mean / variance by linearity depend on E and S only: run them on the side stream, under the valuation
        MCP_CUDA(ctx, cudaEventRecord(ctx->evFork, ctx->stream));
        MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
        MCP_TRY(run_moments_linear(ctx));
        MCP_CUDA(ctx, cudaEventRecord(ctx->evJoin, ctx->stream2));
    }
    if (stream_path) MCP_TRY(risk_run_stream(ctx, t));
    else MCP_TRY(risk_run_dense(ctx, t));
    if (side_moments) MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));

    MCP_TRY(scenarios_tail_ready(ctx)); // (a schedule that never reached the tail rows cannot exist, but the event must not dangle)
    const bool early_stats = ctx->earlyMean || ctx->earlyVar || ctx->earlyVarLoss || ctx->earlyVarTrial || ctx->earlyCvar;
    auto send_stats = [&](cudaStream_t st, bool moments_too) -> int {
        if (moments_too && ctx->earlyMean) MCP_CUDA(ctx, cudaMemcpyAsync(ctx->earlyMean, ctx->dMean.p, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (moments_too && ctx->earlyVar) MCP_CUDA(ctx, cudaMemcpyAsync(ctx->earlyVar, ctx->dVar.p, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (ctx->earlyVarLoss) MCP_CUDA(ctx, cudaMemcpyAsync(ctx->earlyVarLoss, ctx->dVarLoss.p, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (ctx->earlyVarTrial) MCP_CUDA(ctx, cudaMemcpyAsync(ctx->earlyVarTrial, ctx->dVarTrial.p, (size_t)P * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (ctx->earlyCvar) MCP_CUDA(ctx, cudaMemcpyAsync(ctx->earlyCvar, ctx->dCvar.p, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, st));
        return MCP_OK;
    };
    if (ctx->earlyBreachDst || early_stats) {
        // the caller wants the statistics / the raw lists on the host: those copies start on the side stream now, under the encode
        MCP_CUDA(ctx, cudaEventRecord(ctx->evFork, ctx->stream));
        MCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
        MCP_TRY(send_stats(ctx->stream2, true));
        if (ctx->earlyBreachDst)
            MCP_CUDA(ctx, cudaMemcpyAsync(ctx->earlyBreachDst, ctx->dBreach.p, (size_t)P * t * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream2));
    }
    // ---- encode the breach lists (uniform length t) into the JavaFastPFOR format
    MCP_CUDA(ctx, ctx->dEncOff.reserve((size_t)(P + 1) * sizeof(int64_t)));
    int64_t total = 0;
    MCP_TRY(encode_uniform_lists(ctx, codec_id, framing, ctx->dBreach.as<int32_t>(), t, P, ctx->dEncOff, ctx->dEnc, &total)); // (synchronises)
    if (ctx->failPending) {
        ctx->failPending = false;
        const int64_t nf = *reinterpret_cast<int *>(reinterpret_cast<char *>(ctx->hPinned) + 64);
        ctx->runFallbackRows = nf;
        if (nf > 0) { // rare: some rows lost the optimistic bet -- redo them exactly, then encode again
            MCP_TRY(risk_stream_fallback(ctx, t, nf));
            if (ctx->earlyBreachDst || early_stats) { // they already left for the host under the first encode: send the repaired ones
                MCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
                MCP_TRY(send_stats(ctx->stream, false));
                if (ctx->earlyBreachDst)
                    MCP_CUDA(ctx, cudaMemcpyAsync(ctx->earlyBreachDst, ctx->dBreach.p, (size_t)P * t * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            }
            MCP_TRY(encode_uniform_lists(ctx, codec_id, framing, ctx->dBreach.as<int32_t>(), t, P, ctx->dEncOff, ctx->dEnc, &total));
        }
    }
    ctx->runP = P;
    ctx->runRow0 = 0;
    ctx->runN = N;
    ctx->runTail = t;
    ctx->runEncBytes = total;
    ctx->haveRun = true;
    return MCP_OK;
}

} // namespace mcp
