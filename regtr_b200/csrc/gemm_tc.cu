// fp32-accurate dense GEMM on the 5th-generation tensor cores (tcgen05, 3xTF32 split).
//
//   C[M,N] = act( A[M,K] @ B[N,K]^T + bias[N] + R[M,N] )        (all row-major fp32)
//
// It replaces the cuBLAS SIMT sgemm calls behind every nn.Linear on the RegTR hot path (unary
// blocks kpconv_blocks.py:546-561, feat_proj regtr.py:145, attention in/out projections and FFN
// transformers.py:197-238, regressor MLP regtr.py:413-443) and the KPConv weight contraction
// (kpconv_blocks.py:401-406) while keeping the fp32 parity tolerance: a single TF32 pass moves
// the pose by 4e-3 (measured), so every operand is split x = hi + lo with both halves
// exactly representable in TF32 and the product is accumulated as hi*hi + hi*lo + lo*hi in fp32
// TMEM accumulators (dropped term lo*lo ~ 2^-22 relative).
//
// Structure (one 128 x BN output tile per CTA, 6 warps):
//   warp 0     TMA producer: A tile, B_hi tile, B_lo tile per 32-wide k-block (SWIZZLE_128B)
//   warp 1     TMEM allocator + MMA issuer: 4 k-steps x 3 tcgen05.mma.kind::tf32 per stage
//   warps 2-5  operand split (A -> A_hi in place, A_lo to a second buffer) between TMA arrival
//              and MMA issue, then the epilogue: TMEM -> registers -> bias/residual/ReLU -> global
// B (weights) is split once on the host side of the ABI (regtr_split_tf32) and cached.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc.cuh"

namespace {

// Optional bf16 epilogue for the attention in-projection: columns [0, split) are written row-major
// to `qk` (the q and k halves), columns >= split transposed to `vt` (one row per channel).
struct QkvOut {
    __nv_bfloat16* qk;
    int ld_qk;
    __nv_bfloat16* vt;
    int ld_vt;
    int split;
};

// Optional InstanceNorm statistics of the OUTPUT (per cloud, per column: mean and 1/sqrt(var + eps) over the
// cloud's rows, kpconv_blocks.py:497-519), accumulated in the epilogue so that no separate pass re-reads C.
// Every value is converted to FIXED POINT (x * 2^32 and x^2 * 2^24 as int64, both exact products in fp64) and
// summed with integer adds only -- first over the 32 rows of an epilogue warp (through a transposed shared-memory
// tile, lane = column), then across warps and CTAs with 64-bit integer atomics.  Integer addition is associative
// and exact: the totals, and therefore the statistics, are bit-identical from run to run whatever the CTA order
// (no floating-point atomics), and the variance E[x^2] - mean^2 is formed in fp64 from exact sums, so a constant
// column gives exactly 1/sqrt(eps) like the reference's two-pass variance.  Range: |x| < 2^10 with up to 2^19
// rows per cloud (features of this network are O(1) - O(100)).  The last CTA of the launch (completion counter)
// turns the accumulators into (mean, rstd) and leaves accumulators and counter zero for the next launch; a row
// count travels along and poisons the statistics (NaN) if it ever disagrees with the cloud size.
struct InStats {
    const int32_t* offs;        // (n_clouds + 1) row offsets of the clouds
    int n_clouds;
    unsigned long long* acc;    // [n_clouds][N][3]: sum * 2^32, sumsq * 2^24, rows; zero on entry, left zero
    int32_t* counter;           // zero on entry, left zero
    float2* stats;              // out [n_clouds][N]
    float eps;
};

constexpr double FX_S = 4294967296.0;        // 2^32
constexpr double FX_SS = 16777216.0;         // 2^24

__device__ __forceinline__ long long fx_s(float v) { return __double2ll_rn((double)v * FX_S); }
__device__ __forceinline__ long long fx_ss(float v) { const double d = (double)v; return __double2ll_rn(d * d * FX_SS); }

// Completion protocol shared by every CTA of a statistics launch (also the capacity-padding CTAs that have no
// tile): the last CTA to arrive finalises all (cloud, column) statistics and restores the zero state.
__device__ __forceinline__ void instats_tail(const InStats& st, int N, int total_ctas) {
    __shared__ int s_last;
    __threadfence();                               // this CTA's accumulator updates are visible before it is counted
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(st.counter, 1);
        s_last = prev == total_ctas - 1;
        if (s_last) *st.counter = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < st.n_clouds * N; i += blockDim.x) {
        const int c = i / N;
        unsigned long long* a = st.acc + (size_t)i * 3;
        const double s = (double)(long long)__ldcg(a) * (1.0 / FX_S), ss = (double)(long long)__ldcg(a + 1) * (1.0 / FX_SS);
        const long long rows = (long long)__ldcg(a + 2);
        const int n = st.offs[c + 1] - st.offs[c];
        const double dn = n > 0 ? (double)n : 1.0;
        const double mean = s / dn;
        double var = ss / dn - mean * mean;          // biased variance (InstanceNorm)
        var = var > 0.0 ? var : 0.0;
        float2 o = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)st.eps)));
        if (rows != (long long)n) o = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));   // inconsistent: poison
        st.stats[i] = o;
        a[0] = 0ull; a[1] = 0ull; a[2] = 0ull;
    }
}

constexpr int BM = 128;
constexpr int BK = 32;                       // fp32 elements = 128 bytes = one swizzle span
constexpr uint32_t HI_MASK = 0xFFFFE000u;    // keep sign, exponent and 10 mantissa bits

template <int BN, int ST> struct Cfg {
    static constexpr int STAGES = ST;
    static constexpr int A_BYTES = BM * BK * 4;          // 16 KB
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    // The tensor core adds into its fp32 accumulator with truncation, so a long run of
    // accumulations drifts (measured: error grows ~linearly with K).  The k-steps of a k-block
    // are therefore spread over NACC independent TMEM accumulators that are summed with
    // round-to-nearest fp32 adds in the epilogue (4x fewer hardware accumulations each).
    static constexpr int NACC = (BN == 256 || (BN == 128 && ST == 2)) ? 2 : 4;   // co-resident CTAs share 512 columns
    static constexpr int TMEM_COLS = NACC * BN;
};

// round-to-nearest-even to TF32 precision (10 mantissa bits); unbiased, so split errors do not
// accumulate linearly along K as plain truncation does
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t u = __float_as_uint(x);
    u += 0x0FFFu + ((u >> 13) & 1u);
    return __uint_as_float(u & HI_MASK);
}

__global__ void k_split_tf32(const float* __restrict__ x, long long n, float* __restrict__ hi, float* __restrict__ lo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i], h = tf32_hi(v);
    hi[i] = h;
    lo[i] = tf32_hi(v - h);
}

template <int BN, int ST>
__global__ void __launch_bounds__(192, ST == 2 ? 2 : 1)
k_gemm_tf32x3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
              const __grid_constant__ CUtensorMap tmBlo, float* __restrict__ C, int ldc,
              const float* __restrict__ bias, const float* __restrict__ R, int ldr, int M, int N, int K,
              const int32_t* __restrict__ m_dev, int relu, int kb_per_split, size_t split_stride, QkvOut qkv) {
    using P = Cfg<BN, ST>;
    extern __shared__ unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m_dev) M = min(M, *m_dev);
    if (m0 >= M) return;                                   // capacity padding tile (uniform exit)

    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    auto stage_A = [&](int s) { return reinterpret_cast<float*>(base + s * P::STAGE_BYTES); };
    auto stage_Alo = [&](int s) { return reinterpret_cast<float*>(base + s * P::STAGE_BYTES + P::A_BYTES); };
    auto stage_Bhi = [&](int s) { return reinterpret_cast<float*>(base + s * P::STAGE_BYTES + 2 * P::A_BYTES); };
    auto stage_Blo = [&](int s) { return reinterpret_cast<float*>(base + s * P::STAGE_BYTES + 2 * P::A_BYTES + P::B_BYTES); };
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + P::STAGES * P::STAGE_BYTES);
    uint64_t* full = bars;                    // TMA bytes landed           (count 1 + tx)
    uint64_t* split = bars + P::STAGES;       // A split done               (count 128)
    uint64_t* empty = bars + 2 * P::STAGES;   // MMAs of the stage retired  (count 1, tcgen05.commit)
    uint64_t* tmem_full = bars + 3 * P::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * P::STAGES + 1);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&split[s], 128); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmBhi); tc::tma_prefetch_desc(&tmBlo);
    }
    if (warp == 1) tc::tmem_alloc<P::TMEM_COLS>(tmem_slot);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_d = *tmem_slot;
    // split-K: CTA z accumulates k-blocks [kb0, kb0 + nkb) into its own partial output plane
    const int nkb_total = (K + BK - 1) / BK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int nkb = min(kb_per_split, nkb_total - kb0);
    C += (size_t)blockIdx.z * split_stride;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % P::STAGES;
                const uint32_t ph = (kb / P::STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                tc::mbar_arrive_expect_tx(&full[s], P::A_BYTES + 2 * P::B_BYTES);
                tc::tma_load_2d(stage_A(s), &tmA, &full[s], (kb0 + kb) * BK, m0);
                tc::tma_load_2d(stage_Bhi(s), &tmBhi, &full[s], (kb0 + kb) * BK, n0);
                tc::tma_load_2d(stage_Blo(s), &tmBlo, &full[s], (kb0 + kb) * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc(tc::FMT_TF32, BM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % P::STAGES;
                const uint32_t ph = (kb / P::STAGES) & 1;
                tc::mbar_wait(&split[s], ph);
                tc::fence_after_thread_sync();
                const uint64_t dA = tc::umma_desc_sw128_kmajor(tc::smem_u32(stage_A(s)));
                const uint64_t dAlo = tc::umma_desc_sw128_kmajor(tc::smem_u32(stage_Alo(s)));
                const uint64_t dBhi = tc::umma_desc_sw128_kmajor(tc::smem_u32(stage_Bhi(s)));
                const uint64_t dBlo = tc::umma_desc_sw128_kmajor(tc::smem_u32(stage_Blo(s)));
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {             // UMMA_K = 8 tf32 = 32 bytes
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    const uint32_t acc = tmem_d + (uint32_t)((k % P::NACC) * BN);
                    const uint32_t first = (kb == 0 && k < P::NACC) ? 0u : 1u;
                    tc::umma_tf32(acc, dAlo + adv, dBhi + adv, idesc, first);   // small terms first
                    tc::umma_tf32(acc, dA + adv, dBlo + adv, idesc, 1);
                    tc::umma_tf32(acc, dA + adv, dBhi + adv, idesc, 1);
                }
                tc::umma_commit(&empty[s]);
            }
            tc::umma_commit(tmem_full);
        }
    } else {
        const int t = threadIdx.x - 64;                        // 0..127
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % P::STAGES;
            const uint32_t ph = (kb / P::STAGES) & 1;
            tc::mbar_wait(&full[s], ph);
            float4* a = reinterpret_cast<float4*>(stage_A(s));
            float4* al = reinterpret_cast<float4*>(stage_Alo(s));
#pragma unroll
            for (int i = 0; i < P::A_BYTES / 16 / 128; ++i) {  // 8 float4 per thread; swizzle-agnostic
                const float4 v = a[t + i * 128];
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                a[t + i * 128] = h;
                al[t + i * 128] = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&split[s]);
        }
        // ---- epilogue: this warp owns TMEM lanes 32*(warp%4) .. +31 = output rows
        tc::mbar_wait(tmem_full, 0);
        tc::fence_after_thread_sync();
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < M;
        float* crow = C + (size_t)row * ldc;
        const float* rrow = R ? R + (size_t)row * ldr : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tc::tmem_ld_32x32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int a = 1; a < P::NACC; ++a) {
                float u[32];
                tc::tmem_ld_32x32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0), u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
            const int col0 = n0 + c0;
            if (!row_ok || col0 >= N) continue;
            if (qkv.qk) {                                      // bf16 epilogue (N % 32 == 0 guaranteed by the host)
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += bias[col0 + j];
                }
                if (col0 < qkv.split) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        pk[j] = *reinterpret_cast<const uint32_t*>(&b);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(qkv.qk + (size_t)row * qkv.ld_qk + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        qkv.vt[(size_t)(col0 - qkv.split + j) * qkv.ld_vt + row] = __float2bfloat16_rn(v[j]);
                }
                continue;
            }
            if (col0 + 32 <= N && (ldc & 3) == 0 && (!R || (ldr & 3) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + col0 + j); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
                    if (rrow) { const float4 r = *reinterpret_cast<const float4*>(rrow + col0 + j); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(crow + col0 + j) = o;
                }
            } else {
                for (int j = 0; j < 32 && col0 + j < N; ++j) {
                    float o = v[j];
                    if (bias) o += bias[col0 + j];
                    if (rrow) o += rrow[col0 + j];
                    if (relu) o = fmaxf(o, 0.f);
                    crow[col0 + j] = o;
                }
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<P::TMEM_COLS>(tmem_d);
}

// ---- A operand through TENSOR MEMORY ---------------------------------------------------------------------
// The shared-memory variant above pays for the 3xTF32 split of A three times in shared-memory bandwidth: the
// converter warps read the TMA tile and write hi + lo back (48 KB per k-block) and the tensor core then reads
// A_hi twice and A_lo once (another 48 KB), independent of the tile width -- at N = 32/64 that, not the MMA,
// is the k-block time.  Here the converter warps read the TMA tile once (thread = tile row, the TMEM lane it
// owns), split in registers and store hi | lo with tcgen05.st into a double-buffered 2 x (32 + 32)-column
// TMEM region; the MMAs take A from TMEM ([a_tmem] operand form) and only B from shared memory.  Shared
// memory traffic per k-block drops from 96 KB + 3 B to 32 KB + 3 B, and the issuer never waits on a
// shared-memory round trip.  lo = x - hi is exact in fp32 and is rounded to nearest TF32 as well (the tensor
// core would otherwise truncate it: a one-sided 2^-21 bias that adds up linearly along K and moved the
// ill-conditioned random-weight pose by 1e-4).
template <int BN, int NACC, int ST> struct CfgT {
    static constexpr int STAGES = ST;
    static constexpr int A_BYTES = BM * BK * 4;          // 16 KB (fp32 tile as loaded by TMA)
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int ACC_COLS = NACC * BN;           // interleaved accumulators (see Cfg)
    static constexpr int A_COLS = 2 * 2 * BK;            // 2 buffers x (hi | lo) x 32 columns
    static constexpr int NEED = ACC_COLS + A_COLS;
    static constexpr int TMEM_COLS = NEED <= 128 ? 128 : (NEED <= 256 ? 256 : 512);
    static constexpr int MIN_CTAS = (TMEM_COLS <= 256 && SMEM <= 113 * 1024) ? 2 : 1;
};

template <int BN, int NACC, int ST>
__global__ void __launch_bounds__(192, CfgT<BN, NACC, ST>::MIN_CTAS)
k_gemm_tf32x3_ts(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                 const __grid_constant__ CUtensorMap tmBlo, float* __restrict__ C, int ldc,
                 const float* __restrict__ bias, const float* __restrict__ R, int ldr, int M, int N, int K,
                 const int32_t* __restrict__ m_dev, int relu, int kb_per_split, size_t split_stride, QkvOut qkv,
                 InStats ist) {
    using P = CfgT<BN, NACC, ST>;
    extern __shared__ unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m_dev) M = min(M, *m_dev);
    if (m0 >= M) {                                         // capacity padding tile (uniform exit)
        if (ist.acc) instats_tail(ist, N, gridDim.x * gridDim.y);
        return;
    }

    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    auto stage_A = [&](int s) { return base + s * P::STAGE_BYTES; };
    auto stage_Bhi = [&](int s) { return base + s * P::STAGE_BYTES + P::A_BYTES; };
    auto stage_Blo = [&](int s) { return base + s * P::STAGE_BYTES + P::A_BYTES + P::B_BYTES; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + P::STAGES * P::STAGE_BYTES);
    uint64_t* full = bars;                    // TMA bytes landed                      (count 1 + tx)
    uint64_t* empty = bars + P::STAGES;       // MMAs of the stage retired             (count 1, tcgen05.commit)
    uint64_t* aready = bars + 2 * P::STAGES;  // A buffer b holds the split tile       (count 4: one per converter warp)
    uint64_t* afree = aready + 2;             // MMAs reading A buffer b retired       (count 1, tcgen05.commit)
    uint64_t* tmem_full = afree + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&aready[b], 4); tc::mbar_init(&afree[b], 1); }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmBhi); tc::tma_prefetch_desc(&tmBlo);
    }
    if (warp == 1) tc::tmem_alloc<P::TMEM_COLS>(tmem_slot);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t tmem_a = tmem_d + (uint32_t)P::ACC_COLS;
    const int nkb_total = (K + BK - 1) / BK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int nkb = min(kb_per_split, nkb_total - kb0);
    C += (size_t)blockIdx.z * split_stride;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % P::STAGES;
                const uint32_t ph = (kb / P::STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                tc::mbar_arrive_expect_tx(&full[s], P::A_BYTES + 2 * P::B_BYTES);
                tc::tma_load_2d(stage_A(s), &tmA, &full[s], (kb0 + kb) * BK, m0);
                tc::tma_load_2d(stage_Bhi(s), &tmBhi, &full[s], (kb0 + kb) * BK, n0);
                tc::tma_load_2d(stage_Blo(s), &tmBlo, &full[s], (kb0 + kb) * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc(tc::FMT_TF32, BM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % P::STAGES, ab = kb & 1;
                const uint32_t ph = (kb / P::STAGES) & 1, aph = (kb >> 1) & 1;
                tc::mbar_wait(&full[s], ph);               // B tiles of the stage (this thread's own acquire)
                tc::mbar_wait(&aready[ab], aph);           // A split into TMEM buffer ab
                tc::fence_after_thread_sync();
                const uint64_t dBhi = tc::umma_desc_sw128_kmajor(tc::smem_u32(stage_Bhi(s)));
                const uint64_t dBlo = tc::umma_desc_sw128_kmajor(tc::smem_u32(stage_Blo(s)));
                const uint32_t a_hi = tmem_a + (uint32_t)(ab * 2 * BK), a_lo = a_hi + BK;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {             // UMMA_K = 8 tf32
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    const uint32_t acc = tmem_d + (uint32_t)((k % NACC) * BN);
                    const uint32_t first = (kb == 0 && k < NACC) ? 0u : 1u;
                    tc::umma_tf32_ts(acc, a_lo + 8 * k, dBhi + adv, idesc, first);   // small terms first
                    tc::umma_tf32_ts(acc, a_hi + 8 * k, dBlo + adv, idesc, 1);
                    tc::umma_tf32_ts(acc, a_hi + 8 * k, dBhi + adv, idesc, 1);
                }
                tc::umma_commit(&empty[s]);                // smem stage reusable
                tc::umma_commit(&afree[ab]);               // TMEM A buffer reusable
            }
            tc::umma_commit(tmem_full);
        }
    } else {
        // ---- converter warps: thread = tile row = TMEM lane 32 * (warp % 4) + lane
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % P::STAGES, ab = kb & 1;
            const uint32_t ph = (kb / P::STAGES) & 1, aph = (kb >> 1) & 1;
            tc::mbar_wait(&full[s], ph);
            tc::mbar_wait(&afree[ab], aph ^ 1);            // the MMAs of this buffer's previous use have retired
            tc::fence_after_thread_sync();
            // SWIZZLE_128B tile: 16-byte chunk c of row r sits at chunk (c ^ (r & 7)) of its 128-byte row
            const unsigned char* rowp = stage_A(s) + r * 128;
            const uint32_t dst = tmem_a + lane_sel + (uint32_t)(ab * 2 * BK);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(rowp + (((half * 4 + c) ^ (r & 7)) << 4));
                    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t h = (__float_as_uint(f[e]) + 0x1000u) & HI_MASK;   // RN (ties away) to TF32
                        hi[4 * c + e] = h;
                        lo[4 * c + e] = (__float_as_uint(f[e] - __uint_as_float(h)) + 0x1000u) & HI_MASK;   // RN: unbiased
                    }
                }
                tc::tmem_st_32x16(dst + 16 * half, hi);
                tc::tmem_st_32x16(dst + BK + 16 * half, lo);
            }
            tc::tmem_st_wait();
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&aready[ab]);
        }
        // ---- epilogue: this warp owns TMEM lanes 32*(warp%4) .. +31 = output rows
        tc::mbar_wait(tmem_full, 0);
        tc::fence_after_thread_sync();
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < M;
        float* crow = C + (size_t)row * ldc;
        const float* rrow = (R && row_ok) ? R + (size_t)row * ldr : nullptr;
        // statistics: clouds covered by this warp's 32 rows (one, except at a cloud boundary / the last rows)
        int my_cloud = 0, c_first = 0, c_last = -1;
        if (ist.acc) {
            my_cloud = row_ok ? regtr_cloud_of(ist.offs, ist.n_clouds, row) : -1;
            c_first = __shfl_sync(0xffffffffu, my_cloud, 0);                 // lane 0's row is < M (m0 < M, q*32 may not be)
            const int last_row = min(m0 + q * 32 + 31, M - 1);
            c_last = last_row >= m0 + q * 32 ? regtr_cloud_of(ist.offs, ist.n_clouds, last_row) : -1;
            if (c_first < 0) c_last = -1;                                    // the whole warp is beyond M
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tc::tmem_ld_32x32(tmem_d + lane_sel + (uint32_t)c0, v);
#pragma unroll
            for (int a = 1; a < NACC; ++a) {
                float u[32];
                tc::tmem_ld_32x32(tmem_d + lane_sel + (uint32_t)(a * BN + c0), u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
            const int col0 = n0 + c0;
            if (col0 >= N) continue;                           // warp-uniform
            if (qkv.qk) {                                      // bf16 epilogue (N % 32 == 0 guaranteed by the host)
                if (!row_ok) continue;
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += bias[col0 + j];
                }
                if (col0 < qkv.split) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        pk[j] = *reinterpret_cast<const uint32_t*>(&b);
                    }
                    uint4* dstq = reinterpret_cast<uint4*>(qkv.qk + (size_t)row * qkv.ld_qk + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dstq[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        qkv.vt[(size_t)(col0 - qkv.split + j) * qkv.ld_vt + row] = __float2bfloat16_rn(v[j]);
                }
                continue;
            }
            if (col0 + 32 <= N && (ldc & 3) == 0 && (!R || (ldr & 3) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + col0 + j); v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w; }
                    if (rrow) { const float4 rr = *reinterpret_cast<const float4*>(rrow + col0 + j); v[j] += rr.x; v[j + 1] += rr.y; v[j + 2] += rr.z; v[j + 3] += rr.w; }
                    if (relu) { v[j] = fmaxf(v[j], 0.f); v[j + 1] = fmaxf(v[j + 1], 0.f); v[j + 2] = fmaxf(v[j + 2], 0.f); v[j + 3] = fmaxf(v[j + 3], 0.f); }
                    if (row_ok) *reinterpret_cast<float4*>(crow + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (col0 + j < N) {
                        if (bias) v[j] += bias[col0 + j];
                        if (rrow) v[j] += rrow[col0 + j];
                        if (relu) v[j] = fmaxf(v[j], 0.f);
                        if (row_ok) crow[col0 + j] = v[j];
                    }
                }
            }
            if (ist.acc && c_last >= c_first) {                // host guarantees N % 32 == 0 in this mode
                // transpose through shared memory (the pipeline stages are idle now): lane = column afterwards
                float* tile = reinterpret_cast<float*>(base) + q * (32 * 33);
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = v[j];
                __syncwarp();
                const bool uniform = (c_first == c_last) && (m0 + q * 32 + 31 < M);   // 32 rows of one cloud
                for (int c = c_first; c <= c_last; ++c) {      // one iteration except at a cloud boundary
                    long long s = 0, ss = 0, rows = 0;
                    if (uniform) {
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) { const float xv = tile[r * 33 + lane]; s += fx_s(xv); ss += fx_ss(xv); }
                        rows = 32;
                    } else {
                        for (int r = 0; r < 32; ++r) {
                            const int rc = __shfl_sync(0xffffffffu, my_cloud, r);
                            if (rc == c) { const float xv = tile[r * 33 + lane]; s += fx_s(xv); ss += fx_ss(xv); ++rows; }
                        }
                    }
                    if (rows) {
                        unsigned long long* dst = ist.acc + ((size_t)c * N + col0 + lane) * 3;
                        atomicAdd(dst, (unsigned long long)s);
                        atomicAdd(dst + 1, (unsigned long long)ss);
                        atomicAdd(dst + 2, (unsigned long long)rows);
                    }
                }
                __syncwarp();                                  // tile is rewritten by the next chunk
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<P::TMEM_COLS>(tmem_d);
    if (ist.acc) instats_tail(ist, N, gridDim.x * gridDim.y);
}

// C = act(sum_z P[z] + bias + R): deterministic split-K reduction (fixed order), 4 columns / thread
__global__ void k_splitk_reduce(const float* __restrict__ P, int splits, size_t split_stride, float* __restrict__ C,
                                int ldc, const float* __restrict__ bias, const float* __restrict__ R, int ldr, int M,
                                int N, const int32_t* __restrict__ m_dev, int relu) {
    if (m_dev) M = min(M, *m_dev);
    const int n4 = N >> 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)M * n4) return;
    const int r = (int)(t / n4), c = (int)(t % n4) * 4;
    float4 acc = *reinterpret_cast<const float4*>(P + (size_t)r * N + c);
    for (int z = 1; z < splits; ++z) {
        const float4 v = *reinterpret_cast<const float4*>(P + (size_t)z * split_stride + (size_t)r * N + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (bias) { acc.x += bias[c]; acc.y += bias[c + 1]; acc.z += bias[c + 2]; acc.w += bias[c + 3]; }
    if (R) {
        const float* rr = R + (size_t)r * ldr + c;
        acc.x += rr[0]; acc.y += rr[1]; acc.z += rr[2]; acc.w += rr[3];
    }
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    float* o = C + (size_t)r * ldc + c;
    o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
}

// Split-K reduction that also accumulates the InstanceNorm statistics of C (see InStats): warp = RS_ROWS consecutive
// rows x 128 columns, lane = 4 columns; the per-lane fixed-point column sums are flushed to the accumulators
// whenever the cloud changes and at the end.  Used for the small-M long-K contractions of the coarse levels.
constexpr int RS_ROWS = 8;      // rows per warp of k_splitk_reduce_stats (small M: parallelism over the rows matters)

__global__ void __launch_bounds__(128)
k_splitk_reduce_stats(const float* __restrict__ P, int splits, size_t split_stride, float* __restrict__ C, int ldc,
                      const float* __restrict__ bias, const float* __restrict__ R, int ldr, int M, int N,
                      const int32_t* __restrict__ m_dev, int relu, InStats ist) {
    if (m_dev) M = min(M, *m_dev);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (blockIdx.x * 4 + warp) * RS_ROWS, r1 = min(r0 + RS_ROWS, M);
    const int c = blockIdx.y * 128 + 4 * lane;
    const bool col_ok = c < N;
    long long s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0}, rows = 0;
    int cur = -1;
    auto flush = [&]() {
        if (cur >= 0 && col_ok && rows) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned long long* dst = ist.acc + ((size_t)cur * N + c + j) * 3;
                atomicAdd(dst, (unsigned long long)s[j]);
                atomicAdd(dst + 1, (unsigned long long)ss[j]);
                atomicAdd(dst + 2, (unsigned long long)rows);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[j] = 0; ss[j] = 0; }
        rows = 0;
    };
    int next_start = 0;                            // first row of the cloud after `cur`
    if (r0 < r1) {
        cur = regtr_cloud_of(ist.offs, ist.n_clouds, r0);       // one search per warp; then walk the boundaries
        next_start = ist.offs[cur + 1];
    }
    float4 vals[RS_ROWS];                          // all rows' loads in flight before the sequential part
#pragma unroll
    for (int u = 0; u < RS_ROWS; ++u) {
        const int r = r0 + u;
        vals[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < r1 && col_ok) {
            float4 acc = *reinterpret_cast<const float4*>(P + (size_t)r * N + c);
            for (int z = 1; z < splits; ++z) {
                const float4 v = *reinterpret_cast<const float4*>(P + (size_t)z * split_stride + (size_t)r * N + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            if (bias) { acc.x += bias[c]; acc.y += bias[c + 1]; acc.z += bias[c + 2]; acc.w += bias[c + 3]; }
            if (R) {
                const float* rr = R + (size_t)r * ldr + c;
                acc.x += rr[0]; acc.y += rr[1]; acc.z += rr[2]; acc.w += rr[3];
            }
            if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
            float* o = C + (size_t)r * ldc + c;
            o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
            vals[u] = acc;
        }
    }
#pragma unroll
    for (int u = 0; u < RS_ROWS; ++u) {
        const int r = r0 + u;
        if (r < r1) {
            if (r >= next_start) {                 // crossed into a later cloud (skipping empty ones)
                flush();
                do { ++cur; next_start = ist.offs[cur + 1]; } while (r >= next_start);
            }
            const float4 acc = vals[u];
            s[0] += fx_s(acc.x); s[1] += fx_s(acc.y); s[2] += fx_s(acc.z); s[3] += fx_s(acc.w);
            ss[0] += fx_ss(acc.x); ss[1] += fx_ss(acc.y); ss[2] += fx_ss(acc.z); ss[3] += fx_ss(acc.w);
            ++rows;
        }
    }
    flush();
    instats_tail(ist, N, gridDim.x * gridDim.y);
}

// split count: only for skinny problems (few output tiles) with a long K
int choose_splits(int M, int N, int K, int bn) {
    const int tiles = regtr_cdiv(M, BM) * regtr_cdiv(N, bn);
    const int nkb = regtr_cdiv(K, BK);
    if (tiles >= 74 || nkb < 16 || (N & 3)) return 1;
    int s = regtr_cdiv(148, tiles);
    if (s > nkb / 8) s = nkb / 8;
    if (s > 8) s = 8;
    return s < 1 ? 1 : s;
}

// Widest tile that covers N: the pipeline keeps several independent pairs in flight, so the machine is
// filled by OTHER forwards and what counts is the total CTA time (A is streamed / split once per
// N-tile): measured 632 -> 696 pairs/s against a "fill 148 SMs per launch" heuristic.
int choose_bn(int M, int N) {
    (void)M;
    return N > 64 ? 128 : (N > 32 ? 64 : 32);
}

// ---- host: tensor maps (driver entry point resolved through the runtime, no libcuda link)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// row-major fp32 matrix [rows, cols] with leading dimension ld; box = [box_rows, 32 cols], 128B swizzle
bool make_map(CUtensorMap* m, const float* ptr, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int ST>
int launch_gemm(const float* A, int lda, const float* Bhi, const float* Blo, int ldb, float* C, int ldc,
                const float* bias, const float* R, int ldr, int M, int N, int K, const int32_t* m_dev, int relu,
                int splits, float* ws, cudaStream_t st, QkvOut qkv = QkvOut{nullptr, 0, nullptr, 0, 0}) {
    CUtensorMap tA, tBh, tBl;
    if (!make_map(&tA, A, M, K, lda, BM) || !make_map(&tBh, Bhi, N, K, ldb, BN) || !make_map(&tBl, Blo, N, K, ldb, BN))
        return REGTR_ERR_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tf32x3<BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg<BN, ST>::SMEM);
        if (e != cudaSuccess) return -(1000 + (int)e);
        attr_set = true;
    }
    const int nkb = regtr_cdiv(K, BK);
    if (splits <= 1) {
        dim3 grid(regtr_cdiv(M, BM), regtr_cdiv(N, BN), 1);
        k_gemm_tf32x3<BN, ST><<<grid, 192, Cfg<BN, ST>::SMEM, st>>>(tA, tBh, tBl, C, ldc, bias, R, ldr, M, N, K, m_dev,
                                                                    relu, nkb, 0, qkv);
        REGTR_CHECK_LAUNCH();
        return REGTR_OK;
    }
    const int per = regtr_cdiv(nkb, splits);
    const int z = regtr_cdiv(nkb, per);                     // every plane gets >= 1 k-block
    const size_t stride = (size_t)M * N;
    dim3 grid(regtr_cdiv(M, BM), regtr_cdiv(N, BN), z);
    k_gemm_tf32x3<BN, ST><<<grid, 192, Cfg<BN, ST>::SMEM, st>>>(tA, tBh, tBl, ws, N, nullptr, nullptr, 0, M, N, K, m_dev, 0,
                                                                per, stride, QkvOut{nullptr, 0, nullptr, 0, 0});
    REGTR_CHECK_LAUNCH();
    k_splitk_reduce<<<regtr_cdiv((long long)M * (N / 4), 256), 256, 0, st>>>(ws, z, stride, C, ldc, bias, R, ldr, M, N,
                                                                            m_dev, relu);
    REGTR_CHECK_LAUNCH();
    return REGTR_OK;
}

template <int BN, int NACC, int ST>
int launch_gemm_ts(const float* A, int lda, const float* Bhi, const float* Blo, int ldb, float* C, int ldc,
                   const float* bias, const float* R, int ldr, int M, int N, int K, const int32_t* m_dev, int relu,
                   int splits, float* ws, cudaStream_t st, QkvOut qkv = QkvOut{nullptr, 0, nullptr, 0, 0},
                   InStats ist = InStats{nullptr, 0, nullptr, nullptr, nullptr, 0.f}) {
    using P = CfgT<BN, NACC, ST>;
    CUtensorMap tA, tBh, tBl;
    if (!make_map(&tA, A, M, K, lda, BM) || !make_map(&tBh, Bhi, N, K, ldb, BN) || !make_map(&tBl, Blo, N, K, ldb, BN))
        return REGTR_ERR_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tf32x3_ts<BN, NACC, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM);
        if (e != cudaSuccess) return -(1000 + (int)e);
        attr_set = true;
    }
    const int nkb = regtr_cdiv(K, BK);
    if (splits <= 1) {
        dim3 grid(regtr_cdiv(M, BM), regtr_cdiv(N, BN), 1);
        k_gemm_tf32x3_ts<BN, NACC, ST><<<grid, 192, P::SMEM, st>>>(tA, tBh, tBl, C, ldc, bias, R, ldr, M, N, K, m_dev, relu,
                                                                   nkb, 0, qkv, ist);
        REGTR_CHECK_LAUNCH();
        return REGTR_OK;
    }
    const int per = regtr_cdiv(nkb, splits);
    const int z = regtr_cdiv(nkb, per);                     // every plane gets >= 1 k-block
    const size_t stride = (size_t)M * N;
    dim3 grid(regtr_cdiv(M, BM), regtr_cdiv(N, BN), z);
    k_gemm_tf32x3_ts<BN, NACC, ST><<<grid, 192, P::SMEM, st>>>(tA, tBh, tBl, ws, N, nullptr, nullptr, 0, M, N, K, m_dev, 0,
                                                               per, stride, QkvOut{nullptr, 0, nullptr, 0, 0},
                                                               InStats{nullptr, 0, nullptr, nullptr, nullptr, 0.f});
    REGTR_CHECK_LAUNCH();
    if (ist.acc) {
        k_splitk_reduce_stats<<<dim3(regtr_cdiv(M, 4 * RS_ROWS), regtr_cdiv(N, 128)), 128, 0, st>>>(ws, z, stride, C, ldc, bias, R, ldr,
                                                                                          M, N, m_dev, relu, ist);
    } else {
        k_splitk_reduce<<<regtr_cdiv((long long)M * (N / 4), 256), 256, 0, st>>>(ws, z, stride, C, ldc, bias, R, ldr, M, N,
                                                                                m_dev, relu);
    }
    REGTR_CHECK_LAUNCH();
    return REGTR_OK;
}

// development A/B switch: REGTR_GEMM_IMPL=ss selects the shared-memory-A kernel
bool gemm_use_ts() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("REGTR_GEMM_IMPL"); v = !(e && e[0] == 's'); }
    return v != 0;
}

}  // namespace

extern "C" {

int regtr_split_tf32(const float* x, long long n, float* hi, float* lo, void* stream_) {
    if (n < 0) return REGTR_ERR_ARG;
    if (n == 0) return REGTR_OK;
    if (!x || !hi || !lo) return REGTR_ERR_ARG;
    k_split_tf32<<<regtr_cdiv(n, 256), 256, 0, (cudaStream_t)stream_>>>(x, n, hi, lo);
    REGTR_CHECK_LAUNCH();
    return REGTR_OK;
}

size_t regtr_gemm_ws_bytes(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 256;
    const int s = choose_splits(M, N, K, choose_bn(M, N));
    return s > 1 ? regtr_align((size_t)s * M * N * sizeof(float)) : 256;
}

static int gemm_dispatch(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                         const float* bias, const float* R, int ldr, int M, int N, int K, const int32_t* m_dev,
                         int relu, void* ws, size_t ws_bytes, void* stream_, InStats ist) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (M < 0 || N <= 0 || K <= 0) return REGTR_ERR_ARG;
    if (M == 0) return REGTR_OK;
    if (!A || !B_hi || !B_lo || !C) return REGTR_ERR_ARG;
    // TMA: 16-byte aligned bases and row pitches
    if ((lda & 3) || (ldb & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B_hi & 15) || ((uintptr_t)B_lo & 15))
        return REGTR_ERR_UNSUPPORTED;
    // Tile width: the widest BN that still yields >= ~1 wave of CTAs (148 SMs); short-K problems
    // take 2 pipeline stages so that two CTAs fit one SM (their fixed prologue/epilogue overlap);
    // skinny long-K problems are split along K (deterministic two-pass reduction).
    const int bn = choose_bn(M, N);
    int splits = choose_splits(M, N, K, bn);
    if (splits > 1 && (!ws || ws_bytes < regtr_gemm_ws_bytes(M, N, K))) return REGTR_ERR_WORKSPACE;
    const bool shortk = K <= 128;
    if (gemm_use_ts()) {
        // accumulation runs per TMEM accumulator: k-blocks per split / NACC (the tensor core adds with truncation)
        const int nkb_split = regtr_cdiv(regtr_cdiv(K, BK), splits);
#define REGTR_TS_CASE(BN_, NACC_, ST_)                                                                                 \
        return launch_gemm_ts<BN_, NACC_, ST_>(A, lda, B_hi, B_lo, ldb, C, ldc, bias, R, ldr, M, N, K, m_dev, relu, splits, \
                                               (float*)ws, st, QkvOut{nullptr, 0, nullptr, 0, 0}, ist)
        if (bn == 128) { if (nkb_split <= 16) REGTR_TS_CASE(128, 1, 2); REGTR_TS_CASE(128, 2, 4); }
        if (bn == 64) REGTR_TS_CASE(64, 2, 3);
        REGTR_TS_CASE(32, 4, 4);
#undef REGTR_TS_CASE
    }
    if (ist.acc) return REGTR_ERR_UNSUPPORTED;             // statistics epilogue: TMEM-A kernel only
#define REGTR_GEMM_CASE(BN_, ST_)                                                                              \
    return launch_gemm<BN_, ST_>(A, lda, B_hi, B_lo, ldb, C, ldc, bias, R, ldr, M, N, K, m_dev, relu, splits, \
                                 (float*)ws, st)
    // 2-stage variants keep the CTA under half an SM's shared memory (2 CTAs/SM): latency is hidden by
    // the co-resident CTA instead of a deeper pipeline (measured: GEMMs were 43 % of the pair time with
    // 1 CTA/SM because a resident CTA mostly waits and blocks other pairs' CTAs from the SM)
    // Small grids (<= one CTA per SM) are latency-critical: 4 stages put the whole K run in flight at once.
    (void)shortk;
    const bool small_grid = false;   // measured: deeper pipelines for small grids lose 2 % under 6-way overlap
    if (bn == 128) { if (K <= 128) REGTR_GEMM_CASE(128, 2); REGTR_GEMM_CASE(128, 3); }
    if (bn == 64) { if (small_grid) REGTR_GEMM_CASE(64, 4); REGTR_GEMM_CASE(64, 2); }
    if (small_grid) REGTR_GEMM_CASE(32, 4);
    REGTR_GEMM_CASE(32, 2);
#undef REGTR_GEMM_CASE
}

int regtr_gemm_tf32x3(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                      const float* bias, const float* R, int ldr, int M, int N, int K, const int32_t* m_dev,
                      int relu, void* ws, size_t ws_bytes, void* stream_) {
    return gemm_dispatch(A, lda, B_hi, B_lo, ldb, C, ldc, bias, R, ldr, M, N, K, m_dev, relu, ws, ws_bytes, stream_,
                         InStats{nullptr, 0, nullptr, nullptr, nullptr, 0.f});
}

size_t regtr_instnorm_acc_bytes(int n_clouds, int C) {
    return sizeof(unsigned long long) * 3 * (size_t)(n_clouds > 0 ? n_clouds : 1) * (size_t)(C > 0 ? C : 1) + 256;
}

int regtr_gemm_tf32x3_instats(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                              int M, int N, int K, const int32_t* m_dev, const int32_t* offs, int n_clouds, float eps,
                              void* acc, float* stats, void* ws, size_t ws_bytes, void* stream_) {
    if (!offs || n_clouds <= 0 || !acc || !stats) return REGTR_ERR_ARG;
    if (N % 32 != 0 || ((uintptr_t)acc & 255)) return REGTR_ERR_UNSUPPORTED;
    // acc: regtr_instnorm_acc_bytes(n_clouds, N): one 256-byte header (completion counter) + the accumulators
    InStats ist{offs, n_clouds, (unsigned long long*)((char*)acc + 256), (int32_t*)acc, (float2*)stats, eps};
    return gemm_dispatch(A, lda, B_hi, B_lo, ldb, C, ldc, nullptr, nullptr, 0, M, N, K, m_dev, 0, ws, ws_bytes, stream_, ist);
}

// In-projection of the attention block with the bf16 epilogue consumed by regtr_mha_bf16_tc_fwd:
// qk_out [M, split] bf16 (ld_qk), vt_out [N - split, ld_vt] bf16 (transposed v), bias added first.
int regtr_gemm_tf32x3_qkv_bf16(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb,
                               const float* bias, int M, int N, int K, int split, void* qk_out, int ld_qk,
                               void* vt_out, int ld_vt, const int32_t* m_dev, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (M < 0 || N <= 0 || K <= 0 || split <= 0 || split > N) return REGTR_ERR_ARG;
    if (M == 0) return REGTR_OK;
    if (!A || !B_hi || !B_lo || !qk_out || !vt_out) return REGTR_ERR_ARG;
    if ((N & 31) || (split & 31) || (ld_qk & 7) || ((uintptr_t)qk_out & 15) || (lda & 3) || (ldb & 3) ||
        ((uintptr_t)A & 15) || ((uintptr_t)B_hi & 15) || ((uintptr_t)B_lo & 15) || ld_vt < M)
        return REGTR_ERR_UNSUPPORTED;
    QkvOut q{(__nv_bfloat16*)qk_out, ld_qk, (__nv_bfloat16*)vt_out, ld_vt, split};
    float* dummy = reinterpret_cast<float*>(qk_out);      // C is never written in this mode
    const int bn = choose_bn(M, N);
    const bool shortk = K <= 128;
    if (gemm_use_ts()) {
#define REGTR_TSQ_CASE(BN_, NACC_, ST_)                                                                                 \
        return launch_gemm_ts<BN_, NACC_, ST_>(A, lda, B_hi, B_lo, ldb, dummy, 0, bias, nullptr, 0, M, N, K, m_dev, 0, 1, \
                                               nullptr, st, q)
        if (bn == 128) { if (K <= 512) REGTR_TSQ_CASE(128, 1, 2); REGTR_TSQ_CASE(128, 2, 4); }
        if (bn == 64) REGTR_TSQ_CASE(64, 2, 3);
        REGTR_TSQ_CASE(32, 4, 4);
#undef REGTR_TSQ_CASE
    }
#define REGTR_QKV_CASE(BN_, ST_)                                                                                  \
    return launch_gemm<BN_, ST_>(A, lda, B_hi, B_lo, ldb, dummy, 0, bias, nullptr, 0, M, N, K, m_dev, 0, 1, nullptr, \
                                 st, q)
    (void)shortk;
    if (bn == 128) { if (K <= 128) REGTR_QKV_CASE(128, 2); REGTR_QKV_CASE(128, 3); }
    if (bn == 64) REGTR_QKV_CASE(64, 2);
    REGTR_QKV_CASE(32, 2);
#undef REGTR_QKV_CASE
}

}  // extern "C"
