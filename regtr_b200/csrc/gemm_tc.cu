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

int regtr_gemm_tf32x3(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                      const float* bias, const float* R, int ldr, int M, int N, int K, const int32_t* m_dev,
                      int relu, void* ws, size_t ws_bytes, void* stream_) {
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
