// fp32-accurate variable-length multi-head attention core on the tcgen05 tensor cores (3xTF32 split, fp32
// accumulation in TMEM, fp32 softmax): the parity-mode MHA core of TransformerCrossEncoderLayer.forward_pre
// (/root/reference/src/models/transformer/transformers.py:197-226) with both contractions -- S = Q K^T and
// O = P V -- fed by TMA and issued as tcgen05.mma.kind::tf32.  head_dim = 32.
//
// Inputs come from the in-projection GEMM's split epilogue (gemm_tc.cu, regtr_gemm_tf32x3_qkv_split):
//   qk4 [N_tokens, 4E] fp32 : [Q_hi | Q_lo | K_hi | K_lo], q pre-scaled by scale * log2(e); x = hi + lo with both
//                             halves exactly representable in TF32                       (TMA, SWIZZLE_128B)
//   vt2 [2E, ld_vt]    fp32 : v TRANSPOSED (rows = channels, hi then lo; tokens contiguous), so that P V is a
//                             K-major x K-major MMA                                       (TMA, SWIZZLE_128B)
// CTA = (problem, head, tile of 128 queries); 6 warps:
//   warp 0     TMA producer (Q_hi / Q_lo once; per 64-key tile K_hi [+ K_lo, Vt_hi, Vt_lo in pass 2], 2 stages)
//   warp 1     TMEM allocator + MMA issuer
//   warps 2-5  softmax: thread t owns query row t == TMEM lane t
// Two passes over the keys:
//   pass 1  S ~ Q_hi K_hi^T (one TF32 MMA per k-step: the row maximum only stabilises the exponentials, any
//           value within a few per cent of it serves) -> row maximum m;
//   pass 2  S = Q_lo K_hi^T + Q_hi K_lo^T + Q_hi K_hi^T (fp32-accurate), p = exp2(s - m), row sums in registers,
//           p split into (hi, lo) TF32 halves and stored with tcgen05.st into TENSOR MEMORY -- P_hi over the S
//           columns it came from, P_lo beside them -- where the P V MMAs read it as their A operand
//           (O += P_lo V_hi + P_hi V_lo + P_hi V_hi).  P never touches shared memory, the accumulator is never
//           rescaled.
// S is double buffered in TMEM so that Q K^T of tile i+1 overlaps the softmax of tile i; 224 TMEM columns and
// 97 KB of shared memory per CTA: two CTAs per SM.
#include "common.cuh"
#include "tc.cuh"

namespace {

constexpr int HD = 32;          // head dim (floats: one 128-byte swizzle row)
constexpr int BQ = 128;         // queries per CTA
constexpr int BKEY = 64;        // keys per tile
constexpr int TMEM_COLS = 256;  // S0 / P_hi0 [0,64)  S1 / P_hi1 [64,128)  P_lo [128,192)  O even tiles [192,224)  O odd tiles [224,256)
constexpr uint32_t COL_PLO = 128, COL_O = 192;

constexpr int Q_BYTES = BQ * HD * 4;        // 16 KB per half (hi / lo)
constexpr int K_BYTES = BKEY * HD * 4;      // 8 KB per half
constexpr int V_HALF = HD * 32 * 4;         // 4 KB: 32 channel rows x 32 keys; a 64-key tile = 2 halves per (hi / lo)
constexpr int STAGE_BYTES = 2 * K_BYTES + 4 * V_HALF;      // K_hi K_lo | Vt_hi(0,1) Vt_lo(0,1) = 32 KB
constexpr int SMEM_BYTES = 2 * Q_BYTES + 2 * STAGE_BYTES + 1024 + 256;
constexpr uint32_t HI_MASK = 0xFFFFE000u;

__device__ __forceinline__ uint32_t tf32_rn(float x) { return (__float_as_uint(x) + 0x1000u) & HI_MASK; }
__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(192, 2)
k_mha_tf32_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmVt, float* __restrict__ O, int ldo, int E,
              const int32_t* __restrict__ q_start, const int32_t* __restrict__ q_len,
              const int32_t* __restrict__ k_start, const int32_t* __restrict__ k_len,
              const int32_t* __restrict__ tile_base, int n_prob) {
    extern __shared__ unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int prob = blockIdx.z, qt = blockIdx.x;
    const int head = blockIdx.y;
    if (tile_base) {                                             // linear 128-query tile index -> (problem, tile)
        if (qt >= tile_base[n_prob]) return;
        int p = 0;
        while (p + 1 < n_prob && tile_base[p + 1] <= qt) ++p;
        prob = p;
        qt -= tile_base[p];
    }
    const int ql = q_len[prob];
    if (qt * BQ >= ql) return;                                   // uniform exit
    const int q0 = q_start[prob] + qt * BQ, k0 = k_start[prob], kl = k_len[prob];
    if (kl <= 0) {                                                // no keys: zero rows (uniform exit)
        for (int t = threadIdx.x; t < BQ * HD; t += blockDim.x) {
            const int r = t / HD, d = t % HD;
            if (qt * BQ + r < ql) O[(size_t)(q0 + r) * ldo + head * HD + d] = 0.f;
        }
        return;
    }
    // Key tiles start at a token index aligned to 4 (16 bytes of fp32): the inner TMA coordinate of the transposed
    // V must be 16-byte aligned; keys outside [k0, k0 + kl) are masked in the softmax.
    const int ka = k0 & ~3;
    const int n_kt = (k0 - ka + kl + BKEY - 1) / BKEY;
    const int n_it = 2 * n_kt;                                   // pass 1 (max) + pass 2 (exp, PV)

    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sQhi = base;
    unsigned char* sQlo = base + Q_BYTES;
    auto stage = [&](int s) { return base + 2 * Q_BYTES + s * STAGE_BYTES; };
    auto sKhi = [&](int s) { return stage(s); };
    auto sKlo = [&](int s) { return stage(s) + K_BYTES; };
    auto sVhi = [&](int s, int h) { return stage(s) + 2 * K_BYTES + h * V_HALF; };
    auto sVlo = [&](int s, int h) { return stage(s) + 2 * K_BYTES + (2 + h) * V_HALF; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * Q_BYTES + 2 * STAGE_BYTES);
    uint64_t* q_full = bars;            // 1
    uint64_t* kv_full = bars + 1;       // [2] TMA landed
    uint64_t* kv_empty = bars + 3;      // [2] MMAs that read the stage retired (commit)
    uint64_t* s_full = bars + 5;        // [2] S = QK^T ready in TMEM (commit)
    uint64_t* s_empty = bars + 7;       // [2] pass 1: softmax finished reading S (4 warp arrivals)
    uint64_t* p_full = bars + 9;        // [2] P halves stored to TMEM (4 warp arrivals)
    uint64_t* pv_done = bars + 11;      // [2] PV MMAs of the tile retired (commit): S/P_hi buffer and P_lo reusable
    uint64_t* o_full = bars + 13;       // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    if (warp == 0 && lane == 0) {
        tc::mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&kv_full[i], 1); tc::mbar_init(&kv_empty[i], 1);
            tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 4);
            tc::mbar_init(&p_full[i], 4); tc::mbar_init(&pv_done[i], 1);
        }
        tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmVt);
    }
    if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(q_full, 2 * Q_BYTES);
            tc::tma_load_2d(sQhi, &tmQ, q_full, head * HD, q0);
            tc::tma_load_2d(sQlo, &tmQ, q_full, E + head * HD, q0);
            for (int it = 0; it < n_it; ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                const int kt = it < n_kt ? it : it - n_kt;
                const int key0 = ka + kt * BKEY;
                tc::mbar_wait(&kv_empty[s], ph ^ 1);
                if (it < n_kt) {                                 // pass 1: K_hi only
                    tc::mbar_arrive_expect_tx(&kv_full[s], K_BYTES);
                    tc::tma_load_2d(sKhi(s), &tmK, &kv_full[s], 2 * E + head * HD, key0);
                } else {
                    tc::mbar_arrive_expect_tx(&kv_full[s], 2 * K_BYTES + 4 * V_HALF);
                    tc::tma_load_2d(sKhi(s), &tmK, &kv_full[s], 2 * E + head * HD, key0);
                    tc::tma_load_2d(sKlo(s), &tmK, &kv_full[s], 3 * E + head * HD, key0);
                    tc::tma_load_2d(sVhi(s, 0), &tmVt, &kv_full[s], key0, head * HD);
                    tc::tma_load_2d(sVhi(s, 1), &tmVt, &kv_full[s], key0 + 32, head * HD);
                    tc::tma_load_2d(sVlo(s, 0), &tmVt, &kv_full[s], key0, E + head * HD);
                    tc::tma_load_2d(sVlo(s, 1), &tmVt, &kv_full[s], key0 + 32, E + head * HD);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = tc::umma_idesc(tc::FMT_TF32, BQ, BKEY);
            constexpr uint32_t idesc_o = tc::umma_idesc(tc::FMT_TF32, BQ, HD);
            tc::mbar_wait(q_full, 0);
            const uint64_t dQhi = tc::umma_desc_sw128_kmajor(tc::smem_u32(sQhi));
            const uint64_t dQlo = tc::umma_desc_sw128_kmajor(tc::smem_u32(sQlo));
            auto issue_qk = [&](int it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                tc::mbar_wait(&kv_full[s], ph);
                // S[s] was last used by iteration it - 2: read by the softmax (pass 1) or, as P_hi, by its PV MMAs (pass 2)
                if (it >= 2) {
                    if (it - 2 < n_kt) tc::mbar_wait(&s_empty[s], ((it - 2) >> 1) & 1);
                    else tc::mbar_wait(&pv_done[s], ((it - 2 - n_kt) >> 1) & 1);
                }
                tc::fence_after_thread_sync();
                const uint64_t dKhi = tc::umma_desc_sw128_kmajor(tc::smem_u32(sKhi(s)));
                const uint32_t acc = tmem + (uint32_t)(s * BKEY);
                if (it < n_kt) {
#pragma unroll
                    for (int k = 0; k < HD / 8; ++k)               // UMMA_K = 8 tf32 = 32 bytes
                        tc::umma_tf32(acc, dQhi + (uint64_t)(2 * k), dKhi + (uint64_t)(2 * k), idesc_s, k != 0);
                    tc::umma_commit(&s_full[s]);
                    tc::umma_commit(&kv_empty[s]);              // pass 1: the stage is free once QK^T retired
                } else {
                    const uint64_t dKlo = tc::umma_desc_sw128_kmajor(tc::smem_u32(sKlo(s)));
#pragma unroll
                    for (int k = 0; k < HD / 8; ++k) {
                        const uint64_t adv = (uint64_t)(2 * k);
                        tc::umma_tf32(acc, dQlo + adv, dKhi + adv, idesc_s, k != 0);       // small terms first
                        tc::umma_tf32(acc, dQhi + adv, dKlo + adv, idesc_s, 1);
                        tc::umma_tf32(acc, dQhi + adv, dKhi + adv, idesc_s, 1);
                    }
                    tc::umma_commit(&s_full[s]);
                }
            };
            issue_qk(0);
            for (int it = 0; it < n_it; ++it) {
                if (it + 1 < n_it) issue_qk(it + 1);            // overlaps the softmax of tile `it`
                if (it >= n_kt) {
                    const int s = it & 1;
                    const int j = it - n_kt;                    // pass-2 tile index; its S / P_hi buffer is s
                    tc::mbar_wait(&p_full[s], (j >> 1) & 1);
                    tc::fence_after_thread_sync();
                    const uint32_t p_hi = tmem + (uint32_t)(s * BKEY), p_lo = tmem + COL_PLO;
#pragma unroll
                    for (int k = 0; k < BKEY / 8; ++k) {
                        const int h = k >> 2;
                        const uint64_t adv = (uint64_t)(2 * (k & 3));
                        const uint64_t dVhi = tc::umma_desc_sw128_kmajor(tc::smem_u32(sVhi(s, h))) + adv;
                        const uint64_t dVlo = tc::umma_desc_sw128_kmajor(tc::smem_u32(sVlo(s, h))) + adv;
                        // two output accumulators (even / odd key tiles), added in the epilogue: the tensor core
                        // truncates when it accumulates, and the bias grows with the length of the chain
                        const uint32_t o_acc = tmem + COL_O + (uint32_t)(32 * (j & 1));
                        tc::umma_tf32_ts(o_acc, p_lo + 8 * k, dVhi, idesc_o, ((j >> 1) | k) != 0);
                        tc::umma_tf32_ts(o_acc, p_hi + 8 * k, dVlo, idesc_o, 1);
                        tc::umma_tf32_ts(o_acc, p_hi + 8 * k, dVhi, idesc_o, 1);
                    }
                    tc::umma_commit(&pv_done[s]);
                    tc::umma_commit(&kv_empty[s]);
                }
            }
            tc::umma_commit(o_full);
        }
    } else {
        const int q = warp & 3;                                // TMEM lane quarter of this warp
        const int row = q * 32 + lane;                         // query row within the tile
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float m = -INFINITY, l = 0.f;
        for (int it = 0; it < n_it; ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int kt = it < n_kt ? it : it - n_kt;
            const int jlo = max(0, k0 - (ka + kt * BKEY));      // keys [jlo, jhi) of this tile belong to the problem
            const int jhi = min(BKEY, k0 + kl - (ka + kt * BKEY));
            tc::mbar_wait(&s_full[s], ph);
            tc::fence_after_thread_sync();
            float v[BKEY];
            tc::tmem_ld_32x32(tmem + lane_addr + (uint32_t)(s * BKEY), v);
            tc::tmem_ld_32x32(tmem + lane_addr + (uint32_t)(s * BKEY + 32), v + 32);
            if (it < n_kt) {
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s_empty[s]);    // S buffer may be overwritten
#pragma unroll
                for (int j = 0; j < BKEY; ++j) if (j >= jlo && j < jhi) m = fmaxf(m, v[j]);
            } else {
                const int jt = it - n_kt;
                const uint32_t dst_hi = tmem + lane_addr + (uint32_t)(s * BKEY), dst_lo = tmem + lane_addr + COL_PLO;
                // sweep 1: p = exp2(s - m) (kept in v[]), row sum, P_hi over the S columns this thread just read --
                // no hazard with the PV MMAs of the previous tile, which read the OTHER S / P_hi buffer
#pragma unroll
                for (int c = 0; c < BKEY / 16; ++c) {
                    uint32_t hi[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int j = 16 * c + e;
                        const float p = (j >= jlo && j < jhi) ? fast_exp2(v[j] - m) : 0.f;
                        l += p;
                        v[j] = p;
                        hi[e] = tf32_rn(p);
                    }
                    tc::tmem_st_32x16(dst_hi + 16 * c, hi);
                }
                // P_lo is single buffered: the PV MMAs of the previous pass-2 tile must have retired before it is
                // overwritten (the exponentials above ran while they executed)
                if (jt >= 1) tc::mbar_wait(&pv_done[s ^ 1], ((jt - 1) >> 1) & 1);
                tc::fence_after_thread_sync();
#pragma unroll
                for (int c = 0; c < BKEY / 16; ++c) {
                    uint32_t lo[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float p = v[16 * c + e];
                        lo[e] = tf32_rn(p - __uint_as_float(tf32_rn(p)));
                    }
                    tc::tmem_st_32x16(dst_lo + 16 * c, lo);
                }
                tc::tmem_st_wait();
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&p_full[s]);
            }
        }
        tc::mbar_wait(o_full, 0);
        tc::fence_after_thread_sync();
        float o[HD];
        tc::tmem_ld_32x32(tmem + lane_addr + COL_O, o);
        if (n_kt > 1) {                                          // odd tiles' accumulator (never written otherwise)
            float o1[HD];
            tc::tmem_ld_32x32(tmem + lane_addr + COL_O + 32, o1);
#pragma unroll
            for (int j = 0; j < HD; ++j) o[j] += o1[j];
        }
        if (qt * BQ + row < ql) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            float4* dst = reinterpret_cast<float4*>(O + (size_t)(q0 + row) * ldo + head * HD);
#pragma unroll
            for (int j = 0; j < HD / 4; ++j)
                dst[j] = make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem);
}

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

// row-major fp32 matrix [rows, cols] with leading dimension ld; box = [32 cols, box_rows], 128-byte swizzle
bool make_map_f32(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" int regtr_mha_tf32_tc_fwd(const float* qk4, int ld4, const float* vt2, int ld_vt, int n_tokens, float* O,
                                     int ldo, const int32_t* q_start, const int32_t* q_len, const int32_t* k_start,
                                     const int32_t* k_len, int n_problems, int max_q_len, const int32_t* tile_base,
                                     int max_tiles, int n_heads, int head_dim, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_problems < 0 || max_q_len < 0 || n_heads <= 0 || n_tokens < 0 || max_tiles < 0) return REGTR_ERR_ARG;
    if (head_dim != HD) return REGTR_ERR_UNSUPPORTED;
    if (n_problems == 0 || max_q_len == 0 || n_tokens == 0 || (tile_base && max_tiles == 0)) return REGTR_OK;
    if (!qk4 || !vt2 || !O || !q_start || !q_len || !k_start || !k_len) return REGTR_ERR_ARG;
    const int E = n_heads * HD;
    if ((ld4 & 3) || (ld_vt & 3) || ld4 < 4 * E || ((uintptr_t)qk4 & 15) || ((uintptr_t)vt2 & 15) || (ldo & 3) ||
        n_problems > 65535)
        return REGTR_ERR_UNSUPPORTED;
    CUtensorMap tQ, tK, tV;
    if (!make_map_f32(&tQ, qk4, n_tokens, 4 * E, ld4, BQ) || !make_map_f32(&tK, qk4, n_tokens, 4 * E, ld4, BKEY) ||
        !make_map_f32(&tV, vt2, 2 * E, ld_vt, ld_vt, HD))
        return REGTR_ERR_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_mha_tf32_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return -(1000 + (int)e);
        attr_set = true;
    }
    const dim3 grid = tile_base ? dim3(max_tiles, n_heads, 1) : dim3(regtr_cdiv(max_q_len, BQ), n_heads, n_problems);
    k_mha_tf32_tc<<<grid, 192, SMEM_BYTES, st>>>(tQ, tK, tV, O, ldo, E, q_start, q_len, k_start, k_len, tile_base,
                                                 n_problems);
    REGTR_CHECK_LAUNCH();
    return REGTR_OK;
}
