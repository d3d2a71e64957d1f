// Normalisation / embedding kernels around the GEMMs:
//   per-cloud InstanceNorm (+ residual, + LeakyReLU)  -- kpconv_blocks.py:474-530, 546-561
//   LayerNorm (+ position add)                         -- transformers.py:117-119, 194-232
//   3-D sine position embedding                        -- position_embedding.py:29-50
// (paths relative to /root/reference/src)
#include "common.cuh"

namespace {

constexpr int IN_CH = 1024;   // rows per statistics chunk
constexpr int IN_TY = 8;

// chunk id -> (cloud, first row, last row) ; chunks never straddle clouds.
__device__ __forceinline__ bool chunk_of(const int32_t* __restrict__ offs, int n_clouds, int chunk, int& cloud,
                                         int& r0, int& r1, int& first_chunk, int& n_chunks) {
    int acc = 0;
    for (int c = 0; c < n_clouds; ++c) {
        const int a = offs[c], b = offs[c + 1];
        const int nc = (b - a + IN_CH - 1) / IN_CH;
        if (chunk < acc + nc) {
            cloud = c;
            r0 = a + (chunk - acc) * IN_CH;
            r1 = min(r0 + IN_CH, b);
            first_chunk = acc;
            n_chunks = nc;
            return true;
        }
        acc += nc;
    }
    return false;
}

// partial[chunk][c] = (sum, sum of squares) in fp64 over the chunk's rows.
__global__ void __launch_bounds__(32 * IN_TY)
k_in_stats(const float* __restrict__ x, const int32_t* __restrict__ offs, int n_clouds, int C,
           double2* __restrict__ partial) {
    __shared__ double2 red[IN_TY][32];
    int cloud, r0, r1, fc, nc;
    if (!chunk_of(offs, n_clouds, blockIdx.x, cloud, r0, r1, fc, nc)) return;
    const int c = blockIdx.y * 32 + threadIdx.x;
    double s = 0.0, ss = 0.0;
    if (c < C) {
        for (int r = r0 + threadIdx.y; r < r1; r += IN_TY) {
            const double v = (double)x[(size_t)r * C + c];
            s += v;
            ss += v * v;
        }
    }
    red[threadIdx.y][threadIdx.x] = make_double2(s, ss);
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int t = 1; t < IN_TY; ++t) { s += red[t][threadIdx.x].x; ss += red[t][threadIdx.x].y; }
        partial[(size_t)blockIdx.x * C + c] = make_double2(s, ss);
    }
}

// out = act(norm(x) + res) on the chunk's rows; the cloud's statistics are re-reduced from
// its chunk partials in a fixed order (deterministic, no atomics).
__global__ void __launch_bounds__(32 * IN_TY)
k_in_apply(const float* x, const int32_t* __restrict__ offs, int n_clouds, int C, float eps,
           const double2* __restrict__ partial, const float* res, float slope, float* out) {
    __shared__ float s_mean[32], s_rstd[32];
    int cloud, r0, r1, fc, nc;
    if (!chunk_of(offs, n_clouds, blockIdx.x, cloud, r0, r1, fc, nc)) return;
    const int c = blockIdx.y * 32 + threadIdx.x;
    if (threadIdx.y == 0 && c < C) {
        double s = 0.0, ss = 0.0;
        for (int t = 0; t < nc; ++t) {
            const double2 p = partial[(size_t)(fc + t) * C + c];
            s += p.x;
            ss += p.y;
        }
        const double n = (double)(offs[cloud + 1] - offs[cloud]);
        const double mean = s / n;
        double var = ss / n - mean * mean;          // biased variance (InstanceNorm)
        var = var > 0.0 ? var : 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    if (c >= C) return;
    const float mean = s_mean[threadIdx.x], rstd = s_rstd[threadIdx.x];
    for (int r = r0 + threadIdx.y; r < r1; r += IN_TY) {
        const size_t o = (size_t)r * C + c;
        float v = (x[o] - mean) * rstd;
        if (res) v += res[o];
        if (slope >= 0.f) v = v > 0.f ? v : v * slope;
        out[o] = v;
    }
}

// One warp per row; E <= 1024, multiple of 32.
__global__ void k_layernorm_pos(const float* __restrict__ x, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ pos, int n, int E,
                                float eps, float* __restrict__ y, float* __restrict__ y_pos) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* xr = x + (size_t)row * E;
    float v[32];
    const int per = E / 32;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < per) { v[j] = xr[j * 32 + lane]; s += v[j]; }
    const float mean = warp_sum(s) / (float)E;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < per) { const float d = v[j] - mean; ss += d * d; }
    const float rstd = 1.f / sqrtf(warp_sum(ss) / (float)E + eps);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j >= per) break;
        const int c = j * 32 + lane;
        const float o = (v[j] - mean) * rstd * gamma[c] + beta[c];
        if (y) y[(size_t)row * E + c] = o;
        if (y_pos) y_pos[(size_t)row * E + c] = o + (pos ? pos[(size_t)row * E + c] : 0.f);
    }
}

// out[i, a*n_freq + f] = f even ? sin(v) : cos(v),  v = (xyz[i,a]*scale) / dim_t[f]; zero pad.
__global__ void k_pos_embed_sine(const float* __restrict__ xyz, int n, const float* __restrict__ dim_t, int n_freq,
                                 int d_model, float scale, float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * d_model) return;
    const int i = (int)(t / d_model), c = (int)(t % d_model);
    float o = 0.f;
    if (c < 3 * n_freq) {
        const int a = c / n_freq, f = c % n_freq;
        const float v = __fdiv_rn(__fmul_rn(xyz[3 * i + a], scale), dim_t[f]);
        o = (f & 1) ? cosf(v) : sinf(v);
    }
    out[t] = o;
}

}  // namespace

extern "C" {

size_t regtr_instnorm_ws_bytes(int n_cap, int n_clouds, int C) {
    const size_t chunks = (size_t)regtr_cdiv(n_cap > 0 ? n_cap : 1, IN_CH) + (size_t)(n_clouds > 0 ? n_clouds : 1);
    return regtr_align(chunks * (size_t)(C > 0 ? C : 1) * sizeof(double2));
}

int regtr_instnorm_act(const float* x, const int32_t* offs, int n_clouds, int n_cap, int C, float eps,
                       const float* res, float slope, float* out, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (!offs || n_clouds <= 0 || n_cap < 0 || C <= 0) return REGTR_ERR_ARG;
    if (n_cap == 0) return REGTR_OK;
    if (!x || !out || !ws) return REGTR_ERR_ARG;
    if (ws_bytes < regtr_instnorm_ws_bytes(n_cap, n_clouds, C)) return REGTR_ERR_WORKSPACE;
    const int chunks = regtr_cdiv(n_cap, IN_CH) + n_clouds;
    dim3 grid(chunks, regtr_cdiv(C, 32)), block(32, IN_TY);
    k_in_stats<<<grid, block, 0, st>>>(x, offs, n_clouds, C, (double2*)ws);
    REGTR_CHECK_LAUNCH();
    k_in_apply<<<grid, block, 0, st>>>(x, offs, n_clouds, C, eps, (const double2*)ws, res, slope, out);
    REGTR_CHECK_LAUNCH();
    return REGTR_OK;
}

int regtr_layernorm_pos(const float* x, const float* gamma, const float* beta, const float* pos, int n, int E,
                        float eps, float* y, float* y_pos, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n < 0 || E <= 0 || E % 32 != 0 || E > 1024) return REGTR_ERR_ARG;
    if (n == 0) return REGTR_OK;
    if (!x || !gamma || !beta || (!y && !y_pos)) return REGTR_ERR_ARG;
    k_layernorm_pos<<<regtr_cdiv((long long)n * 32, 256), 256, 0, st>>>(x, gamma, beta, pos, n, E, eps, y, y_pos);
    REGTR_CHECK_LAUNCH();
    return REGTR_OK;
}

int regtr_pos_embed_sine(const float* xyz, int n, const float* dim_t, int n_freq, int d_model, float scale,
                         float* out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n < 0 || n_freq <= 0 || d_model < 3 * n_freq) return REGTR_ERR_ARG;
    if (n == 0) return REGTR_OK;
    if (!xyz || !dim_t || !out) return REGTR_ERR_ARG;
    k_pos_embed_sine<<<regtr_cdiv((long long)n * d_model, 256), 256, 0, st>>>(xyz, n, dim_t, n_freq, d_model, scale, out);
    REGTR_CHECK_LAUNCH();
    return REGTR_OK;
}

}  // extern "C"
