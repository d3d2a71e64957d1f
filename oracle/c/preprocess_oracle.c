/*
 * ORACLE -- test infrastructure only.  Never imported by the product path
 * (regtr_b200/); used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker / CPU baseline.
 *
 * Plain-C CPU restatement of the two un-vendored third-party operations the
 * reference's PreprocessorGPU calls:
 *
 *   - pytorch3d.ops.ball_query (PyTorch3D 0.6.0), call site
 *     /root/reference/src/models/backbone_kpconv/kpconv.py:261-288
 *   - MinkowskiEngine.SparseTensor(quantization_mode=UNWEIGHTED_AVERAGE)
 *     (ME 0.5.4), call site kpconv.py:213-240
 *
 * Neither library is present in /root/reference, so their published semantics
 * are restated here under the pinned determinism rules of DESIGN.md ("H1"):
 *   ball query : per cloud, per query, scan supports in ascending index, keep j
 *                while ((dx*dx + dy*dy) + dz*dz) < r*r (fp32, no FMA
 *                contraction), first K hits, pad with the total support count.
 *   voxel mean : voxel = floor(p / dl) per axis with IEEE fp32 division;
 *                output order = ascending (cloud, vx, vy, vz); barycentre =
 *                fp32 running sum in ascending original index, then one fp32
 *                division by the member count.
 * "parity unpinned" for the third-party halves: the reference ships no test or
 * golden vector for either op (the GPU path is documented as non-deterministic,
 * Readme.md:99), so these rules are the contract.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- ball query */

/* q:(Nq,3) s:(Ns,3) packed; q_lens/s_lens:(C); out:(Nq,K) int64, pad = Ns. */
int oracle_ball_query(const float *q, const int64_t *q_lens, const float *s,
                      const int64_t *s_lens, int n_clouds, int K, float radius,
                      int64_t *out)
{
    int64_t Ns = 0, Nq = 0;
    for (int c = 0; c < n_clouds; ++c) { Ns += s_lens[c]; Nq += q_lens[c]; }
    const float r2 = radius * radius;
    int64_t q0 = 0, s0 = 0;
    for (int c = 0; c < n_clouds; ++c) {
        const int64_t nq = q_lens[c], ns = s_lens[c];
        for (int64_t i = 0; i < nq; ++i) {
            const float *p = q + 3 * (q0 + i);
            int64_t *row = out + (q0 + i) * (int64_t)K;
            int found = 0;
            for (int64_t j = 0; j < ns && found < K; ++j) {
                const float *t = s + 3 * (s0 + j);
                const float dx = p[0] - t[0], dy = p[1] - t[1], dz = p[2] - t[2];
                const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
                const float d2 = (xx + yy) + zz;
                if (d2 < r2) row[found++] = s0 + j;
            }
            for (; found < K; ++found) row[found] = Ns;
        }
        q0 += nq; s0 += ns;
    }
    (void)Nq;
    return 0;
}

/* ------------------------------------------------------------- grid subsample */

typedef struct { int32_t c, x, y, z; int64_t i; } vox_t;

static int vox_cmp(const void *a, const void *b)
{
    const vox_t *u = (const vox_t *)a, *v = (const vox_t *)b;
    if (u->c != v->c) return u->c < v->c ? -1 : 1;
    if (u->x != v->x) return u->x < v->x ? -1 : 1;
    if (u->y != v->y) return u->y < v->y ? -1 : 1;
    if (u->z != v->z) return u->z < v->z ? -1 : 1;
    return u->i < v->i ? -1 : (u->i > v->i ? 1 : 0);
}

/* xyz:(N,3) lens:(C) -> out_xyz:(<=N,3), out_lens:(C); returns #voxels or <0. */
int64_t oracle_grid_subsample(const float *xyz, const int64_t *lens, int n_clouds,
                              float dl, float *out_xyz, int64_t *out_lens)
{
    int64_t N = 0;
    for (int c = 0; c < n_clouds; ++c) N += lens[c];
    vox_t *v = (vox_t *)malloc(sizeof(vox_t) * (size_t)(N > 0 ? N : 1));
    if (!v) return -1;
    int64_t k = 0;
    for (int c = 0; c < n_clouds; ++c) {
        out_lens[c] = 0;
        for (int64_t j = 0; j < lens[c]; ++j, ++k) {
            v[k].c = c;
            v[k].x = (int32_t)floorf(xyz[3 * k + 0] / dl);
            v[k].y = (int32_t)floorf(xyz[3 * k + 1] / dl);
            v[k].z = (int32_t)floorf(xyz[3 * k + 2] / dl);
            v[k].i = k;
        }
    }
    qsort(v, (size_t)N, sizeof(vox_t), vox_cmp);
    int64_t m = 0;
    for (int64_t a = 0; a < N;) {
        int64_t b = a;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        while (b < N && v[b].c == v[a].c && v[b].x == v[a].x && v[b].y == v[a].y &&
               v[b].z == v[a].z) {
            const float *p = xyz + 3 * v[b].i;
            sx = sx + p[0]; sy = sy + p[1]; sz = sz + p[2];
            ++b;
        }
        const float cnt = (float)(b - a);
        out_xyz[3 * m + 0] = sx / cnt;
        out_xyz[3 * m + 1] = sy / cnt;
        out_xyz[3 * m + 2] = sz / cnt;
        out_lens[v[a].c] += 1;
        ++m;
        a = b;
    }
    free(v);
    return m;
}
