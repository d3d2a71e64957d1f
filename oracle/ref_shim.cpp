// ORACLE -- test infrastructure only (never imported by regtr_b200/).
//
// extern "C" shim around the reference's own CPU C++ core so that it can be
// driven through ctypes.  The reference sources are compiled where they lie
// under /root/reference (see oracle/Makefile, target `ref`); nothing is copied
// into this repository.  The reference's own Python wrappers
// (cpp_wrappers/*/wrapper.cpp) do not build against NumPy 2, so
// this shim replaces only that binding layer:
//
//   regtr_ref_batch_neighbors  -> batch_nanoflann_neighbors
//       /root/reference/src/models/backbone_kpconv/cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332
//       (called by the reference wrapper at cpp_neighbors/wrapper.cpp:198)
//   regtr_ref_batch_subsample  -> batch_grid_subsampling
//       .../cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211
//       (called by the reference wrapper at cpp_subsampling/wrapper.cpp:254)
//
// Semantics differ from the GPU path the product mirrors (K *nearest*, sorted by
// distance, hash-map voxel order): this library is the CPU-baseline
// pre-processor ("ORACLE-C" in SURVEY.md 8c) and a cross-check of level sizes,
// not the index-parity oracle.
#include <cstdint>
#include <cstring>
#include <vector>

#include "cpp_neighbors/neighbors/neighbors.h"
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"

extern "C" {

// Returns max_count (row width).  Call once with out == nullptr to get the
// width, or pass a buffer of capacity `cap_cols` columns; rows are written with
// stride `cap_cols` and truncated to it (the Python caller slices [:, :K] as the
// reference does at kpconv.py:254-258).
int regtr_ref_batch_neighbors(const float* q, int64_t nq, const float* s, int64_t ns,
                              const int32_t* q_lens, const int32_t* s_lens, int n_clouds,
                              float radius, int32_t* out, int cap_cols)
{
    std::vector<PointXYZ> queries((size_t)nq), supports((size_t)ns);
    for (int64_t i = 0; i < nq; ++i) queries[i] = PointXYZ(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
    for (int64_t i = 0; i < ns; ++i) supports[i] = PointXYZ(s[3 * i], s[3 * i + 1], s[3 * i + 2]);
    std::vector<int> qb(q_lens, q_lens + n_clouds), sb(s_lens, s_lens + n_clouds);
    std::vector<int> idx;
    batch_nanoflann_neighbors(queries, supports, qb, sb, idx, radius);
    const int width = nq > 0 ? (int)(idx.size() / (size_t)nq) : 0;
    if (out != nullptr) {
        const int w = width < cap_cols ? width : cap_cols;
        for (int64_t i = 0; i < nq; ++i) {
            for (int j = 0; j < w; ++j) out[i * cap_cols + j] = idx[(size_t)i * width + j];
            for (int j = w; j < cap_cols; ++j) out[i * cap_cols + j] = (int32_t)ns;
        }
    }
    return width;
}

// out_xyz capacity: n points; returns number of subsampled points.
int64_t regtr_ref_batch_subsample(const float* xyz, int64_t n, const int32_t* lens, int n_clouds,
                                  float dl, float* out_xyz, int32_t* out_lens)
{
    std::vector<PointXYZ> pts((size_t)n), sub;
    for (int64_t i = 0; i < n; ++i) pts[i] = PointXYZ(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    std::vector<float> f0, f1;
    std::vector<int> c0, c1, b1;
    std::vector<int> b0(lens, lens + n_clouds);
    batch_grid_subsampling(pts, sub, f0, f1, c0, c1, b0, b1, dl, 0);
    for (size_t i = 0; i < sub.size(); ++i) {
        out_xyz[3 * i] = sub[i].x; out_xyz[3 * i + 1] = sub[i].y; out_xyz[3 * i + 2] = sub[i].z;
    }
    for (int c = 0; c < n_clouds; ++c) out_lens[c] = b1[c];
    return (int64_t)sub.size();
}

}  // extern "C"
