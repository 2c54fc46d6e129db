// TEST INFRASTRUCTURE — not product code.
// extern "C" shim over the UNMODIFIED reference C++ core, compiled from the
// sources where they lie under /root/reference (never copied into this repo).
// It does what the reference's CPython glue does between numpy buffers and
// std::vector<PointXYZ>:
//   cpp_wrappers/cpp_neighbors/wrapper.cpp:188-224   (batch_query)
//   cpp_wrappers/cpp_subsampling/wrapper.cpp:242-300 (subsample_batch)
// The CPython glue itself does not build against numpy >= 2 (NPY_IN_ARRAY is
// gone), hence this shim.  Output: oracle/_ref/libkpref.so (git-ignored).
#include "cpp_neighbors/neighbors/neighbors.h"
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"
#include <cstdlib>
#include <cstring>

extern "C" {

// Returns max_count (row width); *out is malloc'ed [nq * max_count] ints, caller frees with kpref_free.
int kpref_batch_query(const float* q, int nq, const float* s, int ns, const int* q_lens,
                      const int* s_lens, int n_clouds, float radius, int** out) {
  std::vector<PointXYZ> Q((const PointXYZ*)q, (const PointXYZ*)q + nq);
  std::vector<PointXYZ> S((const PointXYZ*)s, (const PointXYZ*)s + ns);
  std::vector<int> QB(q_lens, q_lens + n_clouds), SB(s_lens, s_lens + n_clouds), idx;
  batch_nanoflann_neighbors(Q, S, QB, SB, idx, radius);
  int width = nq > 0 ? (int)(idx.size() / (size_t)nq) : 0;
  *out = (int*)malloc(idx.size() * sizeof(int) + 4);
  memcpy(*out, idx.data(), idx.size() * sizeof(int));
  return width;
}

// Returns M (number of subsampled points); *out is malloc'ed [M*3] floats; out_lens has n_clouds ints.
int kpref_subsample_batch(const float* p, int n, const int* lens, int n_clouds, float dl, int max_p,
                          float** out, int* out_lens) {
  std::vector<PointXYZ> P((const PointXYZ*)p, (const PointXYZ*)p + n), SP;
  std::vector<float> f, sf;
  std::vector<int> c, sc, B(lens, lens + n_clouds), SB;
  batch_grid_subsampling(P, SP, f, sf, c, sc, B, SB, dl, max_p);
  *out = (float*)malloc(SP.size() * sizeof(PointXYZ) + 4);
  memcpy(*out, SP.data(), SP.size() * sizeof(PointXYZ));
  memcpy(out_lens, SB.data(), n_clouds * sizeof(int));
  return (int)SP.size();
}

void kpref_free(void* p) { free(p); }
}
