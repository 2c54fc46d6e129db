// Weighted Kabsch for many correspondence sets on sm_100a.
//
// Replaces compute_rigid_transform() (reference utils/se3_torch.py:131-173) and
// fast_compute_rigid_transform() (:226-274):
//   w' = w > threshold ? w : 0                                   (:240-242, fast variant only)
//   wn = w' / max(sum w', 1e-6);  ca = sum wn*a;  cb = sum wn*b  (:149-152)
//   cov = (a - ca)^T ((b - cb) * wn)                             (:153-155)
//   U S V^T = svd(cov);  R = V diag(1,1,sign det(V U^T)) U^T      (:163-168)
//   t = cb - R ca                                                 (:171)
// One CTA per set: a block-wide fp64 reduction of the 16 moments
//   {sum w, sum w a, sum w b, sum w a b^T}  (cov follows from them in closed form),
// then the CTA's first warp solves the 3x3 SVD with one-sided Jacobi rotations in fp64: lane i owns row i of A and of V,
// the column inner products of a rotation are three-lane shuffle sums, every lane derives the same (c, s) and rotates
// its own row.  The reference works
// in fp32 with LAPACK; parity is on the assembled [R|t] (singular vectors are sign/order
// ambiguous), within 1e-3 deg and 1e-5 m.
#include "common.cuh"

namespace kpreg {
namespace {

constexpr int kKabschThreads = 256;

__device__ void svd3_rotation(const double cov[3][3], double R[3][3]) {
  // One-sided Jacobi (Hestenes) on the columns of A = cov, accumulating V: A V = U S.
  double A[3][3], V[3][3];
  double scale = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { A[i][j] = cov[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; scale = fmax(scale, fabs(cov[i][j])); }
  if (!(scale > 0.0) || !isfinite(scale)) {
    // zero (or non-finite) covariance: LAPACK returns U = V = I for the zero matrix, hence R = I
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
    return;
  }
  // Jacobi sweeps across the warp: lane r (< 3) owns row r of A and of V; all 32 lanes execute the shuffles
  {
    const int lane = threadIdx.x & 31;
    const int r = lane < 3 ? lane : 0;
    double ar[3] = {A[r][0], A[r][1], A[r][2]}, vr[3] = {V[r][0], V[r][1], V[r][2]};
    const double live = lane < 3 ? 1.0 : 0.0;
    auto sum3 = [](double v) {  // lanes 0..2 -> every lane
      return __shfl_sync(0xffffffffu, v, 0) + __shfl_sync(0xffffffffu, v, 1) + __shfl_sync(0xffffffffu, v, 2);
    };
    for (int sweep = 0; sweep < 30; ++sweep) {
      double off = 0.0;
      for (int p = 0; p < 2; ++p) {
        for (int q = p + 1; q < 3; ++q) {
          const double alpha = sum3(live * ar[p] * ar[p]), beta = sum3(live * ar[q] * ar[q]), gamma = sum3(live * ar[p] * ar[q]);
          if (gamma == 0.0) continue;  // warp-uniform
          off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
          const double ap = ar[p], aq = ar[q], vp = vr[p], vq = vr[q];
          ar[p] = c * ap - sn * aq;
          ar[q] = sn * ap + c * aq;
          vr[p] = c * vp - sn * vq;
          vr[q] = sn * vp + c * vq;
        }
      }
      if (off < 1e-15) break;  // warp-uniform
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        A[i][j] = __shfl_sync(0xffffffffu, ar[j], i);
        V[i][j] = __shfl_sync(0xffffffffu, vr[j], i);
      }
  }
  // singular values = column norms; order them descending (the reference flips the LAST column)
  double sig[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (sig[ord[b]] > sig[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
  double U[3][3], Vs[3][3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) Vs[i][j] = V[i][ord[j]];
  const double tiny = 1e-14 * sig[ord[0]];
  // U columns: normalised A columns where the singular value is resolvable, completed orthonormally otherwise
  for (int j = 0; j < 3; ++j) {
    const double s = sig[ord[j]];
    for (int i = 0; i < 3; ++i) U[i][j] = s > tiny ? A[i][ord[j]] / s : 0.0;
  }
  if (!(sig[ord[1]] > tiny)) {
    // rank 1: any unit vector orthogonal to u0
    int m = fabs(U[0][0]) < fabs(U[1][0]) ? (fabs(U[0][0]) < fabs(U[2][0]) ? 0 : 2) : (fabs(U[1][0]) < fabs(U[2][0]) ? 1 : 2);
    double e[3] = {0, 0, 0};
    e[m] = 1.0;
    const double d = U[m][0];
    double nrm = 0.0;
    for (int i = 0; i < 3; ++i) { U[i][1] = e[i] - d * U[i][0]; nrm += U[i][1] * U[i][1]; }
    nrm = sqrt(nrm);
    for (int i = 0; i < 3; ++i) U[i][1] /= nrm;
  }
  if (!(sig[ord[2]] > tiny)) {
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  // R+ = V U^T; if det <= 0 negate V's last column (reference: torch.where(det > 0, pos, neg))
  double Rp[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rp[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + Vs[i][2] * U[j][2];
  const double det = Rp[0][0] * (Rp[1][1] * Rp[2][2] - Rp[1][2] * Rp[2][1]) - Rp[0][1] * (Rp[1][0] * Rp[2][2] - Rp[1][2] * Rp[2][0]) +
                     Rp[0][2] * (Rp[1][0] * Rp[2][1] - Rp[1][1] * Rp[2][0]);
  const double sgn = det > 0.0 ? 1.0 : -1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + sgn * Vs[i][2] * U[j][2];
}

__global__ void __launch_bounds__(kKabschThreads) k_kabsch(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ w,
                                                           const int64_t* __restrict__ offsets, int64_t pts_per_set, float threshold,
                                                           int write_back, float* __restrict__ out) {
  __shared__ double s_part[kKabschThreads / 32][16];
  const int64_t set = blockIdx.x;
  const int64_t begin = offsets ? offsets[set] : set * pts_per_set;
  const int64_t end = offsets ? offsets[set + 1] : begin + pts_per_set;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // moments: m[0] = sum w; m[1..3] = sum w a; m[4..6] = sum w b; m[7..15] = sum w a_i b_j
  double m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = 0.0;
  for (int64_t i = begin + threadIdx.x; i < end; i += kKabschThreads) {
    float wf = 1.0f;
    if (w) {
      wf = w[i];
      if (threshold >= 0.f) {
        const float wt = wf > threshold ? wf : 0.f;
        if (write_back && wt != wf) w[i] = wt;
        wf = wt;
      }
    }
    const double wd = (double)wf;
    const double ax = a[3 * i], ay = a[3 * i + 1], az = a[3 * i + 2];
    const double bx = b[3 * i], by = b[3 * i + 1], bz = b[3 * i + 2];
    m[0] += wd;
    m[1] += wd * ax; m[2] += wd * ay; m[3] += wd * az;
    m[4] += wd * bx; m[5] += wd * by; m[6] += wd * bz;
    m[7] += wd * ax * bx;  m[8] += wd * ax * by;  m[9] += wd * ax * bz;
    m[10] += wd * ay * bx; m[11] += wd * ay * by; m[12] += wd * ay * bz;
    m[13] += wd * az * bx; m[14] += wd * az * by; m[15] += wd * az * bz;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const double v = warp_sum(m[i]);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  if (warp != 0) return;  // the first warp finishes: totals (every lane), warp-level Jacobi SVD, lane 0 writes [R | t]
  double t[16];
  for (int i = 0; i < 16; ++i) {
    t[i] = 0.0;
    for (int ww = 0; ww < kKabschThreads / 32; ++ww) t[i] += s_part[ww][i];
  }
  // weighted: normaliser max(sum w, 1e-6) (reference _EPS); unweighted: the mean (and cov left unscaled, as in the reference)
  const double denom = w ? fmax(t[0], 1e-6) : fmax(t[0], 1.0);
  const double inv = 1.0 / denom;
  const double sw = t[0] * inv;  // sum of normalised weights (1 unless the clamp is active)
  double ca[3], cb[3];
  for (int d = 0; d < 3; ++d) { ca[d] = t[1 + d] * inv; cb[d] = t[4 + d] * inv; }
  // cov_ij = sum wn (a_i - ca_i)(b_j - cb_j) = S_ab/den - ca_i*cb_j*(2 - sw)   [sum wn a = ca, sum wn b = cb]
  double cov[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[i][j] = t[7 + 3 * i + j] * inv - ca[i] * cb[j] * (2.0 - sw);
  double R[3][3];
  svd3_rotation(cov, R);
  if (lane != 0) return;
  float* o = out + set * 12;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) o[4 * i + j] = (float)R[i][j];
    o[4 * i + 3] = (float)(cb[i] - (R[i][0] * ca[0] + R[i][1] * ca[1] + R[i][2] * ca[2]));
  }
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_kabsch(const float* a, const float* b, float* w, const int64_t* offsets, int64_t n_sets, int64_t pts_per_set,
                            float threshold, int write_back, float* out, void* stream_) {
  if (n_sets < 0 || (!offsets && pts_per_set < 0)) return KPREG_E_INVALID;
  if (n_sets == 0) return KPREG_OK;
  if (!a || !b || !out) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  ProfScope prof(KPREG_FAM_KABSCH, stream);
  k_kabsch<<<(unsigned)n_sets, kKabschThreads, 0, stream>>>(a, b, w, offsets, pts_per_set, threshold, write_back, out);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}
