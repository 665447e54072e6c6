// Test-only host instantiation of the product's __host__ __device__ geometry (megba_b200/csrc/ba_geometry.cuh), so the
// closed-form Jacobian can be checked against the oracle in the no-GPU test tier.  Not part of the product library.
#include <cstdint>

#include "../../megba_b200/csrc/ba_geometry.cuh"

template <typename T>
static void run(int64_t n, const T *cam, const T *pt, const T *obs, T *r, T *J) {
  for (int64_t e = 0; e < n; ++e) {
    const T *c = cam + 9 * e, *X = pt + 3 * e;
    const ba::RotConst<T> k = ba::rot_const<T>(c[0], c[1], c[2]);
    T r0, r1, s0, s1;
    ba::edge_residual_jacobian<T>(c, k, X[0], X[1], X[2], obs[2 * e], obs[2 * e + 1], r0, r1, J + 24 * e);
    ba::edge_residual<T>(c, k, X[0], X[1], X[2], obs[2 * e], obs[2 * e + 1], s0, s1);
    r[4 * e] = r0; r[4 * e + 1] = r1; r[4 * e + 2] = s0; r[4 * e + 3] = s1;  // Jacobian path | residual-only path
  }
}

extern "C" void geo_rj_f64(int64_t n, const double *cam, const double *pt, const double *obs, double *r, double *J) { run<double>(n, cam, pt, obs, r, J); }
extern "C" void geo_rj_f32(int64_t n, const float *cam, const float *pt, const float *obs, float *r, float *J) { run<float>(n, cam, pt, obs, r, J); }
