// chop_tracks arithmetic shared by the chop kernels (chop.cu) and the fused raw-row prepare kernel (prepare.cu):
// reference optimize/dataio.py:63-106, numpy's promotion rules op by op (see chop.cu's header).
#pragma once
#include "larnd_common.cuh"

struct ChopGeom {
  float len;       // float32 sqrt(sum(seg**2))
  float dir[3];    // float32 seg / (len + 1e-10)
};

__device__ __forceinline__ ChopGeom chop_geom(const float* tr, const larnd_chop_columns_t& c) {
  ChopGeom g;
  const float sx = __fsub_rn(tr[c.x_end], tr[c.x_start]);
  const float sy = __fsub_rn(tr[c.y_end], tr[c.y_start]);
  const float sz = __fsub_rn(tr[c.z_end], tr[c.z_start]);
  // np.sum over 3 float32 elements: sequential adds, no FMA contraction
  const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz));
  g.len = __fsqrt_rn(s2);
  const float den = __fadd_rn(g.len, 1e-10f);
  g.dir[0] = __fdiv_rn(sx, den);
  g.dir[1] = __fdiv_rn(sy, den);
  g.dir[2] = __fdiv_rn(sz, den);
  return g;
}

__device__ __forceinline__ long long chop_nsteps(float len, float prec32) {
  // np.maximum(np.ceil(length / precision), 1).astype(int): float32 array / weak Python float -> float32
  const float q = ceilf(__fdiv_rn(len, prec32));
  return (long long)fmaxf(q, 1.0f);
}

// start / end point of piece k (of n) along one axis, exactly as k_chop_expand writes them
__device__ __forceinline__ float chop_start(double s0, double d, long long k, double precision) {
  return (float)(s0 + __dmul_rn(__dmul_rn((double)k, precision), d));
}
__device__ __forceinline__ float chop_end(double s0, double d, float e_last, long long k, long long n, double precision) {
  return k == n - 1 ? e_last : (float)(s0 + __dmul_rn(__dmul_rn(precision, (double)(k + 1)), d));
}
__device__ __forceinline__ float chop_dx(const ChopGeom& g, long long k, long long n, double precision, float prec32) {
  return k == n - 1 ? (float)((double)g.len - __dmul_rn(precision, (double)(n - 1))) : prec32;
}
__device__ __forceinline__ float chop_dE(float base, const ChopGeom& g, long long k, long long n, double precision, float prec32) {
  const float len_eps = __fadd_rn(g.len, 1e-10f);            // np.float32 + weak Python float
  return k == n - 1 ? (float)__dmul_rn((double)base, 1.0 - __ddiv_rn(__dmul_rn(precision, (double)(n - 1)), (double)len_eps))
                    : __fdiv_rn(__fmul_rn(base, prec32), len_eps);
}
