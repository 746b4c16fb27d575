// Per-segment physics shared by the LUT prepare kernel and the MC-current kernels:
// shift_tracks (reference sim_jax.py:109-119), quench (quenching_jax.py:38-75) and drift (drifting_jax.py:19-58),
// written op by op with round-to-nearest intrinsics (no FMA contraction) because pixel ids / ticks are derived
// from these values and must match the reference's float32 evaluation bit for bit.
#pragma once
#include "larnd_common.cuh"

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// jnp.floor_divide for floats (_float_divmod): round((a - fmod(a,b))/b) with sign fix-up
__device__ __forceinline__ float floor_divide_f(float a, float b) {
  float mod = fmodf(a, b);
  float div = fdiv(fsub(a, mod), b);
  if (mod != 0.0f && ((b < 0.0f) != (mod < 0.0f))) div = fsub(div, 1.0f);
  return roundf(div);  // lax.round: half away from zero
}
// jnp.remainder for floats
__device__ __forceinline__ float remainder_f(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.0f && ((m < 0.0f) != (b < 0.0f))) m = fadd(m, b);
  return m;
}


struct SegPhys {
  float x, y, z;          // shifted position
  float q;                // electrons after recombination, lifetime and TPC mask
  float recomb, xi, cos2; // recombination factor and its inputs (for the backward pass)
  float td;               // drift time to the anode
  float sl_cm, sT;        // longitudinal / transverse diffusion sigma in cm
  float z_anode, z_cath;
  int plane;
  bool inside;
};

__device__ __forceinline__ SegPhys segment_physics(const float* tr, const larnd_columns_t& cols, const larnd_params_t& p) {
  SegPhys o;
  // shift_tracks
  float x = fsub(tr[cols.x], p.shift_x);
  float y = fsub(tr[cols.y], p.shift_y);
  float z = fsub(tr[cols.z], p.shift_z);
  float dEdx = tr[cols.dEdx], dE = tr[cols.dE];
  // quench
  float recomb, xi, cos2 = 0.0f;
  if (p.recombination_mode == 2) {
    xi = fdiv(fmul(p.kb, dEdx), p.efield_rho);
    recomb = fdiv(p.Ab, fadd(1.0f, xi));
  } else if (p.recombination_mode == 1) {
    float csi = fdiv(fmul(p.beta, dEdx), p.efield_rho);
    recomb = fmaxf(0.0f, fdiv(logf(fadd(p.alpha, csi)), csi));
    xi = csi;
  } else {
    float zs = fsub(tr[cols.z_start], p.shift_z), ze = fsub(tr[cols.z_end], p.shift_z);
    float cosphi = fdiv(fabsf(fsub(ze, zs)), fadd(tr[cols.dx], 1e-10f));
    float c2 = fmul(cosphi, cosphi);
    float bphi = fdiv(p.beta, __fsqrt_rn(fadd(fsub(1.0f, c2), fmul(p.inv_R2, c2))));
    float csi = fdiv(fmul(bphi, dEdx), p.efield_rho);
    recomb = fmaxf(0.0f, fdiv(logf(fadd(p.alpha, csi)), fadd(csi, 1e-10f)));
    xi = csi;
    cos2 = c2;
  }
  float ne = fmul(fmul(recomb, dE), p.MeVToElectrons);
  // drift: TPC membership (first TPC that contains the point, argmax of the boolean row)
  int plane = 0;
  bool inside = false;
  for (int k = p.n_tpc - 1; k >= 0; --k) {
    float za = p.tpc_borders[k][2][0], zc = p.tpc_borders[k][2][1];
    float zmin = fminf(fsub(zc, p.size_margin), fsub(za, p.size_margin));
    float zmax = fmaxf(fadd(zc, p.size_margin), fadd(za, p.size_margin));
    bool c = x >= fsub(p.tpc_borders[k][0][0], p.size_margin) && x <= fadd(p.tpc_borders[k][0][1], p.size_margin) &&
             y >= fsub(p.tpc_borders[k][1][0], p.size_margin) && y <= fadd(p.tpc_borders[k][1][1], p.size_margin) &&
             z >= zmin && z <= zmax;
    if (c) { plane = k; inside = true; }
  }
  o.z_anode = p.tpc_borders[plane][2][0];
  o.z_cath = p.tpc_borders[plane][2][1];
  float dd = fadd(fabsf(fsub(z, o.z_anode)), 1e-6f);
  float td = fdiv(dd, p.vdrift);
  float life = expf(-fdiv(td, p.lifetime));
  o.q = fmul(fmul(ne, life), inside ? 1.0f : 0.0f);
  o.sl_cm = __fsqrt_rn(fmul(fmul(td, 2.0f), p.long_diff));
  o.sT = __fsqrt_rn(fmul(fmul(td, 2.0f), p.tran_diff));
  o.x = x; o.y = y; o.z = z; o.recomb = recomb; o.xi = xi; o.cos2 = cos2; o.td = td; o.plane = plane; o.inside = inside;
  return o;
}
