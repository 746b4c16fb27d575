// K1b: closed-form chain rule from the per-segment derivatives of the accumulate VJP (d/dq, d/dfrac, d/d(a,b,c),
// d/dWx[5], d/dWy[5]) to the LARND_NPARAMS fitted parameters, through the preparation stage of the same segment
// (reference: drifting_jax.py:42-50, quenching_jax.py:18-35, detsim_jax.py:332-341, sim_jax.py:157-168,406-423).
// Shared by the chunk kernel (accumulate_bwd.cu) and the class-sorted kernel (accumulate_bwd_sorted.cu);
// `add(k, v)` accumulates v into parameter gradient k.
#pragma once
#include "larnd_common.cuh"

template <class Add>
__device__ __forceinline__ void chain_rule_segment(const larnd_params_t& p, const float* __restrict__ rec, int64_t n, int64_t s, int idx,
                                                   float q, float dq, float df, float da, float db, float dc, const float (&dwx)[5],
                                                   const float (&dwy)[5], Add&& add) {
  const int* irec = reinterpret_cast<const int*>(rec);
  const int flags = irec[(int64_t)LARND_I_FLAGS * n + s];
  const float sl = rec[(int64_t)LARND_F_SL * n + s];
  const float sT = rec[(int64_t)LARND_F_ST * n + s];
  const float td = rec[(int64_t)LARND_F_TD * n + s];
  const float x0 = rec[(int64_t)LARND_F_X0 * n + s], y0 = rec[(int64_t)LARND_F_Y0 * n + s];
  const float recb = rec[(int64_t)LARND_F_REC * n + s];
  const float ft = rec[(int64_t)LARND_F_FT * n + s];
  const float xi = rec[(int64_t)LARND_F_XI * n + s];
  const float cos2 = rec[(int64_t)LARND_F_COS2 * n + s];
  // Lagrange weights (sim_jax.py:165-168)
  const float t0v = p.long_diff_template[idx - 1], t1v = p.long_diff_template[idx], t2v = p.long_diff_template[idx + 1];
  const float das = ((sl - t1v) + (sl - t2v)) / ((t0v - t1v) * (t0v - t2v));
  const float dbs = ((sl - t0v) + (sl - t2v)) / ((t1v - t0v) * (t1v - t2v));
  const float dcs = ((sl - t0v) + (sl - t1v)) / ((t2v - t0v) * (t2v - t1v));
  const float g_sl = da * das + db * dbs + dc * dcs;
  // diffusion weights W_k = 0.5 (E_{k+1} - E_k), E_k = erf((edge_k - x0)/(sqrt2 sT)) for k = 1..4
  float g_x0 = 0.f, g_y0 = 0.f, g_sT = 0.f;
  if (sT > 0.0f) {
    const float inv = 1.0f / (1.41421354f * sT);
    const float two_over_sqrt_pi = 1.12837917f;
#pragma unroll
    for (int k = 1; k < 5; ++k) {
      const float ux = (p.tran_bin_edges[k] - x0) * inv, uy = (p.tran_bin_edges[k] - y0) * inv;
      const float px_ = two_over_sqrt_pi * expf(-ux * ux), py_ = two_over_sqrt_pi * expf(-uy * uy);
      const float gEx = 0.5f * (dwx[k - 1] - dwx[k]);
      const float gEy = 0.5f * (dwy[k - 1] - dwy[k]);
      g_x0 += gEx * (-px_ * inv);
      g_y0 += gEy * (-py_ * inv);
      g_sT += gEx * (-px_ * ux / sT) + gEy * (-py_ * uy / sT);
    }
  }
  const float v = p.vdrift, tau = p.lifetime, ts = p.t_sampling;
  const float sgn_a = (flags & 2) ? 1.0f : -1.0f, sgn_c = (flags & 4) ? 1.0f : -1.0f;
  const float g_ft = df;
  const float g_td = dq * (-q / tau) + (td > 0.f ? (g_sl * sl + g_sT * sT) / (2.0f * td) : 0.f);
  const float g_v = g_td * (-td / v) + g_ft * (-ft / v) + g_sl * (-sl / v);
  add(LARND_P_SHIFT_Z, g_td * (-sgn_a / v) + g_ft * (-sgn_c / (v * ts)));
  add(LARND_P_LIFETIME, dq * q * td / (tau * tau));
  if (p.long_diff > 0.f) add(LARND_P_LONG_DIFF, g_sl * sl / (2.0f * p.long_diff));
  if (p.tran_diff > 0.f) add(LARND_P_TRAN_DIFF, g_sT * sT / (2.0f * p.tran_diff));
  add(LARND_P_SHIFT_X, -g_x0);
  add(LARND_P_SHIFT_Y, -g_y0);
  add(LARND_P_MEV_TO_ELECTRONS, dq * q / p.MeVToElectrons);
  const float g_rec = (recb != 0.0f) ? dq * q / recb : 0.0f;  // q is linear in the recombination factor
  float g_E = g_v * p.dvdrift_dEfield;
  if (p.recombination_mode == 2) {          // Birks: rec = Ab / (1 + xi), xi = kb dEdx / (E rho)
    const float dn = 1.0f + xi;
    add(LARND_P_AB, g_rec * recb / p.Ab);
    const float g_xi = g_rec * (-recb / dn);
    if (p.kb != 0.f) add(LARND_P_KB, g_xi * xi / p.kb);
    g_E += g_xi * (-xi / p.eField);
    add(LARND_P_LAR_DENSITY, g_xi * (-xi / p.lArDensity));
  } else if (recb > 0.0f) {                 // Box / Ellipsoid: rec = log(alpha + xi) / (xi [+1e-10])
    const float den = (p.recombination_mode == 3) ? xi + 1e-10f : xi;
    const float lg = logf(p.alpha + xi);
    add(LARND_P_ALPHA, g_rec / ((p.alpha + xi) * den));
    const float g_xi = g_rec * (1.0f / ((p.alpha + xi) * den) - lg / (den * den));
    add(LARND_P_BETA, g_xi * xi / p.beta);
    g_E += g_xi * (-xi / p.eField);
    add(LARND_P_LAR_DENSITY, g_xi * (-xi / p.lArDensity));
    if (p.recombination_mode == 3) {
      const float gg = 1.0f - cos2 + p.inv_R2 * cos2;   // b_phi = beta / sqrt(gg)
      add(LARND_P_R_PARAM, g_xi * xi * cos2 / (p.R_param * p.R_param * p.R_param * gg));
    }
  }
  add(LARND_P_EFIELD, g_E);
}
