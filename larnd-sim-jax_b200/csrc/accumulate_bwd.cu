// K3b + K1b: VJP of simulate_wfs (reference sim_jax.py:689-736) w.r.t. the fitted Params leaves, sm_100a.
//
// The reference obtains this by jax.grad through gather / scatter-add / erf / sqrt (optimize/fit_params.py:731).
// Here the backward pass is the *gather* form of K3: a warp owns a segment, walks its 25 + (2n+1)^2 target
// rows, reads the upstream gradient row window g[row, T0-1 .. T0+L] (coalesced, L1/L2-resident because
// consecutive segments hit the same rows) together with the same response rows as the forward pass, and
// accumulates lane-partial derivatives w.r.t. the per-segment continuous quantities
//     q, frac, (a, b, c), Wx[5], Wy[5]
// (everything else on the path is an integer index, hence has zero gradient).  One warp
// reduction per segment, then the closed-form chain rule through drift / quench / diffusion-weight math
// (drifting_jax.py:42-50, quenching_jax.py:18-35, detsim_jax.py:332-341, sim_jax.py:157-168,406-423) gives the
// LARND_NPARAMS parameter gradients; they are block-reduced, written as per-chunk partials and summed in
// double precision by a second tiny kernel (deterministic, no float atomics).
#include "larnd_common.cuh"

namespace {

constexpr int BWD_THREADS = 256;
constexpr int BWD_WARPS = BWD_THREADS / 32;
constexpr int S = LARND_CHUNK;

struct BwdArgs {
  const float* rec;
  int64_t n;
  const float* r0;
  const float* rm;
  const float* c0;
  const float* cm;
  int nt, L, Lp, ny_lut;
  int nticks;
  int nb, half2;
  int n_neigh, P;
  int nxp, nyp;
  RowLookup lk;
  const int32_t* counts;
  const float* g;
  int64_t g_stride;
  float* partials;
  int skip_garbage;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(BWD_THREADS)
k_lut_backward(const __grid_constant__ BwdArgs A, const __grid_constant__ larnd_params_t p) {
  __shared__ int s_rows[BWD_WARPS][25 + 15 * 15];
  __shared__ float s_grad[BWD_WARPS][16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float gacc[LARND_NPARAMS];  // lane 0 only
#pragma unroll
  for (int k = 0; k < LARND_NPARAMS; ++k) gacc[k] = 0.0f;
  const bool bad = A.counts[2] != 0;
  const int64_t s_base = (int64_t)blockIdx.x * S;
  const int ns = bad ? 0 : (int)min((int64_t)S, A.n - s_base);
  RowLookup lk = A.lk;
  lk.n_unique = A.counts[0];
  lk.n_neg = A.counts[1];
  const int nb = A.nb, L = A.L, nt = A.nt;
  const int n_units = 25 + A.P * A.P;
  const int sym = (LARND_NB_TRAN_BINS - 1) / 2;
  const int* irec = reinterpret_cast<const int*>(A.rec);
  const int64_t n = A.n;

  for (int t = wid; t < ns; t += BWD_WARPS) {
    const int64_t s = s_base + t;
    const float q = A.rec[(int64_t)LARND_F_Q * n + s];
    const float f = A.rec[(int64_t)LARND_F_FRAC * n + s];
    const int T0 = irec[(int64_t)LARND_I_T0 * n + s];
    const int bx = irec[(int64_t)LARND_I_BX * n + s], by = irec[(int64_t)LARND_I_BY * n + s];
    const int ep = irec[(int64_t)LARND_I_EP * n + s];
    const int idx = irec[(int64_t)LARND_I_IDX * n + s];
    const int flags = irec[(int64_t)LARND_I_FLAGS * n + s];
    if (!(flags & 1)) continue;  // outside every TPC: q == 0 and d q = 0 (mask factor)
    const float ca_ = A.rec[(int64_t)LARND_F_A * n + s], cb_ = A.rec[(int64_t)LARND_F_B * n + s],
                cc_ = A.rec[(int64_t)LARND_F_C * n + s];
    float wx[5], wy[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      wx[k] = A.rec[(int64_t)(LARND_F_WX0 + k) * n + s];
      wy[k] = A.rec[(int64_t)(LARND_F_WY0 + k) * n + s];
    }
    const int mpx = floordiv_i(bx, nb), mpy = floordiv_i(by, nb);
    const int bxm = bx - mpx * nb, bym = by - mpy * nb;
    // ---- target rows of all units (lane-parallel lookups) ---------------------------------------------
    __syncwarp();
    for (int u = lane; u < n_units; u += 32) {
      int row;
      if (u < 25) {
        const int bi = u / 5, bj = u % 5;
        const int px = floordiv_i(bx + bi - sym, nb), py = floordiv_i(by + bj - sym, nb);
        const int pid = pixel2id_dev(px, py, ep, A.nxp, A.nyp);
        row = lookup_row(lk, pid);  // absent -> dropped
        if (A.skip_garbage && pid < 0) row = -1;
      } else {
        const int v = u - 25;
        const int dx = v / A.P - A.n_neigh, dy = v % A.P - A.n_neigh;
        if (dx == 0 && dy == 0) {
          row = A.skip_garbage ? -1 : 0;
        } else {
          const int pid = pixel2id_dev(mpx + dx, mpy + dy, ep, A.nxp, A.nyp);
          row = lookup_row(lk, pid);
          const bool garbage = row < 0 || pid < 0;
          if (row < 0) row = 0;
          if (A.skip_garbage && garbage) row = -1;
        }
      }
      s_rows[wid][u] = row;
    }
    __syncwarp();
    const int ct = nt - L - T0;
    const int ct1 = min(ct + 1, nt - 1);
    const bool fast = (T0 >= 2) && (T0 + L <= A.nticks - 1);
    float dq = 0.f, df = 0.f, da = 0.f, db = 0.f, dc = 0.f;  // lane partials
    float dw = 0.f;                                           // lanes 0..4: dWx[i]; lanes 5..9: dWy[j]
    const float omf = 1.0f - f;
    for (int u = 0; u < n_units; ++u) {
      const int row = s_rows[wid][u];
      if (row < 0) continue;
      const float* grow = A.g + (int64_t)row * A.g_stride;
      if (u >= 25) {
        const int v = u - 25;
        const int dx = v / A.P - A.n_neigh, dy = v % A.P - A.n_neigh;
        const int ci = abs(2 * bxm - A.half2 - 2 * nb * dx) >> 1, cj = abs(2 * bym - A.half2 - 2 * nb * dy) >> 1;
        const int bin = ci * A.ny_lut + cj;
        const float* rowp = A.r0 + (int64_t)bin * A.Lp;
        const float* crow = A.c0 + (int64_t)bin * nt;
        const float Ca = __ldg(crow + ct), Cb = __ldg(crow + ct1), Cl = __ldg(crow + nt - L);
        const float D = Cl - (Ca * omf + Cb * f);
        const float dD = -(Cb - Ca);
        for (int x = -1 + lane; x <= L; x += 32) {
          const int col = T0 + x;
          float gw, gc;
          if (fast) { gw = gc = __ldg(grow + col); }
          else {
            const bool inb = col >= 1 && col <= A.nticks - 1;
            const float gv = inb ? __ldg(grow + col) : 0.0f;
            gw = (col >= 2) ? gv : 0.0f;
            gc = (col <= A.nticks - 2) ? gv : 0.0f;
          }
          const float v0 = __ldg(rowp + x + 2), v1 = __ldg(rowp + x + 1);
          const float cD = (x == -1) ? gc * f : ((x == 0) ? gc * omf : 0.0f);   // d/d(q D)
          const float cDs = (x == -1) ? gc : ((x == 0) ? -gc : 0.0f);           // coefficient of q*D in d/df
          dq = fmaf(gw, fmaf(f, v0, omf * v1), fmaf(cD, D, dq));
          df = fmaf(q, fmaf(gw, v0 - v1, fmaf(cDs, D, cD * dD)), df);
        }
      } else {
        const int bi = u / 5, bj = u % 5;
        const int bxx = bx + bi - sym, byy = by + bj - sym;
        const int px = floordiv_i(bxx, nb), py = floordiv_i(byy, nb);
        const int cix = abs(2 * (bxx - px * nb) - A.half2) >> 1, ciy = abs(2 * (byy - py * nb) - A.half2) >> 1;
        const int bin = cix * 5 + ciy;
        const float* ra = A.rm + (int64_t)((idx - 1) * 25 + bin) * A.Lp;
        const float* rb = A.rm + (int64_t)(idx * 25 + bin) * A.Lp;
        const float* rc = A.rm + (int64_t)((idx + 1) * 25 + bin) * A.Lp;
        const float* crow = A.cm + (int64_t)(idx * 25 + bin) * nt;
        const float Ca = __ldg(crow + ct), Cb = __ldg(crow + ct1), Cl = __ldg(crow + nt - L);
        const float D = Cl - (Ca * omf + Cb * f);
        const float dD = -(Cb - Ca);
        const float w = wx[bi] * wy[bj];
        const float qb = w * q;
        float P = 0.f, Sa = 0.f, Sb = 0.f, Sc = 0.f, Fd = 0.f;
        for (int x = -1 + lane; x <= L; x += 32) {
          const int col = T0 + x;
          float gw, gc;
          if (fast) { gw = gc = __ldg(grow + col); }
          else {
            const bool inb = col >= 1 && col <= A.nticks - 1;
            const float gv = inb ? __ldg(grow + col) : 0.0f;
            gw = (col >= 2) ? gv : 0.0f;
            gc = (col <= A.nticks - 2) ? gv : 0.0f;
          }
          const float a0 = __ldg(ra + x + 2), a1 = __ldg(ra + x + 1);
          const float b0 = __ldg(rb + x + 2), b1 = __ldg(rb + x + 1);
          const float c0v = __ldg(rc + x + 2), c1v = __ldg(rc + x + 1);
          const float ta = fmaf(f, a0, omf * a1), tb = fmaf(f, b0, omf * b1), tc = fmaf(f, c0v, omf * c1v);
          const float bl0 = fmaf(ca_, a0, fmaf(cb_, b0, cc_ * c0v));
          const float bl1 = fmaf(ca_, a1, fmaf(cb_, b1, cc_ * c1v));
          const float cD = (x == -1) ? gc * f : ((x == 0) ? gc * omf : 0.0f);
          const float cDs = (x == -1) ? gc : ((x == 0) ? -gc : 0.0f);
          Sa = fmaf(gw, ta, Sa);
          Sb = fmaf(gw, tb, Sb);
          Sc = fmaf(gw, tc, Sc);
          P = fmaf(gw, fmaf(ca_, ta, fmaf(cb_, tb, cc_ * tc)), fmaf(cD, D, P));
          Fd = fmaf(gw, bl0 - bl1, fmaf(cDs, D, fmaf(cD, dD, Fd)));
        }
        da = fmaf(qb, Sa, da);
        db = fmaf(qb, Sb, db);
        dc = fmaf(qb, Sc, dc);
        df = fmaf(qb, Fd, df);
        dq = fmaf(w, P, dq);
        const float Pt = warp_sum(P);
        if (lane == bi) dw = fmaf(wy[bj] * q, Pt, dw);
        if (lane == 5 + bj) dw = fmaf(wx[bi] * q, Pt, dw);
      }
    }
    dq = warp_sum(dq); df = warp_sum(df); da = warp_sum(da); db = warp_sum(db); dc = warp_sum(dc);
    float gwx[5], gwy[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      gwx[k] = __shfl_sync(0xffffffffu, dw, k);
      gwy[k] = __shfl_sync(0xffffffffu, dw, 5 + k);
    }
    if (lane == 0) {
      // ---- K1b: chain rule through the per-segment preparation ---------------------------------------
      const float sl = A.rec[(int64_t)LARND_F_SL * n + s];
      const float sT = A.rec[(int64_t)LARND_F_ST * n + s];
      const float td = A.rec[(int64_t)LARND_F_TD * n + s];
      const float x0 = A.rec[(int64_t)LARND_F_X0 * n + s], y0 = A.rec[(int64_t)LARND_F_Y0 * n + s];
      const float recb = A.rec[(int64_t)LARND_F_REC * n + s];
      const float ft = A.rec[(int64_t)LARND_F_FT * n + s];
      const float xi = A.rec[(int64_t)LARND_F_XI * n + s];
      const float cos2 = A.rec[(int64_t)LARND_F_COS2 * n + s];
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
          const float gEx = 0.5f * (gwx[k - 1] - gwx[k]), gEy = 0.5f * (gwy[k - 1] - gwy[k]);
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
      gacc[LARND_P_SHIFT_Z] += g_td * (-sgn_a / v) + g_ft * (-sgn_c / (v * ts));
      gacc[LARND_P_LIFETIME] += dq * q * td / (tau * tau);
      if (p.long_diff > 0.f) gacc[LARND_P_LONG_DIFF] += g_sl * sl / (2.0f * p.long_diff);
      if (p.tran_diff > 0.f) gacc[LARND_P_TRAN_DIFF] += g_sT * sT / (2.0f * p.tran_diff);
      gacc[LARND_P_SHIFT_X] += -g_x0;
      gacc[LARND_P_SHIFT_Y] += -g_y0;
      gacc[LARND_P_MEV_TO_ELECTRONS] += dq * q / p.MeVToElectrons;
      // recombination factor: q is linear in it
      float g_rec = (recb != 0.0f) ? dq * q / recb : 0.0f;
      float g_E = g_v * p.dvdrift_dEfield;
      if (p.recombination_mode == 2) {          // Birks: rec = Ab / (1 + xi), xi = kb dEdx / (E rho)
        const float dn = 1.0f + xi;
        gacc[LARND_P_AB] += g_rec * recb / p.Ab;
        const float g_xi = g_rec * (-recb / dn);
        if (p.kb != 0.f) gacc[LARND_P_KB] += g_xi * xi / p.kb;
        g_E += g_xi * (-xi / p.eField);
        gacc[LARND_P_LAR_DENSITY] += g_xi * (-xi / p.lArDensity);
      } else if (recb > 0.0f) {                 // Box / Ellipsoid: rec = log(alpha + xi) / (xi [+1e-10])
        const float den = (p.recombination_mode == 3) ? xi + 1e-10f : xi;
        const float lg = logf(p.alpha + xi);
        gacc[LARND_P_ALPHA] += g_rec / ((p.alpha + xi) * den);
        const float g_xi = g_rec * (1.0f / ((p.alpha + xi) * den) - lg / (den * den));
        gacc[LARND_P_BETA] += g_xi * xi / p.beta;
        g_E += g_xi * (-xi / p.eField);
        gacc[LARND_P_LAR_DENSITY] += g_xi * (-xi / p.lArDensity);
        if (p.recombination_mode == 3) {
          const float gg = 1.0f - cos2 + p.inv_R2 * cos2;   // b_phi = beta / sqrt(gg)
          gacc[LARND_P_R_PARAM] += g_xi * xi * cos2 / (p.R_param * p.R_param * p.R_param * gg);
        }
      }
      gacc[LARND_P_EFIELD] += g_E;
    }
  }
  // ---- block reduction -> per-chunk partials ----------------------------------------------------------
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < LARND_NPARAMS; ++k) s_grad[wid][k] = gacc[k];
  }
  __syncthreads();
  if (threadIdx.x < LARND_NPARAMS) {
    float v = 0.f;
    for (int w = 0; w < BWD_WARPS; ++w) v += s_grad[w][threadIdx.x];
    A.partials[(int64_t)blockIdx.x * 16 + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ partials, int64_t n_chunks,
                                                         float* __restrict__ grad) {
  __shared__ double sm[256];
  const int pidx = blockIdx.x;  // one block per parameter
  double acc = 0.0;
  for (int64_t c = threadIdx.x; c < n_chunks; c += 256) acc += (double)partials[c * 16 + pidx];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) grad[pidx] += (float)sm[0];
}

}  // namespace

int larnd_launch_accumulate_bwd(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                int32_t npix_capacity, int32_t flags, const float* g_wfs, int64_t g_stride,
                                float* grad_params, const int32_t* counts, cudaStream_t st) {
  if (n == 0) return LARND_OK;
  if (p.number_pix_neighbors > 7) { larnd_set_error("number_pix_neighbors > 7 unsupported"); return LARND_E_ARG; }
  BwdArgs A;
  A.rec = ws.rec; A.n = n;
  A.r0 = lut->r0; A.rm = lut->rm; A.c0 = lut->c0; A.cm = lut->cm;
  A.nt = lut->nt; A.L = lut->L; A.Lp = lut->Lp; A.ny_lut = lut->ny;
  A.nticks = p.n_ticks;
  A.nb = p.nb_sampling_bins_per_pixel;
  A.half2 = 2 * (A.nb / 2) - 1;
  A.n_neigh = p.number_pix_neighbors;
  A.P = 2 * A.n_neigh + 1;
  A.nxp = p.n_pixels_x; A.nyp = p.n_pixels_y;
  A.lk.bitmap = ws.bitmap; A.lk.wprefix = ws.wprefix; A.lk.n_words = ws.n_words; A.lk.pid_offset = ws.pid_offset;
  A.lk.n_unique = 0; A.lk.n_neg = 0; A.lk.npix = npix_capacity;
  A.counts = counts;
  A.g = g_wfs; A.g_stride = g_stride;
  A.partials = ws.partials;
  A.skip_garbage = flags & 1;
  const int64_t chunks = (n + S - 1) / S;
  prof_begin(2, st);
  k_lut_backward<<<(unsigned)chunks, BWD_THREADS, 0, st>>>(A, p);
  prof_end(2, st);
  LARND_LAUNCH_CHECK("k_lut_backward");
  k_reduce_partials<<<LARND_NPARAMS, 256, 0, st>>>(ws.partials, chunks, grad_params);
  LARND_LAUNCH_CHECK("k_reduce_partials");
  return LARND_OK;
}
