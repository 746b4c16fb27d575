// Stream-form operators, sm_100a: the reference functions whose ARGUMENTS are the materialised per-segment streams.
//
// The fused path (prepare.cu -> accumulate*.cu) never builds those streams; these kernels exist so that code written
// against the reference's stage-by-stage API keeps working and composes with the fused kernels:
//   larnd_tracks_stage             <- shift_tracks (sim_jax.py:109-119), quench (quenching_jax.py:38-75),
//                                     drift (drifting_jax.py:19-58): (N, ncols) tracks in -> updated tracks out
//   larnd_signals_stream_forward   <- simulate_signals (sim_jax.py:142-286) with the reference's own argument list
//   larnd_signals_stream_backward  <- its VJP w.r.t. nelectrons, t0_after_diff, long_diff, nelectrons_neigh, t0_neigh
// One warp per stream entry (a (segment, diffusion bin) pair or a (segment, neighbour pixel) pair), lanes <-> the L
// samples of the response window, values formed in the reference's own order of operations.
#include "larnd_common.cuh"
#include "segment_physics.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- tracks stages
constexpr int ST_THREADS = 128;

__global__ void __launch_bounds__(ST_THREADS)
k_tracks_stage(const float* __restrict__ tracks, int64_t n, const __grid_constant__ larnd_columns_t cols,
               const __grid_constant__ larnd_track_columns_t oc, const __grid_constant__ larnd_params_t p, int stages,
               float* __restrict__ out) {
  extern __shared__ float srow[];
  const int ncols = cols.ncols;
  const int stride = ncols | 1;
  const int64_t base = (int64_t)blockIdx.x * ST_THREADS;
  const int rows_here = (int)min((int64_t)ST_THREADS, n - base);
  const int total = rows_here * ncols;
  const float* src = tracks + base * ncols;
  for (int i = threadIdx.x; i < total; i += ST_THREADS) {
    int r = i / ncols, c = i - r * ncols;
    srow[r * stride + c] = __ldg(src + i);
  }
  __syncthreads();
  if ((int)threadIdx.x < rows_here) {
    float* tr = srow + threadIdx.x * stride;
    if (stages & 1) {  // shift_tracks: the nine coordinate columns
      tr[cols.x] = fsub(tr[cols.x], p.shift_x);
      tr[oc.x_start] = fsub(tr[oc.x_start], p.shift_x);
      tr[oc.x_end] = fsub(tr[oc.x_end], p.shift_x);
      tr[cols.y] = fsub(tr[cols.y], p.shift_y);
      tr[oc.y_start] = fsub(tr[oc.y_start], p.shift_y);
      tr[oc.y_end] = fsub(tr[oc.y_end], p.shift_y);
      tr[cols.z] = fsub(tr[cols.z], p.shift_z);
      tr[cols.z_start] = fsub(tr[cols.z_start], p.shift_z);
      tr[cols.z_end] = fsub(tr[cols.z_end], p.shift_z);
    }
    if (stages & 2) {  // quench
      const float dEdx = tr[cols.dEdx], dE = tr[cols.dE];
      float recomb;
      if (p.recombination_mode == 2) {
        recomb = fdiv(p.Ab, fadd(1.0f, fdiv(fmul(p.kb, dEdx), p.efield_rho)));
      } else if (p.recombination_mode == 1) {
        const float csi = fdiv(fmul(p.beta, dEdx), p.efield_rho);
        recomb = fmaxf(0.0f, fdiv(logf(fadd(p.alpha, csi)), csi));
      } else {
        const float cosphi = fdiv(fabsf(fsub(tr[cols.z_end], tr[cols.z_start])), fadd(tr[cols.dx], 1e-10f));
        const float c2 = fmul(cosphi, cosphi);
        const float bphi = fdiv(p.beta, __fsqrt_rn(fadd(fsub(1.0f, c2), fmul(p.inv_R2, c2))));
        const float csi = fdiv(fmul(bphi, dEdx), p.efield_rho);
        recomb = fmaxf(0.0f, fdiv(logf(fadd(p.alpha, csi)), fadd(csi, 1e-10f)));
      }
      tr[oc.n_electrons] = fmul(fmul(recomb, dE), p.MeVToElectrons);
    }
    if (stages & 4) {  // drift
      const float x = tr[cols.x], y = tr[cols.y], z = tr[cols.z];
      int plane = 0;
      bool inside = false;
      for (int k = p.n_tpc - 1; k >= 0; --k) {
        const float za = p.tpc_borders[k][2][0], zc = p.tpc_borders[k][2][1];
        const float zmin = fminf(fsub(zc, p.size_margin), fsub(za, p.size_margin));
        const float zmax = fmaxf(fadd(zc, p.size_margin), fadd(za, p.size_margin));
        const bool c = x >= fsub(p.tpc_borders[k][0][0], p.size_margin) && x <= fadd(p.tpc_borders[k][0][1], p.size_margin) &&
                       y >= fsub(p.tpc_borders[k][1][0], p.size_margin) && y <= fadd(p.tpc_borders[k][1][1], p.size_margin) &&
                       z >= zmin && z <= zmax;
        if (c) { plane = k; inside = true; }
      }
      const float m = inside ? 1.0f : 0.0f;
      const float z_anode = p.tpc_borders[plane][2][0];
      const float td = fdiv(fadd(fabsf(fsub(z, z_anode)), 1e-6f), p.vdrift);
      const float zs = tr[cols.z_start], ze = tr[cols.z_end];
      const float d_lo = fabsf(fsub(fminf(zs, ze), z_anode)), d_hi = fabsf(fsub(fmaxf(zs, ze), z_anode));
      const float t0 = tr[cols.t0];
      tr[oc.pixel_plane] = (float)plane;
      tr[oc.n_electrons] = fmul(fmul(tr[oc.n_electrons], expf(-fdiv(td, p.lifetime))), m);
      tr[oc.long_diff] = __fsqrt_rn(fmul(fmul(td, 2.0f), p.long_diff));
      tr[oc.tran_diff] = __fsqrt_rn(fmul(fmul(td, 2.0f), p.tran_diff));
      tr[oc.t] = fadd(fadd(tr[oc.t], fmul(td, m)), t0);
      tr[oc.t_start] = fadd(fadd(tr[oc.t_start], fmul(fdiv(fminf(d_lo, d_hi), p.vdrift), m)), t0);
      tr[oc.t_end] = fadd(fadd(tr[oc.t_end], fmul(fdiv(fmaxf(d_lo, d_hi), p.vdrift), m)), t0);
    }
  }
  __syncthreads();
  float* dst = out + base * ncols;
  for (int i = threadIdx.x; i < total; i += ST_THREADS) {
    int r = i / ncols, c = i - r * ncols;
    dst[i] = srow[r * stride + c];
  }
}

// ---------------------------------------------------------------------------------------------- simulate_signals
struct StreamArgs {
  const int32_t* unique_pixels; int npix;
  const int32_t* pixels; const float* t0; const float* q; const float* ld; const int32_t* cidx; int64_t n_main;
  const float* qn; const int32_t* rown; const float* t0n; const int32_t* cidxn; int64_t n_neigh; int P2;
  larnd_lut lut;
  float* wfs;
  int32_t* status;   // bit0: main response index outside the 5x5 collecting bins, bit1: neighbour index outside the LUT
  const float* g; int64_t g_stride;
  float *g_q, *g_t0, *g_ld, *g_qn, *g_t0n;
};

struct Entry {
  int row;               // waveform row, < 0: dropped (jax.ops.segment_sum drops ids outside [0, Npix*Nticks))
  float q, frac, ld;
  int ct, st0;
  float a, b, c, x0, x1, x2;
  const float *Ri, *Ra, *Rc;  // compact rows of template idx / idx-1 / idx+1 (sample k at [k]); neighbours: Ri only
  const float* C;             // running sum of the idx row over the full time axis
  bool main;
  int64_t seg;
};

constexpr int SG_WARPS = 8;

__device__ __forceinline__ bool load_entry(const StreamArgs& A, const larnd_params_t& p, int64_t e, Entry& E) {
  const larnd_lut& lut = A.lut;
  const int Nt = lut.nt, L = lut.L;
  float t0;
  E.main = e < A.n_main;
  E.row = -1;
  if (E.main) {
    const int pix = __ldg(A.pixels + e);
    int lo = 0, hi = A.npix;  // searchsorted(unique_pixels, pixels), side='left'
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(A.unique_pixels + mid) < pix) lo = mid + 1; else hi = mid;
    }
    if (lo >= A.npix || __ldg(A.unique_pixels + lo) != pix) return false;  // pix_renum = -1: every flat id negative
    E.row = lo;
    E.q = __ldg(A.q + e);
    E.ld = __ldg(A.ld + e);
    t0 = __ldg(A.t0 + e);
    const int ci = __ldg(A.cidx + 2 * e), cj = __ldg(A.cidx + 2 * e + 1);
    if (ci < 0 || cj < 0 || ci >= 5 || cj >= 5) { atomicOr(A.status, 1); E.row = -1; return false; }
    int l2 = 0, h2 = p.n_templates;
    while (l2 < h2) {
      const int mid = (l2 + h2) >> 1;
      if (p.long_diff_template[mid] < E.ld) l2 = mid + 1; else h2 = mid;
    }
    const int idx = max(1, min(l2, p.n_templates - 2));
    if (idx + 1 >= lut.ntpl) { atomicOr(A.status, 4); E.row = -1; return false; }  // truncated bank: row missing (status bit2)
    E.x0 = p.long_diff_template[idx - 1]; E.x1 = p.long_diff_template[idx]; E.x2 = p.long_diff_template[idx + 1];
    E.a = (E.ld - E.x1) * (E.ld - E.x2) / ((E.x0 - E.x1) * (E.x0 - E.x2));
    E.b = (E.ld - E.x0) * (E.ld - E.x2) / ((E.x1 - E.x0) * (E.x1 - E.x2));
    E.c = (E.ld - E.x0) * (E.ld - E.x1) / ((E.x2 - E.x0) * (E.x2 - E.x1));
    const int64_t r = (int64_t)idx * 25 + ci * 5 + cj;
    E.Ri = lut.rm + r * lut.Lp + 2;
    E.Ra = lut.rm + (r - 25) * lut.Lp + 2;
    E.Rc = lut.rm + (r + 25) * lut.Lp + 2;
    E.C = lut.cm + r * Nt;
  } else {
    const int64_t j = e - A.n_main;
    E.seg = j / A.P2;
    const int row = __ldg(A.rown + j);
    if (row < 0 || row >= A.npix) return false;
    E.row = row;
    E.q = __ldg(A.qn + E.seg);
    t0 = __ldg(A.t0n + E.seg);
    const int ci = __ldg(A.cidxn + 2 * j), cj = __ldg(A.cidxn + 2 * j + 1);
    if (ci < 0 || cj < 0 || ci >= lut.nx || cj >= lut.ny) { atomicOr(A.status, 2); E.row = -1; return false; }
    const int64_t r = (int64_t)ci * lut.ny + cj;
    E.Ri = lut.r0 + r * lut.Lp + 2;
    E.Ra = E.Rc = E.Ri;
    E.a = E.c = 0.f; E.b = 1.f;
    E.C = lut.c0 + r * Nt;
  }
  const float ft = __fdiv_rn(t0, p.t_sampling);
  E.ct = max(0, min((int)floorf(ft), Nt - 1));
  E.frac = __fsub_rn(ft, (float)E.ct);
  E.st0 = Nt - L - E.ct;
  return true;
}

__device__ __forceinline__ int tick_rule(int tt, int nticks) { return (tt <= 0 || tt >= nticks - 1) ? 0 : tt + 1; }
__device__ __forceinline__ int start_rule(int st, int nticks) { return (st <= 0 || st >= nticks - 1) ? 0 : st; }

__global__ void __launch_bounds__(SG_WARPS * 32)
k_signals_stream(const __grid_constant__ StreamArgs A, const __grid_constant__ larnd_params_t p) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * SG_WARPS + (threadIdx.x >> 5);
  if (e >= A.n_main + A.n_neigh) return;
  Entry E;
  if (!load_entry(A, p, e, E)) return;
  const int nticks = p.n_ticks, L = A.lut.L, Nt = A.lut.nt;
  float* base = A.wfs + (int64_t)E.row * nticks;
  const float f = E.frac, omf = 1.0f - f;
  float garbage = 0.f;
  for (int k = lane; k < L; k += 32) {
    float v;
    if (E.main) v = (__ldg(E.Ri + k) * E.b + __ldg(E.Ra + k) * E.a + __ldg(E.Rc + k) * E.c) * E.q;
    else v = __ldg(E.Ri + k) * E.q;
    const int t0k = tick_rule(E.st0 + k, nticks), t1k = tick_rule(E.st0 - 1 + k, nticks);
    if (t0k == 0) garbage += v * omf; else atomicAdd(base + t0k, v * omf);
    if (t1k == 0) garbage += v * f; else atomicAdd(base + t1k, v * f);
  }
  if (lane == 0) {  // boundary correction (sim_jax.py:228-261): template idx only, no +1 shift of the target tick
    const float c0 = __ldg(E.C + E.ct), c1 = __ldg(E.C + min(E.ct + 1, Nt - 1));
    const float interp = c0 * omf + c1 * f;
    const float d = (__ldg(E.C + Nt - L) - interp) * E.q;
    const int s0 = start_rule(E.st0, nticks), s1 = start_rule(E.st0 - 1, nticks);
    if (s0 == 0) garbage += d * omf; else atomicAdd(base + s0, d * omf);
    if (s1 == 0) garbage += d * f; else atomicAdd(base + s1, d * f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) garbage += __shfl_xor_sync(0xffffffffu, garbage, o);
  if (lane == 0 && garbage != 0.f) atomicAdd(base, garbage);
}

__global__ void __launch_bounds__(SG_WARPS * 32)
k_signals_stream_bwd(const __grid_constant__ StreamArgs A, const __grid_constant__ larnd_params_t p) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * SG_WARPS + (threadIdx.x >> 5);
  if (e >= A.n_main + A.n_neigh) return;
  Entry E;
  if (!load_entry(A, p, e, E)) return;   // dropped entries: zero gradient (outputs are zero-initialised)
  const int nticks = p.n_ticks, L = A.lut.L, Nt = A.lut.nt;
  const float* grow = A.g + (int64_t)E.row * A.g_stride;
  // S[1], S[4]: <g, R_idx> over window 0 / 1;  S[0], S[3]: <g, R_{idx-1} - R_idx>;  S[2], S[5]: <g, R_{idx+1} - R_idx>.
  // The three templates are nearly equal and da + db + dc = 0, so d/d(long_diff) is formed from the row DIFFERENCES
  // (a + b + c = 1: blend = R_idx + a (R_{idx-1} - R_idx) + c (R_{idx+1} - R_idx)); summing the rows separately would cancel
  // five digits in float32.
  float S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < L; k += 32) {
    const float g0 = __ldg(grow + tick_rule(E.st0 + k, nticks)), g1 = __ldg(grow + tick_rule(E.st0 - 1 + k, nticks));
    const float ri = __ldg(E.Ri + k);
    S[1] = fmaf(g0, ri, S[1]); S[4] = fmaf(g1, ri, S[4]);
    if (E.main) {
      const float ra = __ldg(E.Ra + k) - ri, rc = __ldg(E.Rc + k) - ri;
      S[0] = fmaf(g0, ra, S[0]); S[3] = fmaf(g1, ra, S[3]);
      S[2] = fmaf(g0, rc, S[2]); S[5] = fmaf(g1, rc, S[5]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < 6; ++i) S[i] += __shfl_xor_sync(0xffffffffu, S[i], o);
  if (lane != 0) return;
  const float f = E.frac, omf = 1.0f - f;
  const float B0 = S[1] + E.a * S[0] + E.c * S[2], B1 = S[4] + E.a * S[3] + E.c * S[5];
  const float c0 = __ldg(E.C + E.ct), c1 = __ldg(E.C + min(E.ct + 1, Nt - 1));
  const float D = __ldg(E.C + Nt - L) - (c0 * omf + c1 * f);
  const float gc0 = __ldg(grow + start_rule(E.st0, nticks)), gc1 = __ldg(grow + start_rule(E.st0 - 1, nticks));
  const float W = omf * gc0 + f * gc1;
  const float dq = omf * B0 + f * B1 + W * D;
  const float df = E.q * ((B1 - B0) + (gc1 - gc0) * D - W * (c1 - c0));
  const float dt0 = df / p.t_sampling;
  if (E.main) {
    const float da = ((E.ld - E.x1) + (E.ld - E.x2)) / ((E.x0 - E.x1) * (E.x0 - E.x2));
    const float dc = ((E.ld - E.x0) + (E.ld - E.x1)) / ((E.x2 - E.x0) * (E.x2 - E.x1));
    A.g_q[e] = dq;
    A.g_t0[e] = dt0;
    A.g_ld[e] = E.q * ((omf * S[0] + f * S[3]) * da + (omf * S[2] + f * S[5]) * dc);
  } else {
    atomicAdd(A.g_qn + E.seg, dq);
    atomicAdd(A.g_t0n + E.seg, dt0);
  }
}

// ---------------------------------------------------------------------------------------------- legacy entry points
// simulate_signals_new (sim_jax.py:456-617) and accumulate_signals (detsim_jax.py:157-205): the pre-"sub-tick split" form
// the reference keeps beside simulate_signals.  Differences, reproduced as written:
//   * the tick is (t0 / t_sampling).astype(int): truncation, no clip, no fractional split — one placement per sample;
//   * main pixels are renumbered with a bare searchsorted (no equality mask): a pixel id absent from unique_pixels lands on
//     the row of the next larger id, id > every entry -> row Npix -> flat index out of range -> dropped by .at[].add;
//   * the boundary correction of the MAIN entries reads response_cum.take(base + ...) WITHOUT the template offset, i.e.
//     the running sum of template 0 (sim_jax.py:593-596), also when the values were blended from templates idx-1..idx+1;
//   * flat indices go through .at[]: negative ones wrap once by Npix*Nticks, the rest out of range is dropped.
// One warp per entry; wfs is accumulated INTO (accumulate_signals adds onto its argument).
struct LegacyArgs {
  const int32_t* unique_pixels; int npix;
  const int32_t* pixels; const float* t0; const float* q; const float* ld; const int32_t* cidx; int64_t n_main;
  const float* qe; const int32_t* rowe; const int32_t* cte; const int32_t* cidxe; int64_t n_entries;
  larnd_lut lut;
  float* wfs;
  int32_t* status;   // bit0 / bit1 / bit2 as above, bit3: tick outside [0, Nt) (the reference would read a neighbouring LUT row)
};

__device__ __forceinline__ void legacy_add(float* wfs, int64_t size, int64_t flat, float v) {
  if (flat < 0) flat += size;
  if (flat >= 0 && flat < size) atomicAdd(wfs + flat, v);
}

__global__ void __launch_bounds__(SG_WARPS * 32)
k_signals_legacy(const __grid_constant__ LegacyArgs A, const __grid_constant__ larnd_params_t p) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * SG_WARPS + (threadIdx.x >> 5);
  if (e >= A.n_main + A.n_entries) return;
  const larnd_lut& lut = A.lut;
  const int Nt = lut.nt, L = lut.L, nticks = p.n_ticks;
  const int64_t size = (int64_t)A.npix * nticks;
  const bool main = e < A.n_main;
  int row, ct;
  float q, a = 0.f, b = 1.f, c = 0.f;
  const float *Ri, *Ra, *Rc, *C0;
  if (main) {
    const int pix = __ldg(A.pixels + e);
    int lo = 0, hi = A.npix;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(A.unique_pixels + mid) < pix) lo = mid + 1; else hi = mid;
    }
    row = lo;                                                    // no equality mask (sim_jax.py:533)
    q = __ldg(A.q + e);
    const float ld = __ldg(A.ld + e);
    ct = (int)__fdiv_rn(__ldg(A.t0 + e), p.t_sampling);         // astype(int): truncation
    const int ci = __ldg(A.cidx + 2 * e), cj = __ldg(A.cidx + 2 * e + 1);
    if (ci < 0 || cj < 0 || ci >= 5 || cj >= 5) { if (lane == 0) atomicOr(A.status, 1); return; }
    int l2 = 0, h2 = p.n_templates;
    while (l2 < h2) {
      const int mid = (l2 + h2) >> 1;
      if (p.long_diff_template[mid] < ld) l2 = mid + 1; else h2 = mid;
    }
    const int idx = max(1, min(l2, p.n_templates - 2));
    if (idx + 1 >= lut.ntpl) { if (lane == 0) atomicOr(A.status, 4); return; }
    const float x0 = p.long_diff_template[idx - 1], x1 = p.long_diff_template[idx], x2 = p.long_diff_template[idx + 1];
    a = (ld - x1) * (ld - x2) / ((x0 - x1) * (x0 - x2));
    b = (ld - x0) * (ld - x2) / ((x1 - x0) * (x1 - x2));
    c = (ld - x0) * (ld - x1) / ((x2 - x0) * (x2 - x1));
    const int64_t r = (int64_t)idx * 25 + ci * 5 + cj;
    Ri = lut.rm + r * lut.Lp + 2;
    Ra = lut.rm + (r - 25) * lut.Lp + 2;
    Rc = lut.rm + (r + 25) * lut.Lp + 2;
    C0 = lut.c0 + ((int64_t)ci * lut.ny + cj) * Nt;              // template 0 (no template offset in the reference)
  } else {
    const int64_t j = e - A.n_main;
    row = __ldg(A.rowe + j);
    q = __ldg(A.qe + j);
    ct = __ldg(A.cte + j);
    const int ci = __ldg(A.cidxe + 2 * j), cj = __ldg(A.cidxe + 2 * j + 1);
    if (ci < 0 || cj < 0 || ci >= lut.nx || cj >= lut.ny) { if (lane == 0) atomicOr(A.status, 2); return; }
    const int64_t r = (int64_t)ci * lut.ny + cj;
    Ri = Ra = Rc = lut.r0 + r * lut.Lp + 2;
    C0 = lut.c0 + r * Nt;
  }
  const int st0 = Nt - L - ct;
  const int64_t base = (int64_t)row * nticks;
  float garbage = 0.f;   // column 0 of this row collects every out-of-window sample
  for (int k = lane; k < L; k += 32) {
    // three separate .at[].add in the reference (b, a, c order); one rounded sum here
    const float v = main ? (__ldg(Ri + k) * q * b + __ldg(Ra + k) * q * a + __ldg(Rc + k) * q * c) : __ldg(Ri + k) * q;
    const int tk = tick_rule(st0 + k, nticks);
    if (tk == 0) garbage += v; else legacy_add(A.wfs, size, base + tk, v);
  }
  if (lane == 0) {
    int ctc = ct;
    if (ct < 0 || ct >= Nt) { atomicOr(A.status, 8); ctc = max(0, min(ct, Nt - 1)); }
    const float d = (__ldg(C0 + Nt - L) - __ldg(C0 + ctc)) * q;
    const int s0 = start_rule(st0, nticks);
    if (s0 == 0) garbage += d; else legacy_add(A.wfs, size, base + s0, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) garbage += __shfl_xor_sync(0xffffffffu, garbage, o);
  if (lane == 0 && garbage != 0.f) legacy_add(A.wfs, size, base, garbage);
}

int stream_check(const larnd_params_t* p, const larnd_lut* lut, int64_t n_main, int64_t n_seg) {
  if (!p || !lut) { larnd_set_error("larnd_signals_stream: null argument"); return LARND_E_ARG; }
  if (n_main < 0 || n_seg < 0) { larnd_set_error("larnd_signals_stream: negative size"); return LARND_E_ARG; }
  if (lut->L != p->signal_length) { larnd_set_error("LUT tables were built for signal_length %d, params say %d", lut->L, p->signal_length); return LARND_E_ARG; }
  if (lut->ntpl > p->n_templates || lut->ntpl < 3) { larnd_set_error("template count mismatch (bank %d vs long_diff_template %d)", lut->ntpl, p->n_templates); return LARND_E_ARG; }
  return LARND_OK;
}

}  // namespace

extern "C" int larnd_tracks_stage(const float* tracks_d, int64_t n, const larnd_columns_t* cols, const larnd_track_columns_t* oc,
                                  const larnd_params_t* p, int32_t stages, float* out_d, void* stream) {
  if (!cols || !oc || !p || (n > 0 && (!tracks_d || !out_d)) || n < 0) { larnd_set_error("larnd_tracks_stage: bad argument"); return LARND_E_ARG; }
  if (p->n_tpc < 1 || p->n_tpc > LARND_MAX_TPC) { larnd_set_error("n_tpc unsupported"); return LARND_E_ARG; }
  if (n == 0) return LARND_OK;
  const size_t smem = (size_t)ST_THREADS * (cols->ncols | 1) * sizeof(float);
  k_tracks_stage<<<(unsigned)((n + ST_THREADS - 1) / ST_THREADS), ST_THREADS, smem, (cudaStream_t)stream>>>(
      tracks_d, n, *cols, *oc, *p, stages, out_d);
  LARND_LAUNCH_CHECK("k_tracks_stage");
  return LARND_OK;
}

static void fill_stream_args(StreamArgs& A, const int32_t* unique_pixels_d, int32_t npix, const int32_t* pixels_d, const float* t0_d,
                             const float* nelectrons_d, const float* long_diff_d, const int32_t* currents_idx_d, int64_t n_main,
                             const float* nelectrons_neigh_d, const int32_t* pix_renumbering_neigh_d, const float* t0_neigh_d,
                             const int32_t* currents_idx_neigh_d, int64_t n_seg, int P2, const larnd_lut* lut) {
  A.unique_pixels = unique_pixels_d; A.npix = npix;
  A.pixels = pixels_d; A.t0 = t0_d; A.q = nelectrons_d; A.ld = long_diff_d; A.cidx = currents_idx_d; A.n_main = n_main;
  A.qn = nelectrons_neigh_d; A.rown = pix_renumbering_neigh_d; A.t0n = t0_neigh_d; A.cidxn = currents_idx_neigh_d;
  A.n_neigh = n_seg * P2; A.P2 = P2; A.lut = *lut;
  A.wfs = nullptr; A.status = nullptr; A.g = nullptr; A.g_stride = 0;
  A.g_q = A.g_t0 = A.g_ld = A.g_qn = A.g_t0n = nullptr;
}

extern "C" int larnd_signals_stream_forward(const int32_t* unique_pixels_d, int32_t npix, const int32_t* pixels_d,
                                            const float* t0_after_diff_d, const float* nelectrons_d, const float* long_diff_d,
                                            const int32_t* currents_idx_d, int64_t n_main, const float* nelectrons_neigh_d,
                                            const int32_t* pix_renumbering_neigh_d, const float* t0_neigh_d,
                                            const int32_t* currents_idx_neigh_d, int64_t n_segments, const larnd_params_t* p,
                                            const larnd_lut_t* lut, float* wfs_d, int32_t* status_d, void* stream) {
  int rc = stream_check(p, lut, n_main, n_segments);
  if (rc) return rc;
  if (!unique_pixels_d || npix < 1 || !wfs_d || !status_d) { larnd_set_error("larnd_signals_stream_forward: bad argument"); return LARND_E_ARG; }
  const int P = 2 * p->number_pix_neighbors + 1;
  cudaStream_t st = (cudaStream_t)stream;
  LARND_CUDA(cudaMemsetAsync(wfs_d, 0, (size_t)npix * p->n_ticks * sizeof(float), st));
  LARND_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
  StreamArgs A;
  fill_stream_args(A, unique_pixels_d, npix, pixels_d, t0_after_diff_d, nelectrons_d, long_diff_d, currents_idx_d, n_main,
                   nelectrons_neigh_d, pix_renumbering_neigh_d, t0_neigh_d, currents_idx_neigh_d, n_segments, P * P, lut);
  A.wfs = wfs_d; A.status = status_d;
  const int64_t entries = A.n_main + A.n_neigh;
  if (entries == 0) return LARND_OK;
  k_signals_stream<<<(unsigned)((entries + SG_WARPS - 1) / SG_WARPS), SG_WARPS * 32, 0, st>>>(A, *p);
  LARND_LAUNCH_CHECK("k_signals_stream");
  return LARND_OK;
}

extern "C" int larnd_signals_stream_backward(const int32_t* unique_pixels_d, int32_t npix, const int32_t* pixels_d,
                                             const float* t0_after_diff_d, const float* nelectrons_d, const float* long_diff_d,
                                             const int32_t* currents_idx_d, int64_t n_main, const float* nelectrons_neigh_d,
                                             const int32_t* pix_renumbering_neigh_d, const float* t0_neigh_d,
                                             const int32_t* currents_idx_neigh_d, int64_t n_segments, const larnd_params_t* p,
                                             const larnd_lut_t* lut, const float* g_wfs_d, int64_t g_row_stride,
                                             float* g_nelectrons_d, float* g_t0_after_diff_d, float* g_long_diff_d,
                                             float* g_nelectrons_neigh_d, float* g_t0_neigh_d, int32_t* status_d, void* stream) {
  int rc = stream_check(p, lut, n_main, n_segments);
  if (rc) return rc;
  if (!unique_pixels_d || npix < 1 || !g_wfs_d || !status_d) { larnd_set_error("larnd_signals_stream_backward: bad argument"); return LARND_E_ARG; }
  const int P = 2 * p->number_pix_neighbors + 1;
  cudaStream_t st = (cudaStream_t)stream;
  StreamArgs A;
  fill_stream_args(A, unique_pixels_d, npix, pixels_d, t0_after_diff_d, nelectrons_d, long_diff_d, currents_idx_d, n_main,
                   nelectrons_neigh_d, pix_renumbering_neigh_d, t0_neigh_d, currents_idx_neigh_d, n_segments, P * P, lut);
  A.status = status_d; A.g = g_wfs_d; A.g_stride = g_row_stride;
  A.g_q = g_nelectrons_d; A.g_t0 = g_t0_after_diff_d; A.g_ld = g_long_diff_d; A.g_qn = g_nelectrons_neigh_d; A.g_t0n = g_t0_neigh_d;
  LARND_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
  if (n_main > 0) {
    if (!g_nelectrons_d || !g_t0_after_diff_d || !g_long_diff_d) { larnd_set_error("larnd_signals_stream_backward: null main gradient"); return LARND_E_ARG; }
    LARND_CUDA(cudaMemsetAsync(g_nelectrons_d, 0, n_main * sizeof(float), st));
    LARND_CUDA(cudaMemsetAsync(g_t0_after_diff_d, 0, n_main * sizeof(float), st));
    LARND_CUDA(cudaMemsetAsync(g_long_diff_d, 0, n_main * sizeof(float), st));
  }
  if (n_segments > 0) {
    if (!g_nelectrons_neigh_d || !g_t0_neigh_d) { larnd_set_error("larnd_signals_stream_backward: null neighbour gradient"); return LARND_E_ARG; }
    LARND_CUDA(cudaMemsetAsync(g_nelectrons_neigh_d, 0, n_segments * sizeof(float), st));
    LARND_CUDA(cudaMemsetAsync(g_t0_neigh_d, 0, n_segments * sizeof(float), st));
  }
  const int64_t entries = A.n_main + A.n_neigh;
  if (entries == 0) return LARND_OK;
  k_signals_stream_bwd<<<(unsigned)((entries + SG_WARPS - 1) / SG_WARPS), SG_WARPS * 32, 0, st>>>(A, *p);
  LARND_LAUNCH_CHECK("k_signals_stream_bwd");
  return LARND_OK;
}

extern "C" int larnd_signals_legacy_forward(const int32_t* unique_pixels_d, int32_t npix, const int32_t* pixels_d,
                                            const float* t0_after_diff_d, const float* nelectrons_d, const float* long_diff_d,
                                            const int32_t* currents_idx_d, int64_t n_main, const float* charge_entries_d,
                                            const int32_t* pix_id_entries_d, const int32_t* cathode_ticks_entries_d,
                                            const int32_t* currents_idx_entries_d, int64_t n_entries, const larnd_params_t* p,
                                            const larnd_lut_t* lut, float* wfs_d, int32_t* status_d, void* stream) {
  int rc = stream_check(p, lut, n_main, n_entries);
  if (rc) return rc;
  if (npix < 1 || !wfs_d || !status_d || (n_main > 0 && (!unique_pixels_d || !pixels_d || !t0_after_diff_d || !nelectrons_d || !long_diff_d || !currents_idx_d)) ||
      (n_entries > 0 && (!charge_entries_d || !pix_id_entries_d || !cathode_ticks_entries_d || !currents_idx_entries_d))) {
    larnd_set_error("larnd_signals_legacy_forward: bad argument");
    return LARND_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LARND_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
  LegacyArgs A{unique_pixels_d, npix, pixels_d, t0_after_diff_d, nelectrons_d, long_diff_d, currents_idx_d, n_main,
               charge_entries_d, pix_id_entries_d, cathode_ticks_entries_d, currents_idx_entries_d, n_entries, *lut, wfs_d, status_d};
  const int64_t entries = n_main + n_entries;
  if (entries == 0) return LARND_OK;
  k_signals_legacy<<<(unsigned)((entries + SG_WARPS - 1) / SG_WARPS), SG_WARPS * 32, 0, st>>>(A, *p);
  LARND_LAUNCH_CHECK("k_signals_legacy");
  return LARND_OK;
}

// current_lut (detsim_jax.py:642-660): t0 = response_full_drift_t - t, response bin of |electron - pixel centre|.
namespace {
__global__ void k_current_lut(const float* __restrict__ el, int64_t n, int ncols, int cx, int cy, int ctm, const float* __restrict__ pc,
                              float full_drift_t, float bin_size, int nx, int ny, float* __restrict__ t0, int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = el + i * ncols;
  const float xd = fabsf(__fsub_rn(r[cx], pc[2 * i])), yd = fabsf(__fsub_rn(r[cy], pc[2 * i + 1]));
  t0[i] = __fsub_rn(full_drift_t, r[ctm]);
  idx[2 * i] = max(0, min((int)__fdiv_rn(xd, bin_size), nx - 1));
  idx[2 * i + 1] = max(0, min((int)__fdiv_rn(yd, bin_size), ny - 1));
}
}  // namespace

extern "C" int larnd_current_lut(const float* electrons_d, int64_t n, int32_t ncols, int32_t col_x, int32_t col_y, int32_t col_t,
                                 const float* pixels_coord_d, float response_full_drift_t, float response_bin_size, int32_t nx,
                                 int32_t ny, float* t0_d, int32_t* currents_idx_d, void* stream) {
  if (n < 0 || ncols < 1 || col_x < 0 || col_y < 0 || col_t < 0 || col_x >= ncols || col_y >= ncols || col_t >= ncols || nx < 1 || ny < 1 ||
      !(response_bin_size > 0) || (n > 0 && (!electrons_d || !pixels_coord_d || !t0_d || !currents_idx_d))) {
    larnd_set_error("larnd_current_lut: bad argument");
    return LARND_E_ARG;
  }
  if (n == 0) return LARND_OK;
  k_current_lut<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(electrons_d, n, ncols, col_x, col_y, col_t, pixels_coord_d,
                                                                             response_full_drift_t, response_bin_size, nx, ny, t0_d, currents_idx_d);
  LARND_LAUNCH_CHECK("k_current_lut");
  return LARND_OK;
}
