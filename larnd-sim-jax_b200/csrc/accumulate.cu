// K3 lut_accumulate (forward), sm_100a.
//
// Replaces the XLA lowering of simulate_signals (reference sim_jax.py:142-286): for every segment
//   * 25 transverse-diffusion bins ("main" contributions): 3-template Lagrange blend of the response rows,
//     deposited twice (weights 1-frac / frac, shifted by one tick)                       sim_jax.py:157-194
//   * (2n+1)^2 neighbour pixels: template 0 rows, full segment charge                     sim_jax.py:197-225
//   * boundary corrections from the running sums of the response                          sim_jax.py:228-261
//   * one scatter-add into the (Npix, Nticks) waveform buffer                             sim_jax.py:265-286
//
// Design (B200): no value/index streams are materialised and no shared-memory atomics are used (fp32
// smem atomics are CAS loops on sm_100a).  A CTA takes a chunk of consecutive segments; its work is cut
// into 25 + (2n+1)^2 *units* (one per diffusion bin / per relative neighbour pixel).  A warp owns a unit
// for the whole chunk and keeps that unit's current waveform-row window in REGISTERS (NS slots of 32
// ticks, lane = tick).  For each segment it gathers the response row with coalesced, L1/L2-resident loads
// and FMAs into the registers; the window is flushed with coalesced red.global.add.f32 only when the
// target row or the tick window changes (consecutive segments of a track share both), so global atomics
// drop from ~2x10^4 per segment to a few hundred per chunk.
#include "larnd_common.cuh"

namespace {

constexpr int ACC_THREADS = 256;
constexpr int ACC_WARPS = ACC_THREADS / 32;
constexpr int S = LARND_CHUNK;

struct AccArgs {
  const float* rec;
  int64_t n;
  const float* r0;
  const float* rm;
  const float* c0;
  const float* cm;
  int nt, L, Lp, ny_lut, nx_lut;
  int nticks;
  int nb, half2;  // bins per pixel; 2*(nb/2) - 1
  int n_neigh, P;
  int nxp, nyp;   // pixels per plane
  RowLookup lk;
  const int32_t* counts;
  float* wfs;
  int skip_garbage;
};

struct ChunkSmem {
  float4 seg[S];   // q, frac, T0 (int bits), bxm | bym << 8
  int2 key[S];     // ep, (mpx & 0xffff) | (mpy << 16)
  float a[S], b[S], c[S];
  int idx[S], bx[S], by[S];
  float wx[LARND_NB_TRAN_BINS][S], wy[LARND_NB_TRAN_BINS][S];
  int next_unit;
};

template <int NS>
__device__ __forceinline__ void flush_row(float (&acc)[NS], float& g0, bool& g0_used, int row, int tbase,
                                          const AccArgs& A, int lane) {
  if (row >= 0) {
    float* base = A.wfs + (int64_t)row * A.nticks;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      int col = tbase + 32 * j + lane;
      if (acc[j] != 0.0f && col >= 1 && col < A.nticks) atomicAdd(base + col, acc[j]);  // RED.E.ADD.F32
    }
    if (g0_used) {
      float g = g0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if (lane == 0 && g != 0.0f) atomicAdd(base, g);  // garbage tick 0
    }
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) acc[j] = 0.0f;
  g0 = 0.0f;
  g0_used = false;
}

// Adds one (segment, target-row) contribution: window samples k in [0,L) at ticks T0+k+1 (weight q1) and
// T0+k (weight q0), plus the boundary correction D at ticks T0 (q1) and T0-1 (q0).  NR response rows are
// blended with coefficients cf[].  x = tick - T0; rows are zero padded so x can be clamped to [-1, L+1].
template <int NS, int NR>
__device__ __forceinline__ void add_contribution(float (&acc)[NS], float& g0, bool& g0_used, int tbase, int T0,
                                                 float q0, float q1, float D, const float* const (&rows)[NR],
                                                 const float (&cf)[NR], const AccArgs& A, int lane) {
  const int L = A.L;
  const int xb0 = tbase - T0;  // x of lane 0, slot 0
  const bool fast = (T0 >= 2) && (T0 + L <= A.nticks - 1);
  if (fast) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int xs = xb0 + 32 * j;
      if (xs + 31 < -1 || xs > L) continue;  // warp-uniform: slot does not overlap [T0-1, T0+L]
      const int x = xs + lane;
      const int xc = min(max(x, -1), L + 1);
      float v0, v1;
      if (NR == 1) {
        v0 = __ldg(rows[0] + xc + 2);
        v1 = __ldg(rows[0] + xc + 1);
      } else {
        v0 = 0.0f; v1 = 0.0f;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          v0 = fmaf(cf[r], __ldg(rows[r] + xc + 2), v0);
          v1 = fmaf(cf[r], __ldg(rows[r] + xc + 1), v1);
        }
      }
      if (x == -1) v0 = D;
      if (x == 0) v1 = D;
      acc[j] = fmaf(q0, v0, fmaf(q1, v1, acc[j]));
    }
  } else {
    // window touches the ends of the readout: reference routes invalid ticks to column 0 (sim_jax.py:177-178,243-244)
    g0_used = true;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int xs = xb0 + 32 * j;
      if (xs + 31 < -1 || xs > L) continue;
      const int x = xs + lane;
      const int col = T0 + x;
      float v0 = 0.0f, v1 = 0.0f;
      const int xc = min(max(x, -1), L + 1);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        float c = (NR == 1) ? 1.0f : cf[r];
        v0 = fmaf(c, __ldg(rows[r] + xc + 2), v0);
        v1 = fmaf(c, __ldg(rows[r] + xc + 1), v1);
      }
      const float wterm = fmaf(q0, v0, q1 * v1);
      const float cterm = (x == -1 ? q0 * D : 0.0f) + (x == 0 ? q1 * D : 0.0f);
      const bool in_regs = (x >= -1 && x <= L);
      const bool valid_w = in_regs && col >= 2 && col <= A.nticks - 1;
      const bool valid_c = in_regs && col >= 1 && col <= A.nticks - 2;
      acc[j] += (valid_w ? wterm : 0.0f) + (valid_c ? cterm : 0.0f);
      g0 += (in_regs && !valid_w ? wterm : 0.0f) + (in_regs && !valid_c ? cterm : 0.0f);
    }
  }
}

__device__ __forceinline__ float boundary_delta(const float* crow, int ct, int nt, int L, float f, int lane) {
  // D = C[Nt-L] - ((1-f) C[ct] + f C[min(ct+1,Nt-1)])   (sim_jax.py:236-241): three lanes fetch, then broadcast
  int ci = lane == 0 ? ct : (lane == 1 ? min(ct + 1, nt - 1) : nt - L);
  float cv = lane < 3 ? __ldg(crow + ci) : 0.0f;
  float ca = __shfl_sync(0xffffffffu, cv, 0);
  float cb = __shfl_sync(0xffffffffu, cv, 1);
  float cl = __shfl_sync(0xffffffffu, cv, 2);
  return cl - (ca * (1.0f - f) + cb * f);
}

template <int NS>
__global__ void __launch_bounds__(ACC_THREADS)
k_lut_accumulate(const __grid_constant__ AccArgs A) {
  __shared__ ChunkSmem sm;
  if (A.counts[2] != 0) return;  // capacity overflow / bad event ids flagged upstream
  const int lane = threadIdx.x & 31;
  const int64_t s_base = (int64_t)blockIdx.x * S;
  const int ns = (int)min((int64_t)S, A.n - s_base);
  const int nb = A.nb;
  // ---- stage the chunk's segment records --------------------------------------------------------------
  for (int t = threadIdx.x; t < ns; t += ACC_THREADS) {
    const int64_t s = s_base + t;
    const float* rec = A.rec;
    const int* irec = reinterpret_cast<const int*>(A.rec);
    const int64_t n = A.n;
    int bx = irec[(int64_t)LARND_I_BX * n + s], by = irec[(int64_t)LARND_I_BY * n + s];
    int mpx = floordiv_i(bx, nb), mpy = floordiv_i(by, nb);
    int bxm = bx - mpx * nb, bym = by - mpy * nb;
    sm.seg[t] = make_float4(rec[(int64_t)LARND_F_Q * n + s], rec[(int64_t)LARND_F_FRAC * n + s],
                            __int_as_float(irec[(int64_t)LARND_I_T0 * n + s]), __int_as_float(bxm | (bym << 8)));
    sm.key[t] = make_int2(irec[(int64_t)LARND_I_EP * n + s], (mpx & 0xffff) | (mpy << 16));
    sm.a[t] = rec[(int64_t)LARND_F_A * n + s];
    sm.b[t] = rec[(int64_t)LARND_F_B * n + s];
    sm.c[t] = rec[(int64_t)LARND_F_C * n + s];
    sm.idx[t] = irec[(int64_t)LARND_I_IDX * n + s];
    sm.bx[t] = bx;
    sm.by[t] = by;
#pragma unroll
    for (int k = 0; k < LARND_NB_TRAN_BINS; ++k) {
      sm.wx[k][t] = rec[(int64_t)(LARND_F_WX0 + k) * n + s];
      sm.wy[k][t] = rec[(int64_t)(LARND_F_WY0 + k) * n + s];
    }
  }
  if (threadIdx.x == 0) sm.next_unit = 0;
  __syncthreads();

  RowLookup lk = A.lk;
  lk.n_unique = A.counts[0];
  lk.n_neg = A.counts[1];
  const int n_units = 25 + A.P * A.P;
  const int slack = (32 * NS - (A.L + 2)) / 2;
  float acc[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) acc[j] = 0.0f;
  float g0 = 0.0f;
  bool g0_used = false;

  for (;;) {
    int unit = 0;
    if (lane == 0) unit = atomicAdd(&sm.next_unit, 1);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= n_units) break;
    int cur_row = -1, tbase = 0;
    int cur_k0 = INT32_MIN, cur_k1 = INT32_MIN;
    if (unit >= 25) {
      // ---------------- neighbour unit: relative pixel (dx, dy), template 0, full charge --------------
      const int u = unit - 25;
      const int dx = u / A.P - A.n_neigh, dy = u % A.P - A.n_neigh;
      const bool centre = (dx == 0 && dy == 0);
      for (int t = 0; t < ns; ++t) {
        const float4 sg = sm.seg[t];
        const int2 key = sm.key[t];
        if (key.x != cur_k0 || key.y != cur_k1) {
          cur_k0 = key.x; cur_k1 = key.y;
          const int mpx = (int)(short)(key.y & 0xffff), mpy = key.y >> 16;
          int row;
          bool garbage;
          if (centre) { row = 0; garbage = true; }  // centre id is overwritten with -999 -> never matches -> row 0
          else {
            int pid = pixel2id_dev(mpx + dx, mpy + dy, key.x, A.nxp, A.nyp);
            row = lookup_row(lk, pid);
            garbage = row < 0 || pid < 0;
            if (row < 0) row = 0;  // sim_jax.py:724-725
          }
          if (A.skip_garbage && garbage) row = -1;
          if (row != cur_row) { flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane); cur_row = row; }
        }
        if (cur_row < 0) continue;
        const float q = sg.x, f = sg.y;
        if (q == 0.0f) continue;
        const int T0 = __float_as_int(sg.z);
        if (T0 - 1 < tbase || T0 + A.L >= tbase + 32 * NS) {
          flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane);
          tbase = T0 - 1 - slack;
        }
        const int bm = __float_as_int(sg.w);
        const int vx = 2 * (bm & 0xff) - A.half2 - 2 * nb * dx;
        const int vy = 2 * (bm >> 8) - A.half2 - 2 * nb * dy;
        const int ci = abs(vx) >> 1, cj = abs(vy) >> 1;
        const int bin = ci * A.ny_lut + cj;
        const float* const rows[1] = {A.r0 + (int64_t)bin * A.Lp};
        const float cf[1] = {1.0f};
        const float D = boundary_delta(A.c0 + (int64_t)bin * A.nt, A.nt - A.L - T0, A.nt, A.L, f, lane);
        add_contribution<NS, 1>(acc, g0, g0_used, tbase, T0, q * f, q * (1.0f - f), D, rows, cf, A, lane);
      }
    } else {
      // ---------------- main unit: diffusion bin (i, j), 3-template blend -----------------------------
      const int bi = unit / LARND_NB_TRAN_BINS, bj = unit % LARND_NB_TRAN_BINS;
      const int sym = (LARND_NB_TRAN_BINS - 1) / 2;
      int cix = 0, ciy = 0;
      for (int t = 0; t < ns; ++t) {
        const float4 sg = sm.seg[t];
        const int ep = sm.key[t].x;
        const int bxx = sm.bx[t] + bi - sym, byy = sm.by[t] + bj - sym;
        const int px = floordiv_i(bxx, nb), py = floordiv_i(byy, nb);
        const int k1 = (px & 0xffff) | (py << 16);
        if (ep != cur_k0 || k1 != cur_k1) {
          cur_k0 = ep; cur_k1 = k1;
          int pid = pixel2id_dev(px, py, ep, A.nxp, A.nyp);
          int row = lookup_row(lk, pid);  // not in the list -> dropped (sim_jax.py:152-154)
          if (A.skip_garbage && pid < 0) row = -1;
          if (row != cur_row) { flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane); cur_row = row; }
        }
        if (cur_row < 0) continue;
        const float q = sg.x, f = sg.y;
        const float qb = (sm.wx[bi][t] * sm.wy[bj][t]) * q;
        if (qb == 0.0f) continue;
        const int T0 = __float_as_int(sg.z);
        if (T0 - 1 < tbase || T0 + A.L >= tbase + 32 * NS) {
          flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane);
          tbase = T0 - 1 - slack;
        }
        cix = abs(2 * (bxx - px * nb) - A.half2) >> 1;
        ciy = abs(2 * (byy - py * nb) - A.half2) >> 1;
        const int idx = sm.idx[t];
        const int bin = cix * 5 + ciy;
        const float* const rows[3] = {A.rm + (int64_t)((idx - 1) * 25 + bin) * A.Lp,
                                      A.rm + (int64_t)(idx * 25 + bin) * A.Lp,
                                      A.rm + (int64_t)((idx + 1) * 25 + bin) * A.Lp};
        const float cf[3] = {sm.a[t], sm.b[t], sm.c[t]};
        const float D = boundary_delta(A.cm + (int64_t)(idx * 25 + bin) * A.nt, A.nt - A.L - T0, A.nt, A.L, f, lane);
        add_contribution<NS, 3>(acc, g0, g0_used, tbase, T0, qb * f, qb * (1.0f - f), D, rows, cf, A, lane);
      }
    }
    flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane);
  }
}

}  // namespace

int larnd_launch_accumulate(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                            int32_t npix_capacity, int32_t flags, float* wfs, const int32_t* counts, cudaStream_t st) {
  if (n == 0) return LARND_OK;
  AccArgs A;
  A.rec = ws.rec; A.n = n;
  A.r0 = lut->r0; A.rm = lut->rm; A.c0 = lut->c0; A.cm = lut->cm;
  A.nt = lut->nt; A.L = lut->L; A.Lp = lut->Lp; A.ny_lut = lut->ny; A.nx_lut = lut->nx;
  A.nticks = p.n_ticks;
  A.nb = p.nb_sampling_bins_per_pixel;
  A.half2 = 2 * (A.nb / 2) - 1;
  A.n_neigh = p.number_pix_neighbors;
  A.P = 2 * A.n_neigh + 1;
  A.nxp = p.n_pixels_x; A.nyp = p.n_pixels_y;
  A.lk.bitmap = ws.bitmap; A.lk.wprefix = ws.wprefix; A.lk.n_words = ws.n_words; A.lk.pid_offset = ws.pid_offset;
  A.lk.n_unique = 0; A.lk.n_neg = 0; A.lk.npix = npix_capacity;
  A.counts = counts;
  A.wfs = wfs;
  A.skip_garbage = flags & 1;
  const int64_t chunks = (n + S - 1) / S;
  const int need = lut->L + 2 + 32;  // window + room for the tick drift inside a chunk
  prof_begin(1, st);
  if (need <= 32 * 4) k_lut_accumulate<4><<<(unsigned)chunks, ACC_THREADS, 0, st>>>(A);
  else if (need <= 32 * 6) k_lut_accumulate<6><<<(unsigned)chunks, ACC_THREADS, 0, st>>>(A);
  else if (need <= 32 * 8) k_lut_accumulate<8><<<(unsigned)chunks, ACC_THREADS, 0, st>>>(A);
  else if (need <= 32 * 12) k_lut_accumulate<12><<<(unsigned)chunks, ACC_THREADS, 0, st>>>(A);
  else if (need <= 32 * 16) k_lut_accumulate<16><<<(unsigned)chunks, ACC_THREADS, 0, st>>>(A);
  else {
    larnd_set_error("signal_length %d too large for the register window (max %d)", lut->L, 32 * 16 - 34);
    return LARND_E_ARG;
  }
  prof_end(1, st);
  LARND_LAUNCH_CHECK("k_lut_accumulate");
  return LARND_OK;
}
