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
#include <stdlib.h>

#include "larnd_common.cuh"

namespace {

constexpr int ACC_THREADS = 256;
constexpr int S = LARND_CHUNK;

struct AccArgs {
  const float* rec;
  int64_t n;
  const float* r0;
  const float* rm;
  const float* c0;
  const float* cm;
  const float* sr;  // neighbourhood sums of r0 / c0 per in-pixel bin (larnd_common.cuh)
  const float* sc;
  int nt, L, Lp, ny_lut, nx_lut;
  int nticks;
  int64_t wstride;  // row stride of wfs in floats
  int nb, half2;  // bins per pixel; 2*(nb/2) - 1
  int n_neigh, P;
  int nxp, nyp;   // pixels per plane
  RowLookup lk;
  const int32_t* counts;
  float* wfs;
  unsigned long long* wfs64;  // deterministic mode: (npix, nticks) 64-bit fixed-point accumulators instead of wfs (else nullptr)
  int skip_garbage;
  int chunk;      // segments per CTA (<= S, larnd_chunk_size)
  int slow_only;  // 1: only segments whose window touches the ends of the readout (the rest is done by accumulate_sorted.cu)
  const int* n_slow;  // slow_only: device count of such segments (k_build_runs), 0 = nothing to do
};

constexpr int KP = 8;            // tick positions per run (impulse train length)
constexpr int SPAN_MAX = KP - 2;  // max (T0max - T0min) inside a run

// A *run* = consecutive segments that share (event, plane, sub-pixel bin, template index) and whose start ticks lie
// within SPAN_MAX of each other.  They read the same response rows for every unit, so their deposits are merged
// into an impulse train over tick positions before the response row is applied once per position:
//   h[m]  = sum q*f at T0 = tmin+m  +  sum q*(1-f) at T0 = tmin+m-1            (window part, sim_jax.py:190-194)
//   moments per T0 position for the boundary correction (sim_jax.py:236-247), which is bilinear in (1-f, f):
//   A1 = sum q(1-f), A2 = sum q(1-f)^2, A3 = sum q f(1-f), B1 = sum q f, B3 = sum q f^2
struct __align__(16) RunInfo {
  int start, len, tmin, span;
  int ep, pxy;      // event*ntpc+plane, main pixel (x | y << 16)
  int mpx, mpy;     // main pixel
  int bxm, bym;     // sub-pixel bin inside the main pixel, 0 .. nb-1
  int idx;          // longitudinal-diffusion template index
  int fast;         // whole window inside the readout (no garbage-tick handling needed)
};

struct ChunkSmem {
  float4 seg[S];   // q, frac, T0 (int bits), bxm | bym << 8
  int2 key[S];     // ep, (mpx & 0xffff) | (mpy << 16)
  float a[S], b[S], c[S];
  int idx[S], bx[S], by[S];
  float wx[LARND_NB_TRAN_BINS][S], wy[LARND_NB_TRAN_BINS][S];
  RunInfo run[S];
  // per-segment products used when a main (diffusion-bin) unit builds its impulse trains: multiply by Wx[i]*Wy[j]
  float4 pf[S];    // q f a, q f b, q f c, q f                (deposit at T0)
  float4 po[S];    // q(1-f) a, q(1-f) b, q(1-f) c, q(1-f)    (deposit at T0+1)
  float4 pm[S];    // q(1-f)^2, q f(1-f), q f^2, unused       (correction moments A2, A3, B3)
  float rh[S][KP], rA1[S][KP], rA2[S][KP], rA3[S][KP], rB1[S][KP], rB3[S][KP];  // neighbour impulse train + moments
  // transverse-diffusion bins that fall on the same pixel with the same response row are merged ("groups"); per
  // in-pixel bin index b: number of groups, pixel offset + 1, response index and member mask of each group
  unsigned char g_n[16], g_ox[16][5], g_ci[16][5], g_mask[16][5];
  unsigned char fastseg[S];
  int nruns;
  int next_unit;
};

// 2^-20 electrons-per-tick resolution, +-8.8e12 range: far below float32 resolution of any waveform sample that matters
constexpr double DET_SCALE = 1048576.0;
__device__ __forceinline__ unsigned long long det_fixed(float v) { return (unsigned long long)__double2ll_rn((double)v * DET_SCALE); }

__global__ void k_det_convert(const unsigned long long* __restrict__ acc, float* __restrict__ wfs, int64_t npix, int nticks, int64_t stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * nticks) return;
  const int64_t r = i / nticks;
  wfs[r * stride + (i - r * nticks)] = (float)((double)(long long)acc[i] * (1.0 / DET_SCALE));
}

template <int NS>
__device__ __forceinline__ void flush_row(float (&acc)[NS], float& g0, bool& g0_used, int row, int tbase,
                                          const AccArgs& A, int lane) {
  if (row >= 0 && A.wfs64) {
    // deterministic mode: every flushed window value is a fixed function of the chunk's segments; adding it as a 64-bit
    // fixed-point integer makes the total independent of the order in which warps and CTAs arrive
    unsigned long long* base = A.wfs64 + (int64_t)row * A.nticks;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      int col = tbase + 32 * j + lane;
      if (acc[j] != 0.0f && col >= 1 && col < A.nticks) atomicAdd(base + col, det_fixed(acc[j]));
    }
    if (g0_used) {
      float g = g0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if (lane == 0 && g != 0.0f) atomicAdd(base, det_fixed(g));
    }
  } else if (row >= 0) {
    float* base = A.wfs + (int64_t)row * A.wstride;
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

// Applies one impulse train to the register window: for every tick position j of the run,
//   acc[col] += h_j * Rblend[col - (tmin + j)]   for col - (tmin+j) in [0, L)
// Lane j holds h[NR][j]; it is broadcast with a shuffle.  Each lane walks one pointer per response row; the slot
// offset is an immediate (32*s floats) and lanes outside the response window are predicated off, so the inner
// loop is one compare + NR loads + NR FMAs per 32 ticks.
template <int NS, int NR, bool DUAL = false>
__device__ __forceinline__ void apply_train(float (&acc)[NS], int tbase, int tmin, int npos, const float (&h)[NR],
                                            const float* const (&rows)[NR], int L, int lane, float* accg = nullptr) {
  // x = col - (tmin + j) for slot 0 of this lane; response sample k lives at row[k + 2]
  int x0 = tbase + lane - tmin;
  const float* p[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) p[r] = rows[r] + (x0 + 2);
  for (int j = 0; j < npos; ++j, --x0) {
    float hj[NR];
    bool any = false;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      hj[r] = __shfl_sync(0xffffffffu, h[r], j);
      any |= (hj[r] != 0.0f);
    }
    if (any) {  // warp-uniform
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const bool in = (unsigned)(x0 + 32 * s) < (unsigned)L;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const float v = in ? __ldg(p[r] + 32 * s) : 0.0f;
          acc[s] = fmaf(hj[r], v, acc[s]);
          if (DUAL) accg[s] = fmaf(-hj[r], v, accg[s]);  // the same deposit leaves waveform row 0 (see the garbage-sum unit)
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) --p[r];
  }
}

// Adds the merged boundary corrections of a run: lane j holds E_j, which belongs to tick tmin - 1 + j.
template <int NS, bool DUAL = false>
__device__ __forceinline__ void add_corrections(float (&acc)[NS], int tbase, int tmin, float E, int lane, float* accg = nullptr) {
  const int off = tmin - 1 - tbase;           // >= 0 by construction of the window
  const int rot = off & 31, s0 = off >> 5;
  const float Er = __shfl_sync(0xffffffffu, E, (lane - rot) & 31);  // lane l now holds E_{(l - rot) mod 32}
  const float lo = (lane >= rot) ? Er : 0.0f;  // ticks in slot s0
  const float hi = (lane < rot) ? Er : 0.0f;   // wrapped into slot s0 + 1 (E_j = 0 for j >= KP keeps this exact)
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (s == s0) {            // warp-uniform; the window construction guarantees s0 + 1 < NS whenever hi != 0
      acc[s] += lo;
      if (DUAL) accg[s] -= lo;
      if (s + 1 < NS) {
        acc[s + 1] += hi;
        if (DUAL) accg[s + 1] -= hi;
      }
    }
  }
}

// merged boundary correction of a run for one response bin: lane m <-> T0 = tmin + m
__device__ __forceinline__ float run_correction(const float* crow, int tmin, int nt, int L, float A1, float A2, float A3,
                                                float B1, float B3, int lane) {
  int ct = nt - L - (tmin + lane);
  ct = max(0, min(ct, nt - 1));
  const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, nt - 1)), Cl = __ldg(crow + nt - L);
  const float e1 = Cl * A1 - Ca * A2 - Cb * A3;  // lands on tick T0      (weight 1-f)
  const float e0 = Cl * B1 - Ca * A3 - Cb * B3;  // lands on tick T0 - 1  (weight f)
  // impulse position j covers tick tmin + j - 1:  E_j = e0_j + e1_{j-1}
  float e1_up = __shfl_up_sync(0xffffffffu, e1, 1);
  if (lane == 0) e1_up = 0.0f;
  return e0 + e1_up;  // callers zero lanes > span + 1
}

template <int NS>
__global__ void __launch_bounds__(ACC_THREADS)
k_lut_accumulate(const __grid_constant__ AccArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChunkSmem& sm = *reinterpret_cast<ChunkSmem*>(smem_raw);
  if (A.counts[2] != 0) return;  // capacity overflow / bad event ids flagged upstream
  const int lane = threadIdx.x & 31;
  const int64_t s_base = (int64_t)blockIdx.x * A.chunk;
  const int ns = (int)min((int64_t)A.chunk, A.n - s_base);
  const int nb = A.nb;
  const int L = A.L;
  if (A.slow_only) {  // nothing to do for a chunk without boundary segments (the common case)
    if (A.n_slow && *A.n_slow == 0) return;  // ... nor for a batch without any (counted by k_build_runs): no record is read
    int slow = 0;
    if ((int)threadIdx.x < ns)
      slow = !seg_is_fast(reinterpret_cast<const int*>(A.rec)[(int64_t)LARND_I_T0 * A.n + s_base + threadIdx.x], L, A.nticks);
    if (!__syncthreads_or(slow)) return;
  }
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
    sm.fastseg[t] = seg_is_fast(irec[(int64_t)LARND_I_T0 * n + s], L, A.nticks);
    {
      const float q = sm.seg[t].x, f = sm.seg[t].y, o = 1.0f - f;
      const float qf = q * f, qo = q * o;
      sm.pf[t] = make_float4(qf * sm.a[t], qf * sm.b[t], qf * sm.c[t], qf);
      sm.po[t] = make_float4(qo * sm.a[t], qo * sm.b[t], qo * sm.c[t], qo);
      sm.pm[t] = make_float4(qo * o, qf * o, qf * f, 0.0f);
    }
#pragma unroll
    for (int k = 0; k < LARND_NB_TRAN_BINS; ++k) {
      sm.wx[k][t] = rec[(int64_t)(LARND_F_WX0 + k) * n + s];
      sm.wy[k][t] = rec[(int64_t)(LARND_F_WY0 + k) * n + s];
    }
  }
  if (threadIdx.x < nb && threadIdx.x < 16) build_bin_groups(threadIdx.x, nb, A.half2, sm.g_n[threadIdx.x], sm.g_ox[threadIdx.x], sm.g_ci[threadIdx.x], sm.g_mask[threadIdx.x]);
  __syncthreads();
  // ---- cut the chunk into runs (serial, ~ns steps; negligible next to the accumulate work) ------------------
  if (threadIdx.x == 0) {
    int nr = 0;
    int cur = -1, tmin = 0, tmax = 0;
    for (int t = 0; t < ns; ++t) {
      const int T0 = __float_as_int(sm.seg[t].z);
      if (A.slow_only && sm.fastseg[t]) {  // handled by the class-sorted kernel
        if (cur >= 0) { sm.run[cur].len = t - sm.run[cur].start; sm.run[cur].tmin = tmin; sm.run[cur].span = tmax - tmin; }
        cur = -1;
        continue;
      }
      bool fresh = cur < 0;
      if (!fresh) {
        const int t0s = sm.run[cur].start;
        fresh = sm.key[t].x != sm.key[t0s].x || sm.bx[t] != sm.bx[t0s] || sm.by[t] != sm.by[t0s] || sm.idx[t] != sm.idx[t0s] ||
                sm.fastseg[t] != sm.fastseg[t0s] || max(tmax, T0) - min(tmin, T0) > SPAN_MAX;
      }
      if (fresh) {
        if (cur >= 0) { sm.run[cur].len = t - sm.run[cur].start; sm.run[cur].tmin = tmin; sm.run[cur].span = tmax - tmin; }
        cur = nr++;
        sm.run[cur].start = t;
        tmin = tmax = T0;
      } else {
        tmin = min(tmin, T0);
        tmax = max(tmax, T0);
      }
    }
    if (cur >= 0) { sm.run[cur].len = ns - sm.run[cur].start; sm.run[cur].tmin = tmin; sm.run[cur].span = tmax - tmin; }
    sm.nruns = nr;
    sm.next_unit = 0;
  }
  __syncthreads();
  const int nruns = sm.nruns;
  // ---- per-run neighbour impulse train and correction moments (thread per run) ------------------------------
  for (int r = threadIdx.x; r < nruns; r += ACC_THREADS) {
    RunInfo& R = sm.run[r];
    const int t0s = R.start;
    R.ep = sm.key[t0s].x; R.pxy = sm.key[t0s].y; R.idx = sm.idx[t0s];
    R.mpx = floordiv_i(sm.bx[t0s], nb); R.mpy = floordiv_i(sm.by[t0s], nb);
    R.bxm = sm.bx[t0s] - R.mpx * nb; R.bym = sm.by[t0s] - R.mpy * nb;
    R.fast = (R.tmin >= 2) && (R.tmin + R.span + L <= A.nticks - 1);
    for (int k = 0; k < KP; ++k) { sm.rh[r][k] = 0.f; sm.rA1[r][k] = 0.f; sm.rA2[r][k] = 0.f; sm.rA3[r][k] = 0.f; sm.rB1[r][k] = 0.f; sm.rB3[r][k] = 0.f; }
    for (int t = t0s; t < t0s + R.len; ++t) {
      // merge the 5 per-axis diffusion weights into per-group weights (in place: slot g <- sum of the group's members)
      float vx[5], vy[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) { vx[k] = sm.wx[k][t]; vy[k] = sm.wy[k][t]; }
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        const int mx = g < sm.g_n[R.bxm] ? sm.g_mask[R.bxm][g] : 0, my = g < sm.g_n[R.bym] ? sm.g_mask[R.bym][g] : 0;
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) { sx += (mx >> k & 1) ? vx[k] : 0.f; sy += (my >> k & 1) ? vy[k] : 0.f; }
        sm.wx[g][t] = sx;
        sm.wy[g][t] = sy;
      }
      const float4 sg = sm.seg[t];
      const float q = sg.x, f = sg.y, omf = 1.0f - f;
      const int m = __float_as_int(sg.z) - R.tmin;
      sm.rh[r][m] += q * f;
      sm.rh[r][m + 1] += q * omf;
      sm.rA1[r][m] += q * omf; sm.rA2[r][m] += q * omf * omf; sm.rA3[r][m] += q * f * omf;
      sm.rB1[r][m] += q * f;   sm.rB3[r][m] += q * f * f;
    }
  }
  __syncthreads();

  RowLookup lk = A.lk;
  lk.n_unique = A.counts[0];
  lk.n_neg = A.counts[1];
  // units: 25 diffusion bins, (2n+1)^2 relative neighbour pixels, and one "garbage-sum" unit (see below)
  const int n_neigh_units = A.P * A.P;
  const int n_units = 25 + n_neigh_units + (A.skip_garbage ? 0 : 1);
  float acc[NS], accg[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) { acc[j] = 0.0f; accg[j] = 0.0f; }
  float g0 = 0.0f, g0g = 0.0f;
  bool g0_used = false, g0g_used = false;

  for (;;) {
    int unit = 0;
    if (lane == 0) unit = atomicAdd(&sm.next_unit, 1);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= n_units) break;
    int cur_row = -1, tbase = 0;
    int cur_k0 = INT32_MIN, cur_k1 = INT32_MIN;
    if (unit >= 25) {
      // ---------------- neighbour units: template 0, full segment charge (sim_jax.py:197-225,250-261) ----------
      // A neighbour pixel that is not a main pixel of the batch is routed to waveform row 0 by the reference
      // (sim_jax.py:724-725); along a track that is ~85 % of the (2n+1)^2 neighbours.  By linearity
      //     row0 += sum_{all neighbours} h * R_u  -  sum_{neighbours with their own row} h * R_u
      // so one unit applies the precomputed neighbourhood-sum row (A.sr / A.sc) to row 0, and only the neighbours
      // that own a row do real work: they add to their row and subtract the same deposit from row 0.
      const bool sum_unit = unit == 25 + n_neigh_units;
      const int u = unit - 25;
      const int dx = sum_unit ? 0 : u / A.P - A.n_neigh, dy = sum_unit ? 0 : u % A.P - A.n_neigh;
      const bool centre = (dx == 0 && dy == 0);
      if (centre && !sum_unit) continue;  // the centre id is overwritten with -999: never matches, lives in row 0 only
      const bool dual = !sum_unit && !A.skip_garbage;
      for (int r = 0; r < nruns; ++r) {
        const RunInfo R = sm.run[r];
        if (R.ep != cur_k0 || R.pxy != cur_k1) {
          cur_k0 = R.ep; cur_k1 = R.pxy;
          int row = 0;
          if (!sum_unit) {
            const int pid = pixel2id_dev(R.mpx + dx, R.mpy + dy, R.ep, A.nxp, A.nyp);
            row = lookup_row(lk, pid);
            if (row <= 0) row = -1;                              // row 0 / not a main pixel: covered by the sum unit
            else if (A.skip_garbage && pid < 0) row = -1;
          }
          if (row != cur_row) {
            flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane);
            if (dual) flush_row<NS>(accg, g0g, g0g_used, 0, tbase, A, lane);
            cur_row = row;
          }
        }
        if (cur_row < 0) continue;
        const int span = R.span;
        if (R.tmin - 1 < tbase || R.tmin + span + L >= tbase + 32 * NS) {
          flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane);
          if (dual) flush_row<NS>(accg, g0g, g0g_used, 0, tbase, A, lane);
          tbase = R.tmin - 1 - (32 * NS - (L + 2 + span)) / 2;
        }
        const float* rowp;
        const float* crow;
        if (sum_unit) {
          const int sb = R.bxm * nb + R.bym;
          rowp = A.sr + (int64_t)sb * A.Lp;
          crow = A.sc + (int64_t)sb * A.nt;
        } else {
          const int vx = 2 * R.bxm - A.half2 - 2 * nb * dx;
          const int vy = 2 * R.bym - A.half2 - 2 * nb * dy;
          const int bin = (abs(vx) >> 1) * A.ny_lut + (abs(vy) >> 1);
          rowp = A.r0 + (int64_t)bin * A.Lp;
          crow = A.c0 + (int64_t)bin * A.nt;
        }
        const float* const rows[1] = {rowp};
        if (R.fast) {
          const int kk = min(lane, KP - 1);
          const float hl = lane < KP ? sm.rh[r][kk] : 0.f;
          float E = run_correction(crow, R.tmin, A.nt, L, sm.rA1[r][kk], sm.rA2[r][kk], sm.rA3[r][kk], sm.rB1[r][kk], sm.rB3[r][kk], lane);
          if (lane > span + 1) E = 0.f;
          const float h[1] = {hl};
          if (dual) {
            apply_train<NS, 1, true>(acc, tbase, R.tmin, span + 2, h, rows, L, lane, accg);
            add_corrections<NS, true>(acc, tbase, R.tmin, E, lane, accg);
          } else {
            apply_train<NS, 1>(acc, tbase, R.tmin, span + 2, h, rows, L, lane);
            add_corrections<NS>(acc, tbase, R.tmin, E, lane);
          }
        } else {
          // run touches the ends of the readout window: per-segment path with garbage-tick handling
          const float cf[1] = {1.0f};
          for (int t = R.start; t < R.start + R.len; ++t) {
            const float4 sg = sm.seg[t];
            const float q = sg.x, f = sg.y;
            if (q == 0.0f) continue;
            const int T0 = __float_as_int(sg.z);
            const float D = boundary_delta(crow, A.nt - L - T0, A.nt, L, f, lane);
            add_contribution<NS, 1>(acc, g0, g0_used, tbase, T0, q * f, q * (1.0f - f), D, rows, cf, A, lane);
            if (dual) add_contribution<NS, 1>(accg, g0g, g0g_used, tbase, T0, -q * f, -q * (1.0f - f), D, rows, cf, A, lane);
          }
        }
      }
      if (dual) flush_row<NS>(accg, g0g, g0g_used, 0, tbase, A, lane);
    } else {
      // ---------------- main unit: group (gx, gy) of merged diffusion bins, 3-template blend -----------
      // the 25 bins of the stencil collapse to n_gx * n_gy (9 .. 25, 19 on average) distinct (pixel, response row) pairs
      const int bi = unit / LARND_NB_TRAN_BINS, bj = unit % LARND_NB_TRAN_BINS;  // group indices
      for (int r = 0; r < nruns; ++r) {
        const RunInfo R = sm.run[r];
        if (bi >= sm.g_n[R.bxm] || bj >= sm.g_n[R.bym]) continue;   // this run has fewer groups
        const int px = R.mpx + (int)sm.g_ox[R.bxm][bi] - 1, py = R.mpy + (int)sm.g_ox[R.bym][bj] - 1;
        const int k1 = (px & 0xffff) | (py << 16);
        if (R.ep != cur_k0 || k1 != cur_k1) {
          cur_k0 = R.ep; cur_k1 = k1;
          int pid = pixel2id_dev(px, py, R.ep, A.nxp, A.nyp);
          int row = lookup_row(lk, pid);  // not in the list -> dropped (sim_jax.py:152-154)
          if (A.skip_garbage && pid < 0) row = -1;
          if (row != cur_row) { flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane); cur_row = row; }
        }
        if (cur_row < 0) continue;
        const int span = R.span;
        if (R.tmin - 1 < tbase || R.tmin + span + L >= tbase + 32 * NS) {
          flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane);
          tbase = R.tmin - 1 - (32 * NS - (L + 2 + span)) / 2;
        }
        const int cix = sm.g_ci[R.bxm][bi], ciy = sm.g_ci[R.bym][bj];
        const int idx = R.idx;
        const int bin = cix * 5 + ciy;
        const float* const rows[3] = {A.rm + (int64_t)((idx - 1) * 25 + bin) * A.Lp,
                                      A.rm + (int64_t)(idx * 25 + bin) * A.Lp,
                                      A.rm + (int64_t)((idx + 1) * 25 + bin) * A.Lp};
        const float* crow = A.cm + (int64_t)(idx * 25 + bin) * A.nt;
        if (R.fast) {
          // build this bin's impulse trains (one per template) and correction moments: lane <-> tick position
          float h[3] = {0.f, 0.f, 0.f};
          float A1 = 0.f, A2 = 0.f, A3 = 0.f, B1 = 0.f, B3 = 0.f;
          for (int t = R.start; t < R.start + R.len; ++t) {
            const float w = sm.wx[bi][t] * sm.wy[bj][t];
            const int m = __float_as_int(sm.seg[t].z) - R.tmin;
            const float4 pf = sm.pf[t], po = sm.po[t], pm = sm.pm[t];
            const float w0 = (lane == m) ? w : 0.0f;       // deposits / moments at T0
            const float w1 = (lane == m + 1) ? w : 0.0f;   // (1-f) deposit one tick later
            h[0] = fmaf(w0, pf.x, fmaf(w1, po.x, h[0]));
            h[1] = fmaf(w0, pf.y, fmaf(w1, po.y, h[1]));
            h[2] = fmaf(w0, pf.z, fmaf(w1, po.z, h[2]));
            A1 = fmaf(w0, po.w, A1); A2 = fmaf(w0, pm.x, A2); A3 = fmaf(w0, pm.y, A3);
            B1 = fmaf(w0, pf.w, B1); B3 = fmaf(w0, pm.z, B3);
          }
          float E = run_correction(crow, R.tmin, A.nt, L, A1, A2, A3, B1, B3, lane);
          if (lane > span + 1) E = 0.f;
          apply_train<NS, 3>(acc, tbase, R.tmin, span + 2, h, rows, L, lane);
          add_corrections<NS>(acc, tbase, R.tmin, E, lane);
        } else {
          for (int t = R.start; t < R.start + R.len; ++t) {
            const float4 sg = sm.seg[t];
            const float q = sg.x, f = sg.y;
            const float qb = (sm.wx[bi][t] * sm.wy[bj][t]) * q;
            if (qb == 0.0f) continue;
            const int T0 = __float_as_int(sg.z);
            const float cf[3] = {sm.a[t], sm.b[t], sm.c[t]};
            const float D = boundary_delta(crow, A.nt - L - T0, A.nt, L, f, lane);
            add_contribution<NS, 3>(acc, g0, g0_used, tbase, T0, qb * f, qb * (1.0f - f), D, rows, cf, A, lane);
          }
        }
      }
    }
    flush_row<NS>(acc, g0, g0_used, cur_row, tbase, A, lane);
  }
}

}  // namespace

int larnd_launch_accumulate(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                            int32_t npix_capacity, int32_t flags, float* wfs, int64_t wfs_stride, const int32_t* counts, cudaStream_t st,
                            unsigned long long* det_acc) {
  if (n == 0) return LARND_OK;
  if (det_acc) {   // deterministic mode: chunk kernel only, 64-bit fixed-point accumulators, converted at the end
    flags = (flags | LARND_FLAG_IMPL_CHUNK) & ~LARND_FLAG_IMPL_SORTED;
    LARND_CUDA(cudaMemsetAsync(det_acc, 0, (size_t)npix_capacity * p.n_ticks * sizeof(unsigned long long), st));
  }
  const bool slow_only = (flags & LARND_ACC_SLOW_ONLY) != 0;
  // (the forward tile kernel addresses the waveform buffer with signed 32-bit element offsets)
  {
    int rc0 = larnd_lut_check_neighbours(lut, p.nb_sampling_bins_per_pixel, p.number_pix_neighbors);
    if (rc0) return rc0;
  }
  // (the tile kernel addresses the waveform buffer with signed 32-bit element offsets)
  if (!slow_only && larnd_sorted_supported(p, lut) && (int64_t)npix_capacity * wfs_stride < ((int64_t)1 << 31)) {
    // large batches: class-sorted kernel (accumulate_sorted.cu); LARND_FLAG_IMPL_CHUNK / _SORTED override the size rule
    bool sorted = n >= LARND_SORTED_MIN_SEGMENTS;
    if (flags & LARND_FLAG_IMPL_SORTED) sorted = true;
    if (flags & LARND_FLAG_IMPL_CHUNK) sorted = false;
    if (sorted) return larnd_launch_accumulate_sorted(n, p, lut, ws, npix_capacity, flags, wfs, wfs_stride, counts, st);
  }
  AccArgs A;
  A.rec = ws.rec; A.n = n;
  A.r0 = lut->r0; A.rm = lut->rm; A.c0 = lut->c0; A.cm = lut->cm; A.sr = lut->sr; A.sc = lut->sc;
  A.nt = lut->nt; A.L = lut->L; A.Lp = lut->Lp; A.ny_lut = lut->ny; A.nx_lut = lut->nx;
  A.nticks = p.n_ticks;
  A.wstride = wfs_stride;
  A.nb = p.nb_sampling_bins_per_pixel;
  A.half2 = 2 * (A.nb / 2) - 1;
  A.n_neigh = p.number_pix_neighbors;
  A.P = 2 * A.n_neigh + 1;
  A.nxp = p.n_pixels_x; A.nyp = p.n_pixels_y;
  A.lk.bitmap = ws.bitmap; A.lk.wprefix = ws.wprefix; A.lk.n_words = ws.n_words; A.lk.pid_offset = ws.pid_offset;
  A.lk.n_unique = 0; A.lk.n_neg = 0; A.lk.npix = npix_capacity;
  A.counts = counts;
  A.wfs = wfs;
  A.wfs64 = det_acc;
  A.skip_garbage = flags & LARND_FLAG_SKIP_GARBAGE;
  A.slow_only = slow_only ? 1 : 0;
  A.n_slow = slow_only ? ws.gcnt + 3 /* GC_SLOW, sorted_runs.cuh */ : nullptr;
  A.chunk = (slow_only || det_acc) ? S : larnd_chunk_size(n);   // (the deterministic sums are defined per 128-segment chunk)
  const int64_t chunks = (n + A.chunk - 1) / A.chunk;
  // register window = run window (L + 2 + span) + slack for the tick drift between consecutive runs of a track;
  // more slack = fewer flushes but more predicated-off slots in the inner loop.
  int need = lut->L + 2 + SPAN_MAX;  // measured on B200: the tightest window wins (4 slots vs 6 at L=100: -15 % time)
  const size_t smem = sizeof(ChunkSmem);
  static bool attr_done = false;
  if (!attr_done) {
    LARND_CUDA(cudaFuncSetAttribute(k_lut_accumulate<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_lut_accumulate<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_lut_accumulate<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_lut_accumulate<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_lut_accumulate<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_lut_accumulate<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int ns_sel = (need + 31) / 32;
  if (!slow_only) prof_begin(1, st);
  if (ns_sel <= 4) k_lut_accumulate<4><<<(unsigned)chunks, ACC_THREADS, smem, st>>>(A);
  else if (ns_sel <= 5) k_lut_accumulate<5><<<(unsigned)chunks, ACC_THREADS, smem, st>>>(A);
  else if (ns_sel <= 6) k_lut_accumulate<6><<<(unsigned)chunks, ACC_THREADS, smem, st>>>(A);
  else if (ns_sel <= 8) k_lut_accumulate<8><<<(unsigned)chunks, ACC_THREADS, smem, st>>>(A);
  else if (ns_sel <= 12) k_lut_accumulate<12><<<(unsigned)chunks, ACC_THREADS, smem, st>>>(A);
  else if (ns_sel <= 16) k_lut_accumulate<16><<<(unsigned)chunks, ACC_THREADS, smem, st>>>(A);
  else {
    larnd_set_error("signal_length %d too large for the register window (max %d)", lut->L, 32 * 16 - 2 - SPAN_MAX);
    return LARND_E_ARG;
  }
  if (!slow_only) prof_end(1, st);
  LARND_LAUNCH_CHECK("k_lut_accumulate");
  if (det_acc) {
    const int64_t tot = (int64_t)npix_capacity * p.n_ticks;
    k_det_convert<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(det_acc, wfs, npix_capacity, p.n_ticks, wfs_stride);
    LARND_LAUNCH_CHECK("k_det_convert");
  }
  return LARND_OK;
}
