// K3b' class-sorted VJP of simulate_wfs (reference: jax.grad through sim_jax.py:142-286,375-453), sm_100a —
// the large-batch path of larnd_lut_backward.  Same mathematics as accumulate_bwd.cu, different decomposition.
//
// Runs are sorted by response class (sorted_runs.cuh).  One WARP owns a tile (<= 32 runs of one class, uniform
// number of impulse positions NPOS) and walks its units (merged diffusion-bin groups, neighbour pixels):
//   main unit   lane <-> run.  The upstream-gradient windows of the 32 target rows are staged in shared memory
//               (coalesced loads), the three response rows of the unit sit in shared memory as float4 and are
//               BROADCAST to all lanes, so the correlation
//                     G_r[j] = sum_x g[row, tmin - 1 + x] * R_r[x - 1 - j]
//               is 3*NPOS independent FFMA chains per lane — no cross-lane reduction at all.  Each lane then walks
//               the segments of its run and adds d/dq, d/dfrac, d/d(a,b,c), d/dWx, d/dWy to per-segment accumulators
//               in shared memory (private to the lane: no atomics).
//   neighbour   only (run, neighbour) pairs that own a non-garbage waveform row do work (the upstream gradient of
//               garbage rows is zero, checked on the device by k_garbage_grad_flag): lane <-> tick correlation with
//               the response row in registers and one 10-shuffle reduction per pair.
// Finally lane <-> segment applies the closed-form chain rule through drift / quench / diffusion (K1b) and the 15
// parameter gradients are reduced per warp into partials summed in double by k_reduce_partials.
// If garbage-row gradients are non-zero, or for segments whose window ends beyond the readout, accumulate_bwd.cu's
// kernel does the work instead (device-side switch, no host synchronisation).
#include "sorted_runs.cuh"

namespace {

constexpr int SEGCAP = 128;        // segments staged per batch of runs (a tile is cut into batches if it holds more)
constexpr int GCH = 60;            // ticks of the gradient windows staged per pass: a multiple of every NPOS <= 6
constexpr int GSTR = GCH + 1;      // odd stride: conflict-free for lane <-> run and lane <-> tick accesses
constexpr int RPAD = KPT + 1;      // zero padding on both sides of the response rows in shared memory
constexpr int LMAX_S = 186;        // 32*6 - 2 - SPAN_MAX_S
constexpr int GS = 3 * KPT + 1;    // per-run stride of the G buffer (odd)
constexpr int NACC = 15;           // dq, dfrac, da, db, dc, dWxg[5], dWyg[5]

struct BwdSortArgs {
  SortArgs S;
  const float* g;
  int64_t g_stride;
  float* partials;   // [gridDim][16]
  int force_skip;    // caller promises zero gradients on garbage rows
  const int* garbage_grad_nonzero;
};

struct WarpSmem {
  int4 run[TR];
  int ep[TR], mpx[TR], mpy[TR], soff[TR];
  float gw[TR * GSTR];
  float4 Rs[LMAX_S + 2 * RPAD + KPT];
  float Gs[TR * GS];
  float q[SEGCAP], f[SEGCAP], ca[SEGCAP], cb[SEGCAP], cc[SEGCAP];
  int m[SEGCAP], sid[SEGCAP];
  float acc[NACC][SEGCAP];
  unsigned char g_n[16], g_ox[16][5], g_ci[16][5], g_mask[16][5];
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums 8 lane-partial values over the warp with 10 shuffles; the total of v[j] is returned in lane j (j < 8)
__device__ __forceinline__ float reduce8_to_lane_s(const float (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float w[4], x[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b3 ? w[i] : w[i + 2], keep = b3 ? w[i + 2] : w[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float y = (b2 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, b2 ? x[0] : x[1], 4);
  y += __shfl_xor_sync(0xffffffffu, y, 2);
  y += __shfl_xor_sync(0xffffffffu, y, 1);
  return __shfl_sync(0xffffffffu, y, (lane & 7) * 4);
}

// lane <-> run correlation of the staged gradient windows with the broadcast response rows (see file header)
template <int NPOS>
__device__ __forceinline__ void correlate_runs(WarpSmem& sm, const BwdSortArgs& A, int lane, int row, int tmin, int r_lo, int r_hi,
                                               float* Gout /* &sm.Gs[lane*GS] */) {
  const int L = A.S.L, nticks = A.S.nticks;
  float G[3][NPOS];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int j = 0; j < NPOS; ++j) G[r][j] = 0.0f;
  const int xend = L + NPOS;  // x - 1 - j < L for all x < L + NPOS... samples beyond are zero padding
  for (int c0 = 0; c0 < xend; c0 += GCH) {
    __syncwarp();
    // stage ticks [c0, c0 + GCH) of every live window: lane <-> tick, coalesced
    for (int p = r_lo; p < r_hi; ++p) {
      const int rowp = __shfl_sync(0xffffffffu, row, p);
      const int tminp = __shfl_sync(0xffffffffu, tmin, p);
      if (rowp < 0) continue;
      const float* grow = A.g + (int64_t)rowp * A.g_stride;
#pragma unroll
      for (int xl = lane; xl < GCH; xl += 32) {
        const int col = tminp - 1 + c0 + xl;
        sm.gw[p * GSTR + xl] = (col >= 2 && col <= nticks - 1) ? __ldg(grow + col) : 0.0f;  // window samples live on ticks >= 2
      }
    }
    __syncwarp();
    const int xe = min(GCH, ((xend - c0 + NPOS - 1) / NPOS) * NPOS);
    float4 Rwin[NPOS];
#pragma unroll
    for (int j = 1; j < NPOS; ++j) Rwin[(NPOS - j) % NPOS] = sm.Rs[c0 - 1 - j + RPAD];  // R[x0 - 1 - j]
    const float* gwl = sm.gw + lane * GSTR;
    for (int x0 = 0; x0 < xe; x0 += NPOS) {
#pragma unroll
      for (int u = 0; u < NPOS; ++u) {
        const float gv = gwl[x0 + u];
        Rwin[u] = sm.Rs[c0 + x0 + u - 1 + RPAD];  // R[x - 1], broadcast
#pragma unroll
        for (int j = 0; j < NPOS; ++j) {
          const float4 R = Rwin[(u - j + NPOS) % NPOS];  // R[x - 1 - j]
          G[0][j] = fmaf(gv, R.x, G[0][j]);
          G[1][j] = fmaf(gv, R.y, G[1][j]);
          G[2][j] = fmaf(gv, R.z, G[2][j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NPOS; ++j) { Gout[3 * j] = G[0][j]; Gout[3 * j + 1] = G[1][j]; Gout[3 * j + 2] = G[2][j]; }
}

template <int NS>
__global__ void __launch_bounds__(32)
k_bwd_tiles(const __grid_constant__ BwdSortArgs A, const __grid_constant__ larnd_params_t p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpSmem& sm = *reinterpret_cast<WarpSmem*>(smem_raw);
  const SortArgs& S = A.S;
  const int lane = threadIdx.x;
  float gacc[LARND_NPARAMS];
#pragma unroll
  for (int k = 0; k < LARND_NPARAMS; ++k) gacc[k] = 0.0f;
  const bool garbage_needed = !A.force_skip && (*A.garbage_grad_nonzero != 0);
  const bool dead = S.counts[2] != 0 || garbage_needed;  // accumulate_bwd.cu's kernel takes over
  const int nb = S.nb, L = S.L, nt = S.nt;
  const int64_t n = S.n;
  const int* irec = reinterpret_cast<const int*>(S.rec);
  RowLookup lk = S.lk;
  lk.n_unique = S.counts[0];
  lk.n_neg = S.counts[1];
  if (lane < nb && lane < 16) {
    const int bq = lane;
    int ng = 0;
    for (int i = 0; i < LARND_NB_TRAN_BINS; ++i) {
      int qb = bq + i - (LARND_NB_TRAN_BINS - 1) / 2, ox = 0;
      if (qb < 0) { qb += nb; ox = -1; } else if (qb >= nb) { qb -= nb; ox = 1; }
      const int ci = abs(2 * qb - S.half2) >> 1;
      int g = -1;
      for (int k = 0; k < ng; ++k)
        if (sm.g_ox[bq][k] == ox + 1 && sm.g_ci[bq][k] == ci) g = k;
      if (g < 0) { g = ng++; sm.g_ox[bq][g] = ox + 1; sm.g_ci[bq][g] = ci; sm.g_mask[bq][g] = 0; }
      sm.g_mask[bq][g] |= 1 << i;
    }
    sm.g_n[bq] = ng;
  }
  __syncwarp();
  const int ntiles = dead ? 0 : S.gcnt[1];

  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(S.gcnt + 3, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= ntiles) break;
    const int4 ti = S.tile_info[tile];
    const int cls = ti.x, count = ti.z;
    const int span = cls % (SPAN_MAX_S + 1), cls_b = cls / (SPAN_MAX_S + 1);
    const int bym = cls_b % nb, bxm = (cls_b / nb) % nb, idx = cls_b / (nb * nb);
    const int npos = span + 2;
    // ---- stage the runs --------------------------------------------------------------------------------------
    __syncwarp();
    int len = 0, tmin = 0, start = 0;
    if (lane < count) {
      const int4 e = S.runs[ti.y + lane];
      sm.run[lane] = e;
      start = e.x; len = e.y & 0xffff; tmin = e.z;
      sm.ep[lane] = irec[(int64_t)LARND_I_EP * n + start];
      sm.mpx[lane] = floordiv_i(irec[(int64_t)LARND_I_BX * n + start], nb);
      sm.mpy[lane] = floordiv_i(irec[(int64_t)LARND_I_BY * n + start], nb);
    }
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    const int off_all = inc - len;  // segment offset of this lane's run inside the tile
    // ---- batches of runs whose segments fit the staging buffers ------------------------------------------------
    int r_lo = 0;
    while (r_lo < count) {
      const int base = __shfl_sync(0xffffffffu, off_all, r_lo);
      // runs r_lo .. r_hi-1: the longest prefix (from r_lo) whose segments fit; inc is non-decreasing so `fits` is a prefix
      const unsigned fits = __ballot_sync(0xffffffffu, lane >= r_lo && lane < count && inc - base <= SEGCAP);
      const int r_hi = r_lo + __popc(fits);
      const bool active = lane >= r_lo && lane < r_hi;
      const int so = off_all - base;
      const int nseg = __shfl_sync(0xffffffffu, inc, r_hi - 1) - base;
      __syncwarp();
      if (active) {
        sm.soff[lane] = so;
        for (int t = 0; t < len; ++t) sm.sid[so + t] = start + t;
      }
      __syncwarp();
      for (int i = lane; i < nseg; i += 32) {
        const int64_t s = sm.sid[i];
        sm.q[i] = S.rec[(int64_t)LARND_F_Q * n + s];
        sm.f[i] = S.rec[(int64_t)LARND_F_FRAC * n + s];
        sm.ca[i] = S.rec[(int64_t)LARND_F_A * n + s];
        sm.cb[i] = S.rec[(int64_t)LARND_F_B * n + s];
        sm.cc[i] = S.rec[(int64_t)LARND_F_C * n + s];
        // segments outside every TPC carry q == 0 and dq/dtheta == 0 (mask factor): no gradient (flag -> m = -1)
        const bool inside = irec[(int64_t)LARND_I_FLAGS * n + s] & 1;
        sm.m[i] = inside ? irec[(int64_t)LARND_I_T0 * n + s] : INT32_MIN;
#pragma unroll
        for (int k = 0; k < NACC; ++k) sm.acc[k][i] = 0.0f;
      }
      __syncwarp();
      if (active)
        for (int t = 0; t < len; ++t)
          if (sm.m[so + t] != INT32_MIN) sm.m[so + t] -= tmin;
      __syncwarp();

      // ---------------- main pixels: merged diffusion-bin groups, 3-template blend ----------------------------------
      const int ngx = sm.g_n[bxm], ngy = sm.g_n[bym];
      for (int gi = 0; gi < ngx; ++gi) {
        const int ox = (int)sm.g_ox[bxm][gi] - 1, cix = sm.g_ci[bxm][gi];
        const unsigned mx = sm.g_mask[bxm][gi];
        for (int gj = 0; gj < ngy; ++gj) {
          const int oy = (int)sm.g_ox[bym][gj] - 1, ciy = sm.g_ci[bym][gj];
          const unsigned my = sm.g_mask[bym][gj];
          int row = -1;
          if (active) {
            const int pid = pixel2id_dev(sm.mpx[lane] + ox, sm.mpy[lane] + oy, sm.ep[lane], S.nxp, S.nyp);
            row = lookup_row(lk, pid);      // not a main pixel: dropped (sim_jax.py:152-154)
            if (pid < 0) row = -1;          // garbage rows carry zero gradient here
          }
          if (__ballot_sync(0xffffffffu, row >= 0) == 0u) continue;
          const int bin = cix * 5 + ciy;
          // response rows of the unit -> shared memory, zero padded on both sides
          {
            const float* ra = S.rm + (int64_t)((idx - 1) * 25 + bin) * S.Lp + 2;
            const float* rb = S.rm + (int64_t)(idx * 25 + bin) * S.Lp + 2;
            const float* rc = S.rm + (int64_t)((idx + 1) * 25 + bin) * S.Lp + 2;
            __syncwarp();
            for (int k = lane; k < L + 2 * RPAD + KPT; k += 32) {
              const int kk = k - RPAD;
              const bool in = kk >= 0 && kk < L;
              sm.Rs[k] = in ? make_float4(__ldg(ra + kk), __ldg(rb + kk), __ldg(rc + kk), 0.0f) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          float* Gl = sm.Gs + lane * GS;
          switch (npos) {
            case 2: correlate_runs<2>(sm, A, lane, row, tmin, r_lo, r_hi, Gl); break;
            case 3: correlate_runs<3>(sm, A, lane, row, tmin, r_lo, r_hi, Gl); break;
            case 4: correlate_runs<4>(sm, A, lane, row, tmin, r_lo, r_hi, Gl); break;
            case 5: correlate_runs<5>(sm, A, lane, row, tmin, r_lo, r_hi, Gl); break;
            default: correlate_runs<6>(sm, A, lane, row, tmin, r_lo, r_hi, Gl); break;
          }
          // per segment of the lane's run: products with the segment's own factors
          if (row >= 0) {
            const float* grow = A.g + (int64_t)row * A.g_stride;
            const float* crow = S.cm + (int64_t)(idx * 25 + bin) * nt;
            const float Cl = __ldg(crow + nt - L);
            for (int t = 0; t < len; ++t) {
              const int i = so + t;
              const int m = sm.m[i];
              if (m == INT32_MIN) continue;
              const int64_t s = (int64_t)start + t;
              float gwx = 0.f, gwy = 0.f;
#pragma unroll
              for (int k = 0; k < 5; ++k) {
                if (mx >> k & 1) gwx += S.rec[(int64_t)(LARND_F_WX0 + k) * n + s];
                if (my >> k & 1) gwy += S.rec[(int64_t)(LARND_F_WY0 + k) * n + s];
              }
              const int T0 = tmin + m;
              const float gB = (T0 - 1 >= 1) ? __ldg(grow + T0 - 1) : 0.0f;  // corrections live on ticks >= 1
              const float gA = (T0 >= 1) ? __ldg(grow + T0) : 0.0f;
              int ct = nt - L - T0;
              ct = max(0, min(ct, nt - 1));
              const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, nt - 1));
              const float q = sm.q[i], f = sm.f[i], omf = 1.0f - f;
              const float ca_ = sm.ca[i], cb_ = sm.cb[i], cc_ = sm.cc[i];
              const float* Gm = Gl + 3 * m;
              const float a0 = Gm[0], b0 = Gm[1], c0v = Gm[2], a1 = Gm[3], b1 = Gm[4], c1v = Gm[5];
              const float D = Cl - (Ca * omf + Cb * f), dD = -(Cb - Ca);
              const float gm = fmaf(f, gB, omf * gA);
              const float Sa = fmaf(f, a0, omf * a1), Sb = fmaf(f, b0, omf * b1), Sc = fmaf(f, c0v, omf * c1v);
              const float Pv = fmaf(ca_, Sa, fmaf(cb_, Sb, cc_ * Sc)) + gm * D;
              const float Fd = fmaf(ca_, a0 - a1, fmaf(cb_, b0 - b1, cc_ * (c0v - c1v))) + (gB - gA) * D + gm * dD;
              const float w = gwx * gwy, qb = w * q;
              sm.acc[0][i] = fmaf(w, Pv, sm.acc[0][i]);
              sm.acc[1][i] = fmaf(qb, Fd, sm.acc[1][i]);
              sm.acc[2][i] = fmaf(qb, Sa, sm.acc[2][i]);
              sm.acc[3][i] = fmaf(qb, Sb, sm.acc[3][i]);
              sm.acc[4][i] = fmaf(qb, Sc, sm.acc[4][i]);
              sm.acc[5 + gi][i] += gwy * q * Pv;   // d/dWx_k for every member k of group gi
              sm.acc[10 + gj][i] += gwx * q * Pv;
            }
          }
        }
      }
      // ---------------- neighbour pixels that own a waveform row: template 0, full charge ------------------------------
      for (int u = 0; u < S.P * S.P; ++u) {
        const int dx = u / S.P - S.n_neigh, dy = u % S.P - S.n_neigh;
        if (dx == 0 && dy == 0) continue;  // centre id is -999: garbage row
        int row = -1;
        if (active) {
          const int pid = pixel2id_dev(sm.mpx[lane] + dx, sm.mpy[lane] + dy, sm.ep[lane], S.nxp, S.nyp);
          row = lookup_row(lk, pid);
          if (row <= 0 || pid < 0) row = -1;  // absent -> row 0, ids < 0: garbage rows, zero gradient
        }
        unsigned owned = __ballot_sync(0xffffffffu, row >= 0);
        if (owned == 0u) continue;
        const int vx = 2 * bxm - S.half2 - 2 * nb * dx, vy = 2 * bym - S.half2 - 2 * nb * dy;
        const int bin = (abs(vx) >> 1) * S.ny_lut + (abs(vy) >> 1);
        const float* rowp0 = S.r0 + (int64_t)bin * S.Lp;
        const float* crow = S.c0 + (int64_t)bin * nt;
        float Rw[NS][KPT];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int j = 0; j < KPT; ++j) {
            const int ix = 32 * s + lane + 1 - j;  // sample k = x - 1 - j lives at row[k + 2]
            Rw[s][j] = ((unsigned)ix < (unsigned)S.Lp) ? __ldg(rowp0 + ix) : 0.0f;
          }
        __syncwarp();
        while (owned) {  // lane <-> tick
          const int pr = __ffs(owned) - 1;
          owned &= owned - 1;
          const int rowp = __shfl_sync(0xffffffffu, row, pr);
          const int tminp = __shfl_sync(0xffffffffu, tmin, pr);
          const float* grow = A.g + (int64_t)rowp * A.g_stride;
          float part[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) part[j] = 0.0f;
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const int col = tminp - 1 + 32 * s + lane;
            const float gv = (col >= 2 && col <= S.nticks - 1) ? __ldg(grow + col) : 0.0f;
#pragma unroll
            for (int j = 0; j < KPT; ++j) part[j] = fmaf(gv, Rw[s][j], part[j]);
          }
          const float Gj = reduce8_to_lane_s(part, lane);
          if (lane < KPT) sm.Gs[pr * GS + lane] = Gj;
        }
        __syncwarp();
        if (row >= 0) {
          const float* grow = A.g + (int64_t)row * A.g_stride;
          const float Cl = __ldg(crow + nt - L);
          const float* Gl = sm.Gs + lane * GS;
          for (int t = 0; t < len; ++t) {
            const int i = so + t;
            const int m = sm.m[i];
            if (m == INT32_MIN) continue;
            const int T0 = tmin + m;
            const float gB = (T0 - 1 >= 1) ? __ldg(grow + T0 - 1) : 0.0f;
            const float gA = (T0 >= 1) ? __ldg(grow + T0) : 0.0f;
            int ct = nt - L - T0;
            ct = max(0, min(ct, nt - 1));
            const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, nt - 1));
            const float q = sm.q[i], f = sm.f[i], omf = 1.0f - f;
            const float G0 = Gl[m], G1 = Gl[m + 1];
            const float D = Cl - (Ca * omf + Cb * f), dD = -(Cb - Ca);
            const float gm = fmaf(f, gB, omf * gA);
            sm.acc[0][i] += fmaf(f, G0, omf * G1) + gm * D;
            sm.acc[1][i] += q * ((G0 - G1) + (gB - gA) * D + gm * dD);
          }
        }
        __syncwarp();
      }
      __syncwarp();
      // ---- K1b: chain rule through the per-segment preparation, lane <-> segment ------------------------------------
      int gmapx[5], gmapy[5];  // group of every transverse bin (uniform per class)
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        gmapx[k] = 0; gmapy[k] = 0;
        for (int g = 0; g < ngx; ++g) if (sm.g_mask[bxm][g] >> k & 1) gmapx[k] = g;
        for (int g = 0; g < ngy; ++g) if (sm.g_mask[bym][g] >> k & 1) gmapy[k] = g;
      }
      for (int i = lane; i < nseg; i += 32) {
        if (sm.m[i] == INT32_MIN) continue;
        const int64_t s = sm.sid[i];
        const float q = sm.q[i];
        const float dq = sm.acc[0][i], df = sm.acc[1][i], da = sm.acc[2][i], db = sm.acc[3][i], dc = sm.acc[4][i];
        float dwx[5], dwy[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { dwx[k] = sm.acc[5 + gmapx[k]][i]; dwy[k] = sm.acc[10 + gmapy[k]][i]; }
        const int flags = irec[(int64_t)LARND_I_FLAGS * n + s];
        const float sl = S.rec[(int64_t)LARND_F_SL * n + s];
        const float sT = S.rec[(int64_t)LARND_F_ST * n + s];
        const float td = S.rec[(int64_t)LARND_F_TD * n + s];
        const float x0 = S.rec[(int64_t)LARND_F_X0 * n + s], y0 = S.rec[(int64_t)LARND_F_Y0 * n + s];
        const float recb = S.rec[(int64_t)LARND_F_REC * n + s];
        const float ft = S.rec[(int64_t)LARND_F_FT * n + s];
        const float xi = S.rec[(int64_t)LARND_F_XI * n + s];
        const float cos2 = S.rec[(int64_t)LARND_F_COS2 * n + s];
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
        gacc[LARND_P_SHIFT_Z] += g_td * (-sgn_a / v) + g_ft * (-sgn_c / (v * ts));
        gacc[LARND_P_LIFETIME] += dq * q * td / (tau * tau);
        if (p.long_diff > 0.f) gacc[LARND_P_LONG_DIFF] += g_sl * sl / (2.0f * p.long_diff);
        if (p.tran_diff > 0.f) gacc[LARND_P_TRAN_DIFF] += g_sT * sT / (2.0f * p.tran_diff);
        gacc[LARND_P_SHIFT_X] += -g_x0;
        gacc[LARND_P_SHIFT_Y] += -g_y0;
        gacc[LARND_P_MEV_TO_ELECTRONS] += dq * q / p.MeVToElectrons;
        const float g_rec = (recb != 0.0f) ? dq * q / recb : 0.0f;  // q is linear in the recombination factor
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
      __syncwarp();
      r_lo = r_hi;
    }
  }
#pragma unroll
  for (int k = 0; k < LARND_NPARAMS; ++k) {
    const float v = warp_sum_f(gacc[k]);
    if (lane == 0) A.partials[(int64_t)blockIdx.x * 16 + k] = v;
  }
}

}  // namespace

int larnd_bwd_sorted_slots() { return LARND_BWD_SORTED_SLOTS; }

int larnd_launch_accumulate_bwd_sorted(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                       int32_t npix_capacity, int32_t flags, const float* g_wfs, int64_t g_stride,
                                       float* sorted_partials, int* n_slots_out, const int* gflag, const int32_t* counts,
                                       cudaStream_t st) {
  BwdSortArgs A;
  int rc = sorted_fill_and_build(A.S, n, p, lut, ws, npix_capacity, counts, st);
  if (rc) return rc;
  A.S.wfs = nullptr;
  A.S.skip_garbage = 1;
  A.g = g_wfs; A.g_stride = g_stride;
  A.partials = sorted_partials;
  A.force_skip = flags & 1;
  A.garbage_grad_nonzero = gflag;
  const size_t smem = sizeof(WarpSmem);
  static bool attr_done = false;
  if (!attr_done) {
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int grid = sorted_grid(8, LARND_BWD_SORTED_SLOTS);
  const int need = lut->L + 2 + SPAN_MAX_S;
  if (need <= 32 * 4) k_bwd_tiles<4><<<grid, 32, smem, st>>>(A, p);
  else if (need <= 32 * 5) k_bwd_tiles<5><<<grid, 32, smem, st>>>(A, p);
  else k_bwd_tiles<6><<<grid, 32, smem, st>>>(A, p);
  LARND_LAUNCH_CHECK("k_bwd_tiles");
  *n_slots_out = grid;
  return LARND_OK;
}
