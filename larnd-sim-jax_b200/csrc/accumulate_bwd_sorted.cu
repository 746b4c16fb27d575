// K3b' class-sorted VJP of simulate_wfs (reference: jax.grad through sim_jax.py:142-286,375-453), sm_100a —
// the large-batch path of larnd_lut_backward.  Same mathematics as accumulate_bwd.cu, different decomposition.
//
// Mirror image of accumulate_sorted.cu: runs are sorted by response class (sorted_runs.cuh); a CTA pulls a tile
// (<= 32 runs of one class, uniform number of impulse positions NPOS); its 8 warps pull units (merged diffusion-bin
// groups, neighbour pixels owning a non-garbage row) and keep the unit's response rows in REGISTERS.  For every
// (run, unit) pair the warp loads the upstream-gradient window of the target row (coalesced), correlates it with the
// register-resident response,
//       G_r[j] = sum_x g[row, tmin - 1 + x] * R_r[x - 1 - j],        j < NPOS, r < 3 (main) or 1 (neighbour),
// reduces the 3*NPOS lane-partials with 10-shuffle butterflies, and lanes <-> segments of the run turn G into d/dq,
// d/dfrac, d/d(a,b,c), d/dWx, d/dWy, added to per-segment accumulators in shared memory (float atomics: different warps
// work on different units of the same segments).  Threads <-> segments then apply the closed-form chain rule through
// drift / quench / diffusion (K1b); the 15 parameter gradients are block-reduced into partials summed in double.
// (A tensor-core form of the correlation — mma.sync m16n8k8 TF32 over the 32 runs of a tile — was measured on B200 and
// rejected: 55 ms vs 34 ms per 10 M segments, because the kernel is bound by the latency of the scattered gradient-row
// loads, not by issue slots, and TF32 loses ~1 % on d/d(long_diff), whose three template terms cancel to first order.)
// Work whose only effect is on garbage rows is skipped: their upstream gradient is zero, checked on the device by
// k_garbage_grad_flag.  If it is NOT zero, or for segments whose window ends beyond the readout, accumulate_bwd.cu's
// kernel does the work instead (device-side switch, no host synchronisation).
#include "bwd_chain.cuh"
#include "sorted_runs.cuh"

namespace {

constexpr int BT_THREADS = 256;        // dense-gradient variants (register-resident response: tuned launch bounds below)
#ifndef LARND_BT_THREADS_STEPS
#define LARND_BT_THREADS_STEPS 256
#endif
#ifndef LARND_BWD_STEPS_CTAS
#define LARND_BWD_STEPS_CTAS 4
#endif
constexpr int BT_THREADS_STEPS = LARND_BT_THREADS_STEPS;   // step-event variant
constexpr int BT_WARPS = (BT_THREADS_STEPS > BT_THREADS ? BT_THREADS_STEPS : BT_THREADS) / 32;   // per-warp buffers: the larger CTA
constexpr int SEGMAX_B = TR * MAXLEN;
constexpr int NACC = 15;           // dq, dfrac, da, db, dc, dWxg[5], dWyg[5]
constexpr int GSB = 3 * KPT + 6;   // per-warp G buffer (3*KPT values, padded to a multiple of 8)

struct BwdSortArgs {
  SortArgs S;
  const float* g;
  int64_t g_stride;
  float* partials;   // [gridDim][16]
  int force_skip;    // caller promises zero gradients on garbage rows
  const int* garbage_grad_nonzero;
  StepsView steps;   // compact upstream gradient (STEPS kernels): g / g_stride are unused then
};

constexpr int BSLOTS = 32 / MAXLEN;  // runs whose segments are processed together (lanes <-> (slot, segment))
struct PairSlots {
  float G[BSLOTS][GSB];        // G[NR*j + r] of the run parked in the slot
  float gpos[BSLOTS][KPT + 2]; // upstream gradient at ticks tmin - 1 + j
  int p[BSLOTS];
  __align__(16) int ev[BSLOTS][LARND_STEPS_WORDS];   // STEPS kernels: the step-event record of the slot's target row (cp.async)
};

struct BwdTileSmem {
  int4 run[TR];
  int ep[TR], mpx[TR], mpy[TR], soff[TR];
  float q[SEGMAX_B], f[SEGMAX_B], ca[SEGMAX_B], cb[SEGMAX_B], cc[SEGMAX_B];
  float wxg[5][SEGMAX_B], wyg[5][SEGMAX_B];  // transverse weights merged per group of the class
  int m[SEGMAX_B], sid[SEGMAX_B];            // T0 - tmin (INT32_MIN: outside every TPC, no gradient), global segment index
  unsigned char owner[SEGMAX_B];
  float acc[NACC][SEGMAX_B];
  PairSlots ps[BT_WARPS];
  float red[BT_WARPS][16];                   // per-warp parameter-gradient accumulators (warp-reduced once per tile)
  unsigned char g_n[16], g_ox[16][5], g_ci[16][5], g_mask[16][5];
  signed char udx[225], udy[225];
  int tile, next_unit, nseg;
  int low_end;  // some run of the tile starts below tick 2, or its 32 NS-tick register window ends beyond the row
};

__device__ __forceinline__ float warp_sum_f(float v) {  // [region: reduce helpers]
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

// sums 4 lane-partial values over the warp with 7 shuffles; the total of v[j] is returned in lane j (j < 4)
__device__ __forceinline__ float reduce4_to_lane_s(const float (&v)[4], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8;
  float w[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b4 ? v[i] : v[i + 2], keep = b4 ? v[i + 2] : v[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float x = (b3 ? w[1] : w[0]) + __shfl_xor_sync(0xffffffffu, b3 ? w[0] : w[1], 8);
  x += __shfl_xor_sync(0xffffffffu, x, 4);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return __shfl_sync(0xffffffffu, x, ((lane & 2) << 3) | ((lane & 1) << 3));  // value j lives in lanes with (b4, b3) = (j >> 1, j & 1)
}

// One unit of one tile: every run p of `todo` (bit mask of runs whose target row exists) is correlated with the unit's
// response rows Rw and its segments update the shared accumulators.  NR = 3: main pixel (3-template blend, group gi/gj);
// NR = 1: neighbour pixel (template 0, full charge).
template <int NS, int NR, int NPOS, int KP>
__device__ __forceinline__ void unit_pairs(BwdTileSmem& sm, const BwdSortArgs& A, const float (&Rw)[3][NS][KP], unsigned todo, int row,  // [region: unit_pairs setup]
                                           const float* __restrict__ crow, int gi, int gj, int lane, int warp) {
  const SortArgs& S = A.S;
  const int nt = S.nt, L = S.L, nticks = S.nticks;
  const float Cl = __ldg(crow + nt - L);
  PairSlots& ps = sm.ps[warp];
  // lane <-> run: signed 32-bit element offset of column tmin - 1 of the run's gradient row (npix * nticks < 2^31 is
  // checked by the launcher); tile-level flag instead of a per-run test for windows at the low end of the readout
  const int myoff = row * (int)A.g_stride + (sm.run[lane].z - 1);
  const bool inside = !sm.low_end;
  while (todo) {
    // ---- up to 4 runs: correlate, reduce, park the results in the warp's slots -------------------------------------  // [region: corr: pick run + g loads]
    int nslot = 0;
#pragma unroll 2
    for (; nslot < BSLOTS && todo; ++nslot) {
      const int p = __ffs(todo) - 1;
      todo &= todo - 1;
      const int tmin = sm.run[p].z;
      const float* gp = A.g + (__shfl_sync(0xffffffffu, myoff, p) + lane);
      // upstream-gradient window (coalesced): all loads first
      float graw[NS];
      if (inside) {  // tile-uniform, the common case: every register window of the tile lies inside its row
#pragma unroll
        for (int s = 0; s < NS; ++s) graw[s] = __ldg(gp + 32 * s);
      } else {
        const float* grow = gp - (tmin - 1) - lane;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const int col = tmin - 1 + 32 * s + lane;
          graw[s] = (col >= 1 && col <= nticks - 1) ? __ldg(grow + col) : 0.0f;
        }
      }
      if (lane <= KPT) ps.gpos[nslot][lane] = graw[0];  // lane j <-> position j: gradient at tick tmin - 1 + j
      if (lane == 0) ps.p[nslot] = p;
      float part[NR * NPOS];  // [region: corr: FMAs]
#pragma unroll
      for (int k = 0; k < NR * NPOS; ++k) part[k] = 0.0f;
      if (!inside) {  // window samples live on ticks >= 2, corrections on >= 1 (only differs for runs at the low end of the readout)
#pragma unroll
        for (int s = 0; s < NS; ++s)
          if (tmin - 1 + 32 * s + lane < 2) graw[s] = 0.0f;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float gv = graw[s];
#pragma unroll
        for (int j = 0; j < NPOS; ++j)
#pragma unroll
          for (int r = 0; r < NR; ++r) part[NR * j + r] = fmaf(gv, Rw[r][s][j], part[NR * j + r]);
      }
      // full groups of 8 partials with the 10-shuffle butterfly; the remainder with the cheapest one that fits  // [region: corr: reduce+park]
      // (<= 4 values: 7 shuffles, a single value: plain warp sum)
      constexpr int V = NR * NPOS, V8 = V / 8 * 8, REM = V - V8;
#pragma unroll
      for (int c0 = 0; c0 < V8; c0 += 8) {
        float v8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v8[k] = part[c0 + k];
        const float tot = reduce8_to_lane_s(v8, lane);
        if (lane < 8) ps.G[nslot][c0 + lane] = tot;  // G[NR*j + r]
      }
      if (REM > 4) {
        float v8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v8[k] = (k < REM) ? part[(k < REM) ? V8 + k : 0] : 0.0f;
        const float tot = reduce8_to_lane_s(v8, lane);
        if (lane < 8) ps.G[nslot][V8 + lane] = tot;
      } else if (REM > 1) {
        float v4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v4[k] = (k < REM) ? part[(k < REM) ? V8 + k : 0] : 0.0f;
        const float tot = reduce4_to_lane_s(v4, lane);
        if (lane < 4) ps.G[nslot][V8 + lane] = tot;
      } else if (REM == 1) {
        const float tot = warp_sum_f(part[V8 < V ? V8 : 0]);
        if (lane == 0) ps.G[nslot][V8] = tot;
      }
    }
    __syncwarp();
    // ---- lanes <-> (slot, segment): 4 runs x up to 8 segments in one pass --------------------------------------------  // [region: segment phase]
    const int slot = lane >> 3, t = lane & 7;
    if (slot < nslot) {
      const int p = ps.p[slot];
      const int len = sm.run[p].y & 0xffff;
      const int i = sm.soff[p] + t;
      const int m = t < len ? sm.m[i] : INT32_MIN;
      if (m != INT32_MIN) {
        const float* Gs = ps.G[slot];
        const float gB = ps.gpos[slot][m], gA = ps.gpos[slot][m + 1];   // ticks T0 - 1 and T0
        // running sums of the response at ct(T0), ct + 1 (sim_jax.py:236-247): loaded here, once per segment, instead of
        // per run and position in the correlation phase (where the shared-memory hand-over waited on the loads)
        int ct = nt - L - (sm.run[p].z + m);
        ct = max(0, min(ct, nt - 1));
        const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, nt - 1));
        const float q = sm.q[i], f = sm.f[i], omf = 1.0f - f;
        const float D = Cl - (Ca * omf + Cb * f), dD = -(Cb - Ca);
        const float gm = fmaf(f, gB, omf * gA);
        if (NR == 3) {
          const float a0 = Gs[3 * m], b0 = Gs[3 * m + 1], c0v = Gs[3 * m + 2], a1 = Gs[3 * m + 3], b1 = Gs[3 * m + 4], c1v = Gs[3 * m + 5];
          const float ca_ = sm.ca[i], cb_ = sm.cb[i], cc_ = sm.cc[i];
          const float gwx = sm.wxg[gi][i], gwy = sm.wyg[gj][i];
          const float Sa = fmaf(f, a0, omf * a1), Sb = fmaf(f, b0, omf * b1), Sc = fmaf(f, c0v, omf * c1v);
          const float Pv = fmaf(ca_, Sa, fmaf(cb_, Sb, cc_ * Sc)) + gm * D;
          const float Fd = fmaf(ca_, a0 - a1, fmaf(cb_, b0 - b1, cc_ * (c0v - c1v))) + (gB - gA) * D + gm * dD;
          const float w = gwx * gwy, qb = w * q;
          atomicAdd(&sm.acc[0][i], w * Pv);
          atomicAdd(&sm.acc[1][i], qb * Fd);
          atomicAdd(&sm.acc[2][i], qb * Sa);
          atomicAdd(&sm.acc[3][i], qb * Sb);
          atomicAdd(&sm.acc[4][i], qb * Sc);
          atomicAdd(&sm.acc[5 + gi][i], gwy * q * Pv);   // d/dWx_k for every member k of group gi
          atomicAdd(&sm.acc[10 + gj][i], gwx * q * Pv);
        } else {
          const float G0 = Gs[m], G1 = Gs[m + 1];
          atomicAdd(&sm.acc[0][i], fmaf(f, G0, omf * G1) + gm * D);
          atomicAdd(&sm.acc[1][i], q * ((G0 - G1) + (gB - gA) * D + gm * dD));
        }
      }
    }
    __syncwarp();
  }
}

// The same (unit, runs) work when the upstream gradient is the front end's step function (StepsView): the correlation
//       G_r[j] = sum_k g[row, tmin + j + k] R_r[k]          (k = sample of the L-sample response window)
// with g[col] = sum_{e: pos_e >= col} val_e collapses to one difference of the response's RUNNING SUM per event,
//       G_r[j] = sum_e val_e (C_r[min(khi, pos_e - tmin - j)] - C_r[klo - 1]),      C_r = cumulative response (lut cm / c0),
// so lane <-> (j, r) owns its G entry outright: no gradient loads, no FFMA sweep over the window, no butterfly reduction,
// no register-resident response.  klo / khi carry the readout limits (window deposits live on columns 2 .. n_ticks - 1).
template <int NR>
__device__ __forceinline__ void unit_pairs_steps(BwdTileSmem& sm, const BwdSortArgs& A, const float* const (&crows)[NR], unsigned todo, int row,
                                                 const float* __restrict__ crow, int gi, int gj, int lane, int warp, int npos) {
  const SortArgs& S = A.S;
  const int nt = S.nt, L = S.L, nticks = S.nticks, ntL = nt - L;
  const float Cl = __ldg(crow + ntL);
  PairSlots& ps = sm.ps[warp];
  const int V = NR * npos;
  // crr[k] below = running sum up to sample k of the response window, crr[-1] = everything before it
  // Several runs share the warp: a run needs V = NR * npos lanes for its G entries (and npos + 1 <= KPT + 1 lanes for the
  // boundary-deposit columns), so a main unit (NR = 3, V <= 15 for npos <= 5) takes a half warp and a neighbour unit (NR = 1)
  // a quarter; the events of the parked runs sit in shared memory and are read as broadcasts.
  constexpr int LPR = (NR == 3) ? 16 : 8, RPP = 32 / LPR;
  const bool narrow = V <= LPR;                  // warp-uniform (npos = 6 main units use the whole warp per run)
  const int sub = narrow ? lane / LPR : 0, ll = narrow ? lane % LPR : lane;
  const int jj = (NR == 3) ? ll / 3 : ll, rr = (NR == 3) ? ll - 3 * jj : 0;
  const float* crr = crows[0];
  if (NR == 3) crr = (rr == 0) ? crows[0] : ((rr == 1) ? crows[NR - 2] : crows[NR - 1]);
  crr += ntL;
  const bool mine2 = ll < V;
  const float cbase2 = mine2 ? __ldg(crr - 1) : 0.0f;
  while (todo) {
    // up to BSLOTS runs per pass: their event records (192 bytes each) go straight from global to shared memory with
    // cp.async (twelve 16-byte copies per record, no register staging), then the slots are consumed
    int nslot = 0;
    int nev[BSLOTS];
#pragma unroll
    for (int s = 0; s < BSLOTS; ++s) {
      if (todo) {
        const int pp = __ffs(todo) - 1;
        todo &= todo - 1;
        const int rowp = __shfl_sync(0xffffffffu, row, pp);
        if (lane < LARND_STEPS_WORDS / 4) {
          const unsigned dst = (unsigned)__cvta_generic_to_shared(&ps.ev[s][4 * lane]);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(A.steps.rec + (int64_t)rowp * LARND_STEPS_WORDS + 4 * lane)
                       : "memory");
        }
        if (lane == 0) ps.p[s] = pp;
        nslot = s + 1;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int s = 0; s < BSLOTS; ++s) nev[s] = (s < nslot) ? ps.ev[s][2 * LARND_STEPS_MAX] : 0;
    __syncwarp();
    const int npass = narrow ? (nslot + RPP - 1) / RPP : nslot;
    for (int pass = 0; pass < npass; ++pass) {
      const int s = narrow ? pass * RPP + sub : pass;
      const bool act = s < nslot;
      const int sx = act ? s : 0;
      int nemax = 0;
#pragma unroll
      for (int k = 0; k < BSLOTS; ++k)
        if (narrow ? (k / RPP == pass) : (k == pass)) nemax = max(nemax, nev[k]);
      const int tmin = sm.run[ps.p[sx]].z;
      const int klo = max(0, 2 - tmin - jj), kcap = min(L - 1, nticks - 1 - tmin - jj);
      const float clo = (klo > 0 && mine2) ? __ldg(crr + klo - 1) : cbase2;
      const int colp = tmin - 1 + ll;            // ll <-> position: gradient at column tmin - 1 + ll (boundary deposits)
      const bool colok = colp >= 1 && colp <= nticks - 1 && ll <= KPT;
      const int base = tmin + jj;
      float acc = 0.0f, gp = 0.0f;
      const int* evp = ps.ev[sx];
      const float* evv = reinterpret_cast<const float*>(ps.ev[sx] + LARND_STEPS_MAX);
#pragma unroll 2
      for (int e = 0; e < nemax; ++e) {
        const int pe = evp[e];
        const float ve = evv[e];
        const int khi = min(kcap, pe - base);
        if (mine2 && khi >= klo) acc = fmaf(ve, __ldg(crr + khi) - clo, acc);
        gp += (colok && pe >= colp) ? ve : 0.0f;
      }
      if (act && mine2) ps.G[s][ll] = acc;       // G[NR * j + r]
      if (act && ll <= KPT) ps.gpos[s][ll] = gp;
    }
    __syncwarp();
    const int slot = lane >> 3, t = lane & 7;
    if (slot < nslot) {
      const int p = ps.p[slot];
      const int len = sm.run[p].y & 0xffff;
      const int i = sm.soff[p] + t;
      const int m = t < len ? sm.m[i] : INT32_MIN;
      if (m != INT32_MIN) {
        const float* Gs = ps.G[slot];
        const float gB = ps.gpos[slot][m], gA = ps.gpos[slot][m + 1];
        int ct = nt - L - (sm.run[p].z + m);
        ct = max(0, min(ct, nt - 1));
        const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, nt - 1));
        const float q = sm.q[i], f = sm.f[i], omf = 1.0f - f;
        const float D = Cl - (Ca * omf + Cb * f), dD = -(Cb - Ca);
        const float gm = fmaf(f, gB, omf * gA);
        if (NR == 3) {
          const float a0 = Gs[3 * m], b0 = Gs[3 * m + 1], c0v = Gs[3 * m + 2], a1 = Gs[3 * m + 3], b1 = Gs[3 * m + 4], c1v = Gs[3 * m + 5];
          const float ca_ = sm.ca[i], cb_ = sm.cb[i], cc_ = sm.cc[i];
          const float gwx = sm.wxg[gi][i], gwy = sm.wyg[gj][i];
          const float Sa = fmaf(f, a0, omf * a1), Sb = fmaf(f, b0, omf * b1), Sc = fmaf(f, c0v, omf * c1v);
          const float Pv = fmaf(ca_, Sa, fmaf(cb_, Sb, cc_ * Sc)) + gm * D;
          const float Fd = fmaf(ca_, a0 - a1, fmaf(cb_, b0 - b1, cc_ * (c0v - c1v))) + (gB - gA) * D + gm * dD;
          const float w = gwx * gwy, qb = w * q;
          atomicAdd(&sm.acc[0][i], w * Pv);
          atomicAdd(&sm.acc[1][i], qb * Fd);
          atomicAdd(&sm.acc[2][i], qb * Sa);
          atomicAdd(&sm.acc[3][i], qb * Sb);
          atomicAdd(&sm.acc[4][i], qb * Sc);
          atomicAdd(&sm.acc[5 + gi][i], gwy * q * Pv);
          atomicAdd(&sm.acc[10 + gj][i], gwx * q * Pv);
        } else {
          const float G0 = Gs[m], G1 = Gs[m + 1];
          atomicAdd(&sm.acc[0][i], fmaf(f, G0, omf * G1) + gm * D);
          atomicAdd(&sm.acc[1][i], q * ((G0 - G1) + (gB - gA) * D + gm * dD));
        }
      }
    }
    __syncwarp();
  }
}

template <int NS, int NR, int KP>  // [region: dispatch]
__device__ __forceinline__ void unit_pairs_npos(BwdTileSmem& sm, const BwdSortArgs& A, const float (&Rw)[3][NS][KP], unsigned todo, int row,
                                                const float* crow, int gi, int gj, int lane, int warp, int npos) {
  constexpr int N3 = KP >= 3 ? 3 : KP, N4 = KP >= 4 ? 4 : KP, N5 = KP >= 5 ? 5 : KP;
  if (KP >= 6 && npos >= 6) unit_pairs<NS, NR, KP, KP>(sm, A, Rw, todo, row, crow, gi, gj, lane, warp);
  else if (KP >= 5 && npos == 5) unit_pairs<NS, NR, N5, KP>(sm, A, Rw, todo, row, crow, gi, gj, lane, warp);
  else if (KP >= 4 && npos == 4) unit_pairs<NS, NR, N4, KP>(sm, A, Rw, todo, row, crow, gi, gj, lane, warp);
  else if (KP >= 3 && npos == 3) unit_pairs<NS, NR, N3, KP>(sm, A, Rw, todo, row, crow, gi, gj, lane, warp);
  else unit_pairs<NS, NR, 2, KP>(sm, A, Rw, todo, row, crow, gi, gj, lane, warp);
}

template <int NS, int NR, int KP>  // [region: load_response]
__device__ __forceinline__ void load_response_b(float (&Rw)[3][NS][KP], const float* const (&rows)[NR], int Lp, int lane) {
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        const int ix = 32 * s + lane + 1 - j;  // sample k = x - 1 - j lives at row[k + 2]
        Rw[r][s][j] = ((unsigned)ix < (unsigned)Lp) ? __ldg(rows[r] + ix) : 0.0f;
      }
}

// KP / span range / launch as in k_acc_tiles: the variants holding 3 / 4 response positions (36 / 48 registers, 4 / 3 CTAs
// per SM) serve the tiles of runs with few impulse positions, the KP = KPT kernel the rest (or everything).
template <int NS, int KP, bool STEPS = false>  // [region: kernel prologue]
__global__ void __launch_bounds__(STEPS ? BT_THREADS_STEPS : BT_THREADS,
                                  STEPS ? LARND_BWD_STEPS_CTAS : (NS <= 4 ? (KP <= 3 ? 4 : (KP <= 4 ? 3 : 2)) : (NS == 5 && KP <= 4 ? 2 : 1)))
k_bwd_tiles(const __grid_constant__ BwdSortArgs A, const __grid_constant__ larnd_params_t p, const int span_lo, const int span_hi,
            const int launch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdTileSmem& sm = *reinterpret_cast<BwdTileSmem*>(smem_raw);
  const SortArgs& S = A.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < 16) sm.red[warp][lane] = 0.0f;
  const bool garbage_needed = !STEPS && !A.force_skip && (*A.garbage_grad_nonzero != 0);
  const bool dead = S.counts[2] != 0 || garbage_needed;  // accumulate_bwd.cu's kernel takes over
  const int nb = S.nb;
  const int64_t n = S.n;
  const int* irec = reinterpret_cast<const int*>(S.rec);
  RowLookup lk = S.lk;
  lk.n_unique = S.counts[0];
  lk.n_neg = S.counts[1];
  if (threadIdx.x < nb && threadIdx.x < 16) build_bin_groups(threadIdx.x, nb, S.half2, sm.g_n[threadIdx.x], sm.g_ox[threadIdx.x], sm.g_ci[threadIdx.x], sm.g_mask[threadIdx.x]);
  constexpr int NTHR = STEPS ? BT_THREADS_STEPS : BT_THREADS;
  for (int u = threadIdx.x; u < S.P * S.P; u += NTHR) {
    sm.udx[u] = (signed char)(u / S.P - S.n_neigh);
    sm.udy[u] = (signed char)(u % S.P - S.n_neigh);
  }
  const int tile_lo = S.gcnt[GC_SPAN + span_lo];
  const int ntiles = dead ? 0 : S.gcnt[GC_SPAN + span_hi + 1];
  int* const tile_counter = S.gcnt + GC_BWD + launch;
  const int n_units = 25 + S.P * S.P;

  for (;;) {  // [region: tile pull]
    __syncthreads();
    if (threadIdx.x == 0) { sm.tile = tile_lo + atomicAdd(tile_counter, 1); sm.next_unit = 0; }
    __syncthreads();
    const int tile = sm.tile;
    if (tile >= ntiles) break;
    const int4 ti = S.tile_info[tile];
    const int cls = ti.x, count = ti.z;
    const int ncb = S.ncls / (SPAN_MAX_S + 1);
    const int span = cls / ncb, cls_b = cls % ncb;  // class = span * (ntpl * nb * nb) + (idx * nb + bxm) * nb + bym
    const int bym = cls_b % nb, bxm = (cls_b / nb) % nb, idx = cls_b / (nb * nb);
    const int npos = span + 2;
    // ---- stage the runs (warp 0: lane <-> run) ---------------------------------------------------------------  // [region: stage runs]
    if (warp == 0) {
      int len = 0;
      if (lane < count) {
        const int4 e = S.runs[ti.y + lane];
        sm.run[lane] = e;
        const int64_t s0 = e.x;
        sm.ep[lane] = irec[(int64_t)LARND_I_EP * n + s0];
        sm.mpx[lane] = floordiv_i(irec[(int64_t)LARND_I_BX * n + s0], nb);
        sm.mpy[lane] = floordiv_i(irec[(int64_t)LARND_I_BY * n + s0], nb);
        len = e.y & 0xffff;
      }
      const unsigned low = STEPS ? 0u : __ballot_sync(0xffffffffu, lane < count && (sm.run[lane].z < 2 || sm.run[lane].z - 2 + 32 * NS >= S.nticks));
      if (lane == 0) sm.low_end = low != 0u;
      int inc = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      const int off = inc - len;
      sm.soff[lane] = off;
      for (int t = 0; t < len; ++t) sm.owner[off + t] = (unsigned char)lane;
      if (lane == 31) sm.nseg = inc;
    }
    __syncthreads();
    // ---- stage the segments (thread <-> segment) -----------------------------------------------------------------  // [region: stage segments]
    const int nseg = sm.nseg;
    for (int i = threadIdx.x; i < nseg; i += NTHR) {
      const int r = sm.owner[i];
      const int4 e = sm.run[r];
      const int64_t s = (int64_t)e.x + (i - sm.soff[r]);
      sm.sid[i] = (int)s;
      sm.q[i] = S.rec[(int64_t)LARND_F_Q * n + s];
      sm.f[i] = S.rec[(int64_t)LARND_F_FRAC * n + s];
      sm.ca[i] = S.rec[(int64_t)LARND_F_A * n + s];
      sm.cb[i] = S.rec[(int64_t)LARND_F_B * n + s];
      sm.cc[i] = S.rec[(int64_t)LARND_F_C * n + s];
      // segments outside every TPC carry q == 0 and dq/dtheta == 0 (mask factor): no gradient
      const bool inside = irec[(int64_t)LARND_I_FLAGS * n + s] & 1;
      sm.m[i] = inside ? irec[(int64_t)LARND_I_T0 * n + s] - e.z : INT32_MIN;
      float vx[5], vy[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        vx[k] = S.rec[(int64_t)(LARND_F_WX0 + k) * n + s];
        vy[k] = S.rec[(int64_t)(LARND_F_WY0 + k) * n + s];
      }
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        const int mxg = g < sm.g_n[bxm] ? sm.g_mask[bxm][g] : 0, myg = g < sm.g_n[bym] ? sm.g_mask[bym][g] : 0;
        float sx = 0.0f, sy = 0.0f;
#pragma unroll
        for (int k = 0; k < 5; ++k) { sx += (mxg >> k & 1) ? vx[k] : 0.0f; sy += (myg >> k & 1) ? vy[k] : 0.0f; }
        sm.wxg[g][i] = sx;
        sm.wyg[g][i] = sy;
      }
#pragma unroll
      for (int k = 0; k < NACC; ++k) sm.acc[k][i] = 0.0f;
    }
    __syncthreads();

    for (;;) {  // [region: unit pull]
      int unit = 0;
      if (lane == 0) unit = atomicAdd(&sm.next_unit, 1);
      unit = __shfl_sync(0xffffffffu, unit, 0);
      if (unit >= n_units) break;
      float Rw[STEPS ? 1 : 3][STEPS ? 1 : NS][STEPS ? 1 : KP];
      if (unit < 25) {  // [region: main unit setup]
        // ---------------- merged diffusion-bin group (gi, gj): 3-template blend on a main pixel ----------------
        const int gi = unit / LARND_NB_TRAN_BINS, gj = unit % LARND_NB_TRAN_BINS;
        if (gi >= sm.g_n[bxm] || gj >= sm.g_n[bym]) continue;
        const int bin = (int)sm.g_ci[bxm][gi] * 5 + (int)sm.g_ci[bym][gj];
        const int ox = (int)sm.g_ox[bxm][gi] - 1, oy = (int)sm.g_ox[bym][gj] - 1;
        int row = -1;
        if (lane < count) {
          const int pid = pixel2id_dev(sm.mpx[lane] + ox, sm.mpy[lane] + oy, sm.ep[lane], S.nxp, S.nyp);
          row = lookup_row(lk, pid);  // not a main pixel: dropped (sim_jax.py:152-154)
          if (pid < 0) row = -1;      // garbage rows carry zero gradient here
        }
        const unsigned todo = __ballot_sync(0xffffffffu, row >= 0);
        if (todo == 0u) continue;
        if constexpr (STEPS) {
          const float* const crows[3] = {S.cm + (int64_t)((idx - 1) * 25 + bin) * S.nt, S.cm + (int64_t)(idx * 25 + bin) * S.nt,
                                         S.cm + (int64_t)((idx + 1) * 25 + bin) * S.nt};
          unit_pairs_steps<3>(sm, A, crows, todo, row, crows[1], gi, gj, lane, warp, npos);
        } else {
          const float* const rows[3] = {S.rm + (int64_t)((idx - 1) * 25 + bin) * S.Lp, S.rm + (int64_t)(idx * 25 + bin) * S.Lp,
                                        S.rm + (int64_t)((idx + 1) * 25 + bin) * S.Lp};
          load_response_b<NS, 3, KP>(Rw, rows, S.Lp, lane);
          unit_pairs_npos<NS, 3, KP>(sm, A, Rw, todo, row, S.cm + (int64_t)(idx * 25 + bin) * S.nt, gi, gj, lane, warp, npos);
        }
      } else {  // [region: neigh unit setup]
        // ---------------- neighbour pixels that own a non-garbage waveform row: template 0, full charge -----------
        const int u = unit - 25;
        const int dx = sm.udx[u], dy = sm.udy[u];
        if (dx == 0 && dy == 0) continue;  // centre id is -999: garbage row
        int row = -1;
        if (lane < count) {
          const int pid = pixel2id_dev(sm.mpx[lane] + dx, sm.mpy[lane] + dy, sm.ep[lane], S.nxp, S.nyp);
          row = lookup_row(lk, pid);
          if (row <= 0 || pid < 0) row = -1;  // absent -> row 0, ids < 0: garbage rows, zero gradient
        }
        const unsigned todo = __ballot_sync(0xffffffffu, row >= 0);
        if (todo == 0u) continue;
        const int vx = 2 * bxm - S.half2 - 2 * nb * dx, vy = 2 * bym - S.half2 - 2 * nb * dy;
        const int bin = (abs(vx) >> 1) * S.ny_lut + (abs(vy) >> 1);
        if constexpr (STEPS) {
          const float* const crows[1] = {S.c0 + (int64_t)bin * S.nt};
          unit_pairs_steps<1>(sm, A, crows, todo, row, crows[0], 0, 0, lane, warp, npos);
        } else {
          const float* const rows[1] = {S.r0 + (int64_t)bin * S.Lp};
          load_response_b<NS, 1, KP>(Rw, rows, S.Lp, lane);
          unit_pairs_npos<NS, 1, KP>(sm, A, Rw, todo, row, S.c0 + (int64_t)bin * S.nt, 0, 0, lane, warp, npos);
        }
      }
    }
    __syncthreads();  // [region: chain rule]
    // ---- K1b: chain rule through the per-segment preparation, thread <-> segment -------------------------------------
    int gmapx[5], gmapy[5];  // group of every transverse bin (uniform per class)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      gmapx[k] = 0; gmapy[k] = 0;
      for (int g = 0; g < sm.g_n[bxm]; ++g) if (sm.g_mask[bxm][g] >> k & 1) gmapx[k] = g;
      for (int g = 0; g < sm.g_n[bym]; ++g) if (sm.g_mask[bym][g] >> k & 1) gmapy[k] = g;
    }
    float gacc[LARND_NPARAMS];   // this thread's parameter gradients of the tile; live only during the chain-rule phase
#pragma unroll
    for (int k = 0; k < LARND_NPARAMS; ++k) gacc[k] = 0.0f;
    for (int i = threadIdx.x; i < nseg; i += NTHR) {
      if (sm.m[i] == INT32_MIN) continue;
      const int64_t s = sm.sid[i];
      const float q = sm.q[i];
      const float dq = sm.acc[0][i], df = sm.acc[1][i], da = sm.acc[2][i], db = sm.acc[3][i], dc = sm.acc[4][i];
      float dwx[5], dwy[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) { dwx[k] = sm.acc[5 + gmapx[k]][i]; dwy[k] = sm.acc[10 + gmapy[k]][i]; }
      chain_rule_segment(p, S.rec, n, s, idx, q, dq, df, da, db, dc, dwx, dwy, [&](int k, float v) { gacc[k] += v; });
    }
    if (warp * 32 < nseg) {      // warps without a segment of this tile have nothing to add (warp-uniform)
#pragma unroll
      for (int k = 0; k < LARND_NPARAMS; ++k) {
        const float v = warp_sum_f(gacc[k]);
        if (lane == 0) sm.red[warp][k] += v;
      }
    }
  }
  // ---- block reduction -> per-CTA partials -------------------------------------------------------------------------  // [region: final reduce]
  __syncthreads();
  if (threadIdx.x < LARND_NPARAMS) {
    float v = 0.f;
    for (int w = 0; w < NTHR / 32; ++w) v += sm.red[w][threadIdx.x];
    A.partials[(int64_t)blockIdx.x * 16 + threadIdx.x] = v;
  }
}

}  // namespace

int larnd_launch_accumulate_bwd_sorted(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                       int32_t npix_capacity, int32_t flags, const float* g_wfs, int64_t g_stride,
                                       float* sorted_partials, int* n_slots_out, const int* gflag, const int32_t* counts,
                                       cudaStream_t st, const StepsView* steps) {
  BwdSortArgs A;
  int rc = sorted_fill_and_build(A.S, n, p, lut, ws, npix_capacity, counts, st, (flags & LARND_FLAG_REUSE_RUNS) != 0);
  if (rc) return rc;
  if (steps) {
    // compact upstream gradient: one kernel for every tick span (no register-resident response, so no KP variants)
    A.S.wfs = nullptr;
    A.S.skip_garbage = 1;
    A.g = nullptr; A.g_stride = 0;
    A.partials = sorted_partials;
    A.force_skip = 1;
    A.garbage_grad_nonzero = gflag;
    A.steps = *steps;
    const size_t smem_s = sizeof(BwdTileSmem);
    static bool attr_steps = false;
    if (!attr_steps) {
      LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
      attr_steps = true;
    }
    const int grid = sorted_grid(LARND_BWD_STEPS_CTAS, LARND_BWD_SORTED_SLOTS);
    k_bwd_tiles<4, 2, true><<<grid, BT_THREADS_STEPS, smem_s, st>>>(A, p, 0, SPAN_MAX_S, 2);
    LARND_LAUNCH_CHECK("k_bwd_tiles<steps>");
    *n_slots_out = grid;
    return LARND_OK;
  }
  A.S.wfs = nullptr;
  A.S.skip_garbage = 1;
  A.g = g_wfs; A.g_stride = g_stride;
  A.partials = sorted_partials;
  A.force_skip = flags & 1;
  A.garbage_grad_nonzero = gflag;
  const size_t smem = sizeof(BwdTileSmem);
  static bool attr_done = false;
  if (!attr_done) {
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<4, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<5, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<6, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_bwd_tiles<5, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int need = lut->L + 2 + SPAN_MAX_S;
  const int split = need <= 32 * 4 ? sorted_split_mode(flags) : 0;
  // every launch gets its full residency (2 / 3 / 4 CTAs per SM); only the optional third variant is capped by what is left
  // of the per-CTA partial table
  const int grid2 = sorted_grid(2, LARND_BWD_SORTED_SLOTS / 3);
  const int grid3 = split >= 1 ? sorted_grid(3, LARND_BWD_SORTED_SLOTS / 2 - grid2 / 2) : 0;
  const int grid4 = split >= 2 ? sorted_grid(4, LARND_BWD_SORTED_SLOTS - grid2 - grid3) : 0;
  // the kernels write disjoint slots of the per-CTA partial table
  int big_lo = 0;
  if (split >= 2) {
    k_bwd_tiles<4, 3><<<grid4, BT_THREADS, smem, st>>>(A, p, 0, 1, 0);
    LARND_LAUNCH_CHECK("k_bwd_tiles<3>");
    A.partials += (int64_t)grid4 * 16;
    big_lo = 2;
  }
  if (split >= 1) {
    k_bwd_tiles<4, 4><<<grid3, BT_THREADS, smem, st>>>(A, p, big_lo, 2, 1);
    LARND_LAUNCH_CHECK("k_bwd_tiles<4>");
    A.partials += (int64_t)grid3 * 16;
    big_lo = 3;
  }
  if (need <= 32 * 4) k_bwd_tiles<4, KPT><<<grid2, BT_THREADS, smem, st>>>(A, p, big_lo, SPAN_MAX_S, 2);
  else if (need <= 32 * 5) {
    int n5 = 0;
    if (sorted_split_mode(flags) >= 1) {  // 150-tick windows: 60 response registers, 2 CTAs/SM instead of 1
      k_bwd_tiles<5, 4><<<grid2, BT_THREADS, smem, st>>>(A, p, 0, 2, 1);
      LARND_LAUNCH_CHECK("k_bwd_tiles<5,4>");
      A.partials += (int64_t)grid2 * 16;
      big_lo = 3;
      n5 = grid2;
    }
    k_bwd_tiles<5, KPT><<<grid2, BT_THREADS, smem, st>>>(A, p, big_lo, SPAN_MAX_S, 2);
    LARND_LAUNCH_CHECK("k_bwd_tiles");
    *n_slots_out = grid2 + n5;
    return LARND_OK;
  }
  else k_bwd_tiles<6, KPT><<<grid2, BT_THREADS, smem, st>>>(A, p, 0, SPAN_MAX_S, 2);
  LARND_LAUNCH_CHECK("k_bwd_tiles");
  *n_slots_out = grid2 + grid3 + grid4;
  return LARND_OK;
}
