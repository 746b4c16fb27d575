// K3b + K1b: VJP of simulate_wfs (reference sim_jax.py:689-736) w.r.t. the fitted Params leaves, sm_100a.
//
// The reference obtains this by jax.grad through gather / scatter-add / erf / sqrt (optimize/fit_params.py:731).
// Here the backward pass is the *gather* form of K3 and uses the same run decomposition: consecutive segments that
// share (event, plane, sub-pixel bin, template index) read the same response rows, so for every target row the
// kernel first correlates the upstream gradient window with the response row once per tick position of the run,
//     G[j] = sum_x g[row, tmin + j + x] * R[x],
// and then every segment of the run (one lane per segment) picks its two positions:
//     d/dq += f G[m] + (1-f) G[m+1] + boundary term,      d/dfrac += q (G[m] - G[m+1]) + boundary terms,
// plus, for the 25 diffusion bins, the derivatives w.r.t. the Lagrange weights (a,b,c) and the diffusion weights
// Wx[5], Wy[5].  Everything else on the path is an integer index with zero gradient.  A warp owns a
// run; after walking the 25 + (2n+1)^2 target rows each lane applies the closed-form chain rule through drift /
// quench / diffusion-weight math for its own segment (drifting_jax.py:42-50, quenching_jax.py:18-35,
// detsim_jax.py:332-341, sim_jax.py:157-168,406-423).  The LARND_NPARAMS parameter gradients are reduced per CTA,
// written as per-chunk partials and summed in double by a second tiny kernel (deterministic, no float atomics).
#include <stdlib.h>

#include "bwd_chain.cuh"
#include "larnd_common.cuh"

namespace {

constexpr int BWD_THREADS = 256;
constexpr int BWD_WARPS = BWD_THREADS / 32;
constexpr int S = LARND_CHUNK;
constexpr int KP = 8;
constexpr int SPAN_MAX = KP - 2;

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
  const int* garbage_grad_nonzero;  // device flag written by k_garbage_grad_flag
  StepsView steps;   // compact upstream gradient (larnd_fee_backward_steps) used instead of g when use_steps != 0
  int use_steps;
  int chunk;          // segments per CTA (<= S, larnd_chunk_size)
  int sorted_active;  // the class-sorted kernel (accumulate_bwd_sorted.cu) runs too and takes every segment it can handle
  const int* n_slow;  // sorted_active: device count of the segments it leaves to this kernel (k_build_runs)
};

constexpr int MAXK = 16;        // distinct main pixels per chunk served from the shared row table
constexpr int TWMAX = 15;       // table width 2*max(n,1)+1 for n <= 7

struct BRun {
  int start, len, tmin, span;
  int key;                      // index into the key table, or -1 (table full: look rows up directly)
  int mpx, mpy, ep;
};

struct BwdSmem {
  float4 seg[S];  // q, frac, T0 (int bits), unused
  int ep[S], bx[S], by[S], idx[S], flags[S];
  float a[S], b[S], c[S];
  float wx[5][S], wy[5][S];
  BRun run[S];
  int nruns, nkeys;
  int next_run;
  int kep[MAXK], kpx[MAXK], kpy[MAXK];
  int krow[MAXK][TWMAX * TWMAX];            // raw row of pixel (mpx+dx, mpy+dy): -1 absent, bit 30 = id < 0
  unsigned char klist[MAXK][TWMAX * TWMAX]; // neighbour units (index into the P x P stencil) that have work
  int kcount[MAXK];
  // transverse-diffusion bins that share pixel and response row are merged ("groups"), per in-pixel bin index b:
  unsigned char g_n[16], g_ox[16][5], g_ci[16][5], g_mask[16][5];
  float dwx[BWD_WARPS][5][32], dwy[BWD_WARPS][5][32];
  float grad[BWD_WARPS][16];
  float gl[LARND_NPARAMS][BWD_THREADS];  // per-thread parameter-gradient accumulators (kept out of the register file)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums 8 lane-partial values over the warp with 10 shuffles (instead of 40): after three halving exchanges every lane
// holds one partially reduced value, two more butterflies finish it; the total of v[j] is returned in lane j (j < 8).
__device__ __forceinline__ float reduce8_to_lane(const float (&v)[8], int lane) {
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
  // lane l now holds the total of value index (l >> 2) & 7
  return __shfl_sync(0xffffffffu, y, (lane & 7) * 4);
}

// One row of the upstream gradient: either a pointer into the dense (npix, n_ticks) array or the row's step events
// (larnd_fee_backward_steps), from which any column is synthesised in registers — no gradient traffic at all.
struct GRow {
  const float* grow;
  RowSteps ev;
  bool steps;
  __device__ __forceinline__ void open(const BwdArgs& A, int row, int lane) {
    steps = A.use_steps != 0;
    if (steps) { ev.load(A.steps, row, lane); ev.fix(lane); }
    else grow = A.g + (int64_t)row * A.g_stride;
  }
  // value at this lane's column (warp-collective in the steps form); `ok` masks lanes whose column must read as zero
  __device__ __forceinline__ float at(int col, bool ok) const {
    if (steps) { const float v = ev.value(col); return ok ? v : 0.0f; }
    return ok ? __ldg(grow + col) : 0.0f;
  }
};

// gradient window registers: gr[i] = g[row, tmin + lane + 32 i]; ticks beyond the readout read as zero
template <int NG>
__device__ __forceinline__ void load_gwin(float (&gr)[NG], const GRow& G, int tmin, int nticks, int lane) {
#pragma unroll
  for (int i = 0; i < NG; ++i) {
    const int col = tmin + lane + 32 * i;
    gr[i] = G.at(col, col <= nticks - 1);
  }
}

// G[j] = sum_x g[tmin + j + x] R[x], j < npos <= 8, for one response row; the total for position j ends up in lane j
template <int NG>
__device__ __forceinline__ float correlate(const float (&gr)[NG], const float* rowp, int npos, int L, int lane) {
  float part[8];
  const float* p = rowp + 2 + lane;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    part[j] = 0.0f;
    if (j < npos) {  // warp-uniform
#pragma unroll
      for (int i = 0; i < NG; ++i) {
        const bool in = (unsigned)(lane - j + 32 * i) < (unsigned)L;
        const float v = in ? __ldg(p + (32 * i - j)) : 0.0f;
        part[j] = fmaf(gr[i], v, part[j]);
      }
    }
  }
  return reduce8_to_lane(part, lane);
}

// per-segment (slow path, runs touching the ends of the readout): lane-partial sums of one (segment, row) pair.
template <int NR>
__device__ __forceinline__ void slow_sums(const GRow& G, const float* const (&rows)[NR], int T0, int L, int nticks, int lane,
                                          float (&G0)[NR], float (&G1)[NR], float& gB, float& gA) {
#pragma unroll
  for (int r = 0; r < NR; ++r) { G0[r] = 0.f; G1[r] = 0.f; }
  gB = 0.f; gA = 0.f;
  for (int x0 = -1; x0 <= L; x0 += 32) {   // warp-uniform trip count (the steps form of G.at is warp-collective)
    const int x = x0 + lane;
    const int col = T0 + x;
    const bool inb = x <= L && col >= 1 && col <= nticks - 1;
    const float gv = G.at(col, inb);
    if (x > L) continue;
    const float gw = (col >= 2) ? gv : 0.0f;               // window deposits are valid for ticks 2 .. nticks-1
    const float gc = (col <= nticks - 2) ? gv : 0.0f;      // boundary deposits are valid for ticks 1 .. nticks-2
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      G0[r] = fmaf(gw, __ldg(rows[r] + x + 2), G0[r]);
      G1[r] = fmaf(gw, __ldg(rows[r] + x + 1), G1[r]);
    }
    if (x == -1) gB = gc;
    if (x == 0) gA = gc;
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) { G0[r] = warp_sum(G0[r]); G1[r] = warp_sum(G1[r]); }
  gB = warp_sum(gB);
  gA = warp_sum(gA);
}

__device__ __forceinline__ int raw_row(const RowLookup& lk, int px, int py, int ep, int nxp, int nyp) {
  const int pid = pixel2id_dev(px, py, ep, nxp, nyp);
  const int row = lookup_row(lk, pid);
  return row < 0 ? -1 : (row | (pid < 0 ? (1 << 30) : 0));
}

// Rows 0 .. n_neg of the waveform buffer belong to pixel ids < 0 (padding events and the -1 entries): the reference's own
// callers never look at them (parse_output, sim_jax.py:621), so their upstream gradient is zero for every loss built on
// hits.  This kernel checks that on the device; when it holds, all work whose only effect is on those rows is skipped.
__global__ void k_garbage_grad_flag(const float* __restrict__ g, int64_t g_stride, int nticks, const int32_t* __restrict__ counts,
                                    int* __restrict__ flag) {
  const int n_rows = counts[1] + 1;
  bool nz = false;
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x)
    for (int t = 1 + threadIdx.x; t < nticks; t += blockDim.x) nz |= g[(int64_t)r * g_stride + t] != 0.0f;
  if (__syncthreads_or(nz) && threadIdx.x == 0) atomicOr(flag, 1);
}

#ifndef LARND_BWD_CHUNK_CTAS
#define LARND_BWD_CHUNK_CTAS 3
#endif
template <int NG>
__global__ void __launch_bounds__(BWD_THREADS, LARND_BWD_CHUNK_CTAS)
k_lut_backward(const __grid_constant__ BwdArgs A, const __grid_constant__ larnd_params_t p) {
  const bool skip_garbage = A.skip_garbage || (*A.garbage_grad_nonzero == 0);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool bad = A.counts[2] != 0;
  const int64_t s_base = (int64_t)blockIdx.x * A.chunk;
  const int ns = bad ? 0 : (int)min((int64_t)A.chunk, A.n - s_base);
  const int nb = A.nb, L = A.L, nt = A.nt;
  const int64_t n = A.n;
  const int* irec = reinterpret_cast<const int*>(A.rec);
  const int tm = max(A.n_neigh, 1), TW = 2 * tm + 1;  // row table covers the (2 tm + 1)^2 pixels around a main pixel
  RowLookup lk = A.lk;
  lk.n_unique = A.counts[0];
  lk.n_neg = A.counts[1];
  // With the class-sorted kernel active and zero garbage-row gradients, only segments whose window ends beyond the
  // readout are left for this kernel (none for the usual Nt/L); otherwise it does everything and the sorted kernel idles.
  const bool slow_only = A.sorted_active && skip_garbage;
  if (slow_only) {
    int slow = 0;
    const bool none = A.n_slow && *A.n_slow == 0;   // block-uniform: no boundary segment in the whole batch
    if (!none && (int)threadIdx.x < ns) slow = !seg_is_fast(irec[(int64_t)LARND_I_T0 * n + s_base + threadIdx.x], L, A.nticks);
    if (none || !__syncthreads_or(slow)) {
      if (threadIdx.x < LARND_NPARAMS) A.partials[(int64_t)blockIdx.x * 16 + threadIdx.x] = 0.0f;
      return;
    }
  }
  // ---- stage the chunk ----------------------------------------------------------------------------------
  for (int t = threadIdx.x; t < ns; t += BWD_THREADS) {
    const int64_t s = s_base + t;
    sm.seg[t] = make_float4(A.rec[(int64_t)LARND_F_Q * n + s], A.rec[(int64_t)LARND_F_FRAC * n + s],
                            __int_as_float(irec[(int64_t)LARND_I_T0 * n + s]), 0.0f);
    sm.ep[t] = irec[(int64_t)LARND_I_EP * n + s];
    sm.bx[t] = irec[(int64_t)LARND_I_BX * n + s];
    sm.by[t] = irec[(int64_t)LARND_I_BY * n + s];
    sm.idx[t] = irec[(int64_t)LARND_I_IDX * n + s];
    sm.flags[t] = irec[(int64_t)LARND_I_FLAGS * n + s];
    sm.a[t] = A.rec[(int64_t)LARND_F_A * n + s];
    sm.b[t] = A.rec[(int64_t)LARND_F_B * n + s];
    sm.c[t] = A.rec[(int64_t)LARND_F_C * n + s];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      sm.wx[k][t] = A.rec[(int64_t)(LARND_F_WX0 + k) * n + s];
      sm.wy[k][t] = A.rec[(int64_t)(LARND_F_WY0 + k) * n + s];
    }
  }
  // bin groups of the 5-wide diffusion stencil for every in-pixel bin index b (pure function of b and nb)
  if (threadIdx.x < nb && threadIdx.x < 16) build_bin_groups(threadIdx.x, nb, A.half2, sm.g_n[threadIdx.x], sm.g_ox[threadIdx.x], sm.g_ci[threadIdx.x], sm.g_mask[threadIdx.x]);
  __syncthreads();
  // ---- runs: same key, start ticks within SPAN_MAX, at most 32 segments (one lane per segment) ---------------
  if (threadIdx.x == 0) {
    int nr = 0, cur = -1, tmin = 0, tmax = 0, nk = 0;
    for (int t = 0; t < ns; ++t) {
      if (!(sm.flags[t] & 1)) continue;  // outside every TPC: q == 0 and dq/dtheta == 0 (mask factor) -> no gradient
      const int T0 = __float_as_int(sm.seg[t].z);
      if (slow_only && seg_is_fast(T0, L, A.nticks)) continue;  // done by the class-sorted kernel
      bool fresh = cur < 0;
      if (!fresh) {
        const int t0s = sm.run[cur].start;
        fresh = sm.ep[t] != sm.ep[t0s] || sm.bx[t] != sm.bx[t0s] || sm.by[t] != sm.by[t0s] || sm.idx[t] != sm.idx[t0s] ||
                max(tmax, T0) - min(tmin, T0) > SPAN_MAX || t - t0s >= 32 ||
                t != sm.run[cur].start + sm.run[cur].len;  // masked segment in between: keep runs contiguous
      }
      if (fresh) {
        if (cur >= 0) { sm.run[cur].tmin = tmin; sm.run[cur].span = tmax - tmin; }
        cur = nr++;
        BRun& R = sm.run[cur];
        R.start = t;
        R.len = 1;
        R.ep = sm.ep[t];
        R.mpx = floordiv_i(sm.bx[t], nb);
        R.mpy = floordiv_i(sm.by[t], nb);
        int key = -1;
        for (int k = 0; k < nk; ++k)
          if (sm.kep[k] == R.ep && sm.kpx[k] == R.mpx && sm.kpy[k] == R.mpy) key = k;
        if (key < 0 && nk < MAXK) { key = nk++; sm.kep[key] = R.ep; sm.kpx[key] = R.mpx; sm.kpy[key] = R.mpy; }
        R.key = key;
        tmin = tmax = T0;
      } else {
        sm.run[cur].len += 1;
        tmin = min(tmin, T0);
        tmax = max(tmax, T0);
      }
    }
    if (cur >= 0) { sm.run[cur].tmin = tmin; sm.run[cur].span = tmax - tmin; }
    sm.nruns = nr;
    sm.nkeys = nk;
    sm.next_run = 0;
  }
#pragma unroll
  for (int k = 0; k < LARND_NPARAMS; ++k) sm.gl[k][threadIdx.x] = 0.0f;
  __syncthreads();
  // ---- row table: one warp per distinct main pixel ------------------------------------------------------------------
  for (int k = wid; k < sm.nkeys; k += BWD_WARPS) {
    int cnt = 0;
    for (int e0 = 0; e0 < TW * TW; e0 += 32) {
      const int e = e0 + lane;
      int raw = -1;
      bool work = false;
      if (e < TW * TW) {
        const int dx = e / TW - tm, dy = e % TW - tm;
        raw = raw_row(lk, sm.kpx[k] + dx, sm.kpy[k] + dy, sm.kep[k], A.nxp, A.nyp);
        sm.krow[k][e] = raw;
        // neighbour units with work: inside the stencil, not the centre (-999), and — when garbage gradients are known to
        // be zero — only pixels that own a non-garbage row
        const bool in_stencil = abs(dx) <= A.n_neigh && abs(dy) <= A.n_neigh;
        const bool centre = dx == 0 && dy == 0;
        if (skip_garbage) work = in_stencil && !centre && raw >= 0 && !(raw & (1 << 30));
        else work = in_stencil;
      }
      const unsigned m = __ballot_sync(0xffffffffu, work);
      if (work) sm.klist[k][cnt + __popc(m & ((1u << lane) - 1))] = (unsigned char)e;
      cnt += __popc(m);
    }
    if (lane == 0) sm.kcount[k] = cnt;
  }
  __syncthreads();
  const int nruns = sm.nruns;
#define GACC(k) sm.gl[k][threadIdx.x]

  for (;;) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&sm.next_run, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= nruns) break;
    const BRun R = sm.run[r];
    const int t0s = R.start;
    const int tl = t0s + min(lane, R.len - 1);          // this lane's segment (lanes >= len mirror the last one, unused)
    const bool live = lane < R.len;
    const float4 sg = sm.seg[tl];
    const float q = sg.x, f = sg.y, omf = 1.0f - f;
    const int T0 = __float_as_int(sg.z);
    const int m = T0 - R.tmin;
    const int ep = R.ep, idx = sm.idx[t0s];
    const int mpx = R.mpx, mpy = R.mpy;
    const int bxm = sm.bx[t0s] - mpx * nb, bym = sm.by[t0s] - mpy * nb;
    const bool fast = (R.tmin >= 2) && (R.tmin + R.span + L <= A.nticks - 1);
    const int npos = R.span + 2;
#pragma unroll
    for (int k = 0; k < 5; ++k) { sm.dwx[wid][k][lane] = 0.f; sm.dwy[wid][k][lane] = 0.f; }
    float dq = 0.f, df = 0.f, da = 0.f, db = 0.f, dc = 0.f;  // per-lane (= per-segment) accumulators
    const float ca_ = sm.a[tl], cb_ = sm.b[tl], cc_ = sm.c[tl];
    // boundary-correction tables are read at ct(m) = nt - L - (tmin + lane) by lane m
    int ctl = nt - L - (R.tmin + lane);
    ctl = max(0, min(ctl, nt - 1));
    const int ctl1 = min(ctl + 1, nt - 1);
    auto table_row = [&](int dx, int dy) -> int {
      return R.key >= 0 ? sm.krow[R.key][(dx + tm) * TW + (dy + tm)] : raw_row(lk, mpx + dx, mpy + dy, ep, A.nxp, A.nyp);
    };

    // ---------------- main pixels: merged diffusion-bin groups, 3-template blend ------------------------------------
    const int ngx = sm.g_n[bxm], ngy = sm.g_n[bym];
    for (int gx = 0; gx < ngx; ++gx) {
      const int ox = (int)sm.g_ox[bxm][gx] - 1, cix = sm.g_ci[bxm][gx], mx = sm.g_mask[bxm][gx];
      float gwx = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) gwx += (mx >> i & 1) ? sm.wx[i][tl] : 0.f;
      for (int gy = 0; gy < ngy; ++gy) {
        const int oy = (int)sm.g_ox[bym][gy] - 1, ciy = sm.g_ci[bym][gy], my = sm.g_mask[bym][gy];
        const int raw = table_row(ox, oy);
        if (raw < 0) continue;                                   // not a main pixel: dropped (sim_jax.py:152-154)
        if (skip_garbage && (raw & (1 << 30))) continue;
        const int row = raw & ~(1 << 30);
        float gwy = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) gwy += (my >> j & 1) ? sm.wy[j][tl] : 0.f;
        const float w = gwx * gwy;
        GRow grow;
        grow.open(A, row, lane);
        const int bin = cix * 5 + ciy;
        const float* ra = A.rm + (int64_t)((idx - 1) * 25 + bin) * A.Lp;
        const float* rb = A.rm + (int64_t)(idx * 25 + bin) * A.Lp;
        const float* rc = A.rm + (int64_t)((idx + 1) * 25 + bin) * A.Lp;
        const float* crow = A.cm + (int64_t)(idx * 25 + bin) * nt;
        float Pv = 0.f;
        if (fast) {
          float gr[NG];
          load_gwin<NG>(gr, grow, R.tmin, A.nticks, lane);
          const float Gla = correlate<NG>(gr, ra, npos, L, lane);
          const float Glb = correlate<NG>(gr, rb, npos, L, lane);
          const float Glc = correlate<NG>(gr, rc, npos, L, lane);
          const float gC = grow.at(R.tmin - 1 + min(lane, npos), true);
          const float Cav = __ldg(crow + ctl), Cbv = __ldg(crow + ctl1), Cl = __ldg(crow + nt - L);
          const float a0 = __shfl_sync(0xffffffffu, Gla, m), a1 = __shfl_sync(0xffffffffu, Gla, m + 1);
          const float b0 = __shfl_sync(0xffffffffu, Glb, m), b1 = __shfl_sync(0xffffffffu, Glb, m + 1);
          const float c0v = __shfl_sync(0xffffffffu, Glc, m), c1v = __shfl_sync(0xffffffffu, Glc, m + 1);
          const float gB = __shfl_sync(0xffffffffu, gC, m), gA = __shfl_sync(0xffffffffu, gC, m + 1);
          const float Ca = __shfl_sync(0xffffffffu, Cav, m), Cb = __shfl_sync(0xffffffffu, Cbv, m);
          const float D = Cl - (Ca * omf + Cb * f), dD = -(Cb - Ca);
          const float gm = fmaf(f, gB, omf * gA);
          const float Sa = fmaf(f, a0, omf * a1), Sb = fmaf(f, b0, omf * b1), Sc = fmaf(f, c0v, omf * c1v);
          Pv = fmaf(ca_, Sa, fmaf(cb_, Sb, cc_ * Sc)) + gm * D;
          const float Fd = fmaf(ca_, a0 - a1, fmaf(cb_, b0 - b1, cc_ * (c0v - c1v))) + (gB - gA) * D + gm * dD;
          const float qb = w * q;
          dq = fmaf(w, Pv, dq);
          df = fmaf(qb, Fd, df);
          da = fmaf(qb, Sa, da);
          db = fmaf(qb, Sb, db);
          dc = fmaf(qb, Sc, dc);
        } else {
          for (int t = 0; t < R.len; ++t) {
            const float4 s2 = sm.seg[t0s + t];
            const int T2 = __float_as_int(s2.z);
            const float f2 = s2.y, o2 = 1.0f - f2;
            const float* const rows[3] = {ra, rb, rc};
            float G0[3], G1[3], gB, gA;
            slow_sums<3>(grow, rows, T2, L, A.nticks, lane, G0, G1, gB, gA);
            const int ct = nt - L - T2;
            const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, nt - 1)), Cl = __ldg(crow + nt - L);
            const float D = Cl - (Ca * o2 + Cb * f2), dD = -(Cb - Ca);
            const float gm = fmaf(f2, gB, o2 * gA);
            if (lane == t) {
              const float Sa = fmaf(f2, G0[0], o2 * G1[0]), Sb = fmaf(f2, G0[1], o2 * G1[1]), Sc = fmaf(f2, G0[2], o2 * G1[2]);
              Pv = fmaf(ca_, Sa, fmaf(cb_, Sb, cc_ * Sc)) + gm * D;
              const float Fd = fmaf(ca_, G0[0] - G1[0], fmaf(cb_, G0[1] - G1[1], cc_ * (G0[2] - G1[2]))) + (gB - gA) * D + gm * dD;
              const float qb = w * q;
              dq = fmaf(w, Pv, dq);
              df = fmaf(qb, Fd, df);
              da = fmaf(qb, Sa, da);
              db = fmaf(qb, Sb, db);
              dc = fmaf(qb, Sc, dc);
            }
          }
        }
        // d/dWx_i = sum_j Wy_j q P_ij with P_ij = Pv for every member (i,j) of this group
        const float px_ = gwy * q * Pv, py_ = gwx * q * Pv;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          if (mx >> i & 1) sm.dwx[wid][i][lane] += px_;
          if (my >> i & 1) sm.dwy[wid][i][lane] += py_;
        }
      }
    }
    // ---------------- neighbour pixels: template 0, full charge ---------------------------------------------------------
    const int n_list = R.key >= 0 ? sm.kcount[R.key] : TW * TW;
    for (int li = 0; li < n_list; ++li) {
      const int e = R.key >= 0 ? (int)sm.klist[R.key][li] : li;
      const int dx = e / TW - tm, dy = e % TW - tm;
      int row;
      {
        if (abs(dx) > A.n_neigh || abs(dy) > A.n_neigh) continue;
        const bool centre = dx == 0 && dy == 0;
        const int raw = centre ? -1 : table_row(dx, dy);
        const bool garbage = raw < 0 || (raw & (1 << 30));
        if (skip_garbage && garbage) continue;
        row = raw < 0 ? 0 : (raw & ~(1 << 30));               // absent / centre -> waveform row 0 (sim_jax.py:724-725)
      }
      GRow grow;
      grow.open(A, row, lane);
      const int ci = abs(2 * bxm - A.half2 - 2 * nb * dx) >> 1, cj = abs(2 * bym - A.half2 - 2 * nb * dy) >> 1;
      const int bin = ci * A.ny_lut + cj;
      const float* rowp = A.r0 + (int64_t)bin * A.Lp;
      const float* crow = A.c0 + (int64_t)bin * nt;
      if (fast) {
        float gr[NG];
        load_gwin<NG>(gr, grow, R.tmin, A.nticks, lane);
        const float Gl = correlate<NG>(gr, rowp, npos, L, lane);
        const float gC = grow.at(R.tmin - 1 + min(lane, npos), true);       // tick tmin - 1 + lane
        const float Cav = __ldg(crow + ctl), Cbv = __ldg(crow + ctl1), Cl = __ldg(crow + nt - L);
        const float G0 = __shfl_sync(0xffffffffu, Gl, m), G1 = __shfl_sync(0xffffffffu, Gl, m + 1);
        const float gB = __shfl_sync(0xffffffffu, gC, m), gA = __shfl_sync(0xffffffffu, gC, m + 1);
        const float Ca = __shfl_sync(0xffffffffu, Cav, m), Cb = __shfl_sync(0xffffffffu, Cbv, m);
        const float D = Cl - (Ca * omf + Cb * f), dD = -(Cb - Ca);
        const float gm = fmaf(f, gB, omf * gA);
        dq += fmaf(f, G0, omf * G1) + gm * D;
        df += q * ((G0 - G1) + (gB - gA) * D + gm * dD);
      } else {
        for (int t = 0; t < R.len; ++t) {
          const float4 s2 = sm.seg[t0s + t];
          const int T2 = __float_as_int(s2.z);
          const float f2 = s2.y, o2 = 1.0f - f2;
          const float* const rows[1] = {rowp};
          float G0[1], G1[1], gB, gA;
          slow_sums<1>(grow, rows, T2, L, A.nticks, lane, G0, G1, gB, gA);
          const int ct = nt - L - T2;
          const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, nt - 1)), Cl = __ldg(crow + nt - L);
          const float D = Cl - (Ca * o2 + Cb * f2), dD = -(Cb - Ca);
          const float gm = fmaf(f2, gB, o2 * gA);
          if (lane == t) {
            dq += fmaf(f2, G0[0], o2 * G1[0]) + gm * D;
            df += s2.x * ((G0[0] - G1[0]) + (gB - gA) * D + gm * dD);
          }
        }
      }
    }
    // ---- K1b: chain rule through the per-segment preparation, one lane per segment ---------------------------
    if (live) {
      float dwx[5], dwy[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) { dwx[k] = sm.dwx[wid][k][lane]; dwy[k] = sm.dwy[wid][k][lane]; }
      chain_rule_segment(p, A.rec, n, s_base + tl, idx, q, dq, df, da, db, dc, dwx, dwy, [&](int k, float v) { GACC(k) += v; });
    }
    __syncwarp();
  }
  // ---- warp + block reduction -> per-chunk partials ----------------------------------------------------------
#pragma unroll
  for (int k = 0; k < LARND_NPARAMS; ++k) {
    const float v = warp_sum(GACC(k));
    if (lane == 0) sm.grad[wid][k] = v;
  }
#undef GACC
  __syncthreads();
  if (threadIdx.x < LARND_NPARAMS) {
    float v = 0.f;
    for (int w = 0; w < BWD_WARPS; ++w) v += sm.grad[w][threadIdx.x];
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

template <int NG>
int launch_bwd(const BwdArgs& A, const larnd_params_t& p, int64_t chunks, cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = sizeof(BwdSmem);
  if (!attr_done) {
    LARND_CUDA(cudaFuncSetAttribute(k_lut_backward<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  k_lut_backward<NG><<<(unsigned)chunks, BWD_THREADS, smem, st>>>(A, p);
  prof_end(2, st);
  LARND_LAUNCH_CHECK("k_lut_backward");
  return LARND_OK;
}

}  // namespace

int larnd_launch_accumulate_bwd(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                int32_t npix_capacity, int32_t flags, const float* g_wfs, int64_t g_stride,
                                float* grad_params, const int32_t* counts, cudaStream_t st, const StepsView* steps) {
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
  A.chunk = larnd_chunk_size(n);
  const int64_t chunks = (n + A.chunk - 1) / A.chunk;
  int* gflag = reinterpret_cast<int*>(ws.partials + (size_t)ws.n_chunks_max * 16) - 4;  // last 16 bytes of the partials area
  LARND_CUDA(cudaMemsetAsync(gflag, 0, sizeof(int), st));
  A.use_steps = steps ? 1 : 0;
  if (steps) {
    A.steps = *steps;          // rows of pixel ids < 0 carry no event by construction: the garbage flag stays 0
    A.skip_garbage = 1;
  } else {
    A.steps = StepsView{nullptr};
    k_garbage_grad_flag<<<32, 256, 0, st>>>(g_wfs, g_stride, p.n_ticks, counts, gflag);
    LARND_LAUNCH_CHECK("k_garbage_grad_flag");
  }
  A.garbage_grad_nonzero = gflag;
  // large batches: the class-sorted kernel (accumulate_bwd_sorted.cu) does the bulk.  LARND_FLAG_IMPL_CHUNK / _SORTED
  // override the size rule (the tests force both paths on small batches).
  bool sorted = larnd_sorted_supported(p, lut) && n >= LARND_SORTED_MIN_SEGMENTS;
  if (flags & LARND_FLAG_IMPL_SORTED) sorted = larnd_sorted_supported(p, lut) != 0;
  if (flags & LARND_FLAG_IMPL_CHUNK) sorted = false;
  // (the tile kernel addresses the gradient rows with signed 32-bit element offsets)
  if (!steps && (int64_t)npix_capacity * g_stride >= ((int64_t)1 << 31)) sorted = false;
  if (steps && lut->nt - lut->L < 1) sorted = false;   // the running-sum form needs one sample in front of the response window
  A.sorted_active = sorted ? 1 : 0;
  A.n_slow = sorted ? ws.gcnt + 3 /* GC_SLOW, sorted_runs.cuh */ : nullptr;
  prof_begin(2, st);
  if (sorted) {
    float* sorted_partials = ws.partials + (size_t)ws.n_chunks_max * 16;
    int n_slots = 0;
    int rc0 = larnd_launch_accumulate_bwd_sorted(n, p, lut, ws, npix_capacity, flags, g_wfs, g_stride, sorted_partials, &n_slots,
                                                 gflag, counts, st, steps);
    if (rc0) return rc0;
    k_reduce_partials<<<LARND_NPARAMS, 256, 0, st>>>(sorted_partials, n_slots, grad_params);
    LARND_LAUNCH_CHECK("k_reduce_partials");
  }
  // The chunk kernel below takes what the tile kernel leaves: segments whose window ends beyond the readout — none when the
  // response is no longer than the readout (T0 + L <= nt <= n_ticks - 2) — and everything when garbage rows carry gradient
  // (decided on the device unless the caller promised otherwise, or the gradient comes as step events).  When the host knows
  // there is nothing left, the pass (78 k CTAs writing zero partials + their reduction at spill size) is not launched.
  if (sorted && lut->nt <= p.n_ticks - 2 && (steps || (flags & 1))) {
    prof_end(2, st);
    return LARND_OK;
  }
  const int need = lut->L + SPAN_MAX + 2;  // gradient window of a run: ticks tmin .. tmin + span + 1 + L - 1
  int rc;
  if (need <= 32 * 2) rc = launch_bwd<2>(A, p, chunks, st);
  else if (need <= 32 * 4) rc = launch_bwd<4>(A, p, chunks, st);
  else if (need <= 32 * 6) rc = launch_bwd<6>(A, p, chunks, st);
  else if (need <= 32 * 10) rc = launch_bwd<10>(A, p, chunks, st);
  else if (need <= 32 * 16) rc = launch_bwd<16>(A, p, chunks, st);
  else { larnd_set_error("signal_length %d too large for the backward register window", lut->L); return LARND_E_ARG; }
  if (rc) return rc;
  k_reduce_partials<<<LARND_NPARAMS, 256, 0, st>>>(ws.partials, chunks, grad_params);
  LARND_LAUNCH_CHECK("k_reduce_partials");
  return LARND_OK;
}
