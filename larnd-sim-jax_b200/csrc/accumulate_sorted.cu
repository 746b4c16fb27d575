// K3' class-sorted lut_accumulate (forward), sm_100a — the large-batch path of simulate_signals
// (reference sim_jax.py:142-286; same arithmetic as accumulate.cu, different work decomposition).
//
// accumulate.cu walks a chunk of consecutive segments and re-reads the response rows from L1/L2 for every
// (run, unit) pair: ~75 % of its instructions are gather/bookkeeping around the multiply-adds.  Here the runs of
// the whole batch are first SORTED BY RESPONSE CLASS (longitudinal template index, sub-pixel bin inside the pixel):
// every run of a class reads the same response rows for the same unit, so a warp loads the rows of its unit ONCE
// into registers — for every tick of its 32*NS-tick window and every impulse position j < KPT the sample
// R[x - 1 - j] — and then streams the runs of the class through pure FFMAs:
//
//   k_build_runs     chunk of 128 segments -> runs (same definition as accumulate.cu, only "fast" segments whose
//                    whole window lies inside the readout), class histogram                          [global atomics]
//   k_class_scan     exclusive scan of the histogram -> class offsets, tile table (<= 32 runs of one class)
//   k_scatter_runs   counting-sort scatter of the run records
//   k_acc_tiles      persistent CTAs pull tiles; per tile 8 warps pull units (merged diffusion-bin groups,
//                    the neighbourhood-sum row, neighbour pixels that own a waveform row).  Per (unit, tile):
//                    lane <-> run builds the impulse trains + boundary corrections in shared memory, then
//                    lane <-> tick applies them from the register-resident response and flushes every
//                    (run, unit) window with coalesced red.global.add.f32.
//
// Segments whose window touches the ends of the readout (garbage tick 0 handling, sim_jax.py:177-178,243-244)
// are left to accumulate.cu's per-segment path (larnd_launch_accumulate with mode = slow-only).
#include "sorted_runs.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ tile consumer
constexpr int SEGMAX = TR * MAXLEN;  // segments per tile

struct TileSmem {
  int4 run[TR];                 // start, len | span << 16, tmin, class
  int ep[TR], mpx[TR], mpy[TR], soff[TR];
  // per-segment data of the tile, staged once (thread <-> segment) and read by every unit's build phase
  float qf[SEGMAX], qo[SEGMAX], ca[SEGMAX], cb[SEGMAX], cc[SEGMAX], fr[SEGMAX];
  int m[SEGMAX];                // T0 - tmin of the run
  float wxg[5][SEGMAX], wyg[5][SEGMAX];  // transverse weights merged per group of the class
  unsigned char owner[SEGMAX];
  float hN[TR][KPT];            // neighbour impulse train (full segment charge)
  float mom[TR][MS];            // neighbour correction moments A1,A2,A3,B1,B3 per position
  float ph[NW][TR * HS];        // per-warp trains of the current unit: [run][3*j + template]
  float pE[NW][TR * ES];        // per-warp merged boundary corrections: [run][position]
  unsigned char g_n[16], g_ox[16][5], g_ci[16][5], g_mask[16][5];
  signed char udx[225], udy[225];  // relative pixel of every neighbour unit (P <= 15)
  int tile, next_unit, nseg;
  int low_end;  // some run of the tile starts below tick 2 (its windows need the garbage-column handling)
};

template <int NS, int NR, int KP>
__device__ __forceinline__ void load_response(float (&Rw)[3][NS][KP], const float* const (&rows)[NR], int Lp, int lane) {
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

// Adds one (run, unit) window to a waveform row: acc = window part, Ev = merged boundary correction of this lane
// (slot 0, lanes < ES).  Column of (slot s, lane) is tmin - 1 + 32 s + lane.  Window samples are valid on columns
// >= 2, corrections on columns >= 1; what falls below goes to the garbage column 0 (sim_jax.py:177-178,243-244).
// dst = address of this lane's tick of slot 0 (row base + tmin - 1 + lane); fast = no run of the tile starts below tick 2
// (the window end is inside the row for every run of the sorted path: seg_is_fast).
template <int NS, bool LP>
__device__ __forceinline__ void emit_window(const float (&acc)[NS], float Ev, float* dst, int tmin, int nticks, int lane,
                                            float sign, const bool (&act)[NS], bool fast) {
  if (fast) {  // tile-uniform, the common case: every active tick of the window is a regular column of the row
    // act[s] = this lane's tick of slot s lies inside the run window (L + 2 + span ticks, the same for every run of the
    // tile): hoisted predicates instead of a zero test per slot; zero-valued adds inside the window are harmless
    if (LP) {  // kernel-level constant: L + NPOS lies in [32 (NS - 1), 32 NS) for every NPOS
      // the usual shape (the window ends inside the last slot): every lane of the other slots is active, so they need no
      // predicate -- ptxas wraps every predicated reduction in BSSY / BRA / BSYNC
      atomicAdd(dst, sign * (acc[0] + Ev));  // RED.E.ADD.F32, coalesced
#pragma unroll
      for (int s = 1; s < NS - 1; ++s) atomicAdd(dst + 32 * s, sign * acc[s]);
      if (act[NS - 1]) atomicAdd(dst + 32 * (NS - 1), sign * acc[NS - 1]);
    } else {
      if (act[0]) atomicAdd(dst, sign * (acc[0] + Ev));
#pragma unroll
      for (int s = 1; s < NS; ++s)
        if (act[s]) atomicAdd(dst + 32 * s, sign * acc[s]);
    }
  } else {
    float* rowbase = dst - (tmin - 1) - lane;
    float g = 0.0f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int col = tmin - 1 + 32 * s + lane;
      const float w = acc[s], e = (s == 0) ? Ev : 0.0f;
      const float v = (col >= 2 ? w : 0.0f) + (col >= 1 ? e : 0.0f);
      g += (col < 2 ? w : 0.0f) + (col < 1 ? e : 0.0f);
      if (v != 0.0f && col < nticks) atomicAdd(rowbase + col, sign * v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
    if (lane == 0 && g != 0.0f) atomicAdd(rowbase, sign * g);
  }
}

// Consume loop of one unit: every run of `todo` gets its window computed from the register-resident response and flushed.
// Two runs are in flight per iteration (independent FFMA chains hide the 4-cycle dependency latency) and the number of
// impulse positions is a compile-time constant (uniform per tile: the class key contains the tick span).
template <int NS, int NR, int NPOS, bool LP, int KP, bool TWO>
__device__ __forceinline__ void consume_pairs(const SortArgs& A, const TileSmem& sm, const float (&Rw)[3][NS][KP], unsigned todo, int row,
                                              const float* __restrict__ hbuf, int hstride, const float* __restrict__ Ebuf,
                                              float* __restrict__ row0, int mode /* 0 own row, 1 sum row -> row0, 2 own row and -row0 */,
                                              int lane) {
  const bool fast = !sm.low_end;
  // lane <-> run: 32-bit (signed: tmin - 1 may be negative on row 0) element offset of column tmin - 1 of the run's row
  // inside the target buffer (the launcher keeps npix * nticks below 2^31 for this kernel); one shuffle + one wide add per
  // window instead of 64-bit row arithmetic
  float* const base = mode == 1 ? row0 : A.wfs;
  const int myoff = (mode == 1 ? 0 : row * A.nticks) + (sm.run[lane].z - 1);
  bool act[NS];  // window length of every run of the tile: L + 2 + span = L + NPOS ticks
#pragma unroll
  for (int s = 0; s < NS; ++s) act[s] = 32 * s + lane < A.L + NPOS;
  while (todo) {
    const int p0 = __ffs(todo) - 1;
    todo &= todo - 1;
    const bool two = TWO && NR == 3 && todo != 0u;  // neighbour units (one template, dual flush) run one window at a time: measured faster
    const int p1 = two ? __ffs(todo) - 1 : p0;
    if (two) todo &= todo - 1;
    float* const d0 = base + (__shfl_sync(0xffffffffu, myoff, p0) + lane);
    float* const d1 = base + (__shfl_sync(0xffffffffu, myoff, p1) + lane);
    const int t0 = sm.run[p0].z, t1 = sm.run[p1].z;
    const float* h0 = hbuf + p0 * hstride;
    const float* h1 = hbuf + p1 * hstride;
    float a0[NS], a1[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { a0[s] = 0.0f; a1[s] = 0.0f; }
#pragma unroll
    for (int j = 0; j < NPOS; ++j) {
      float u[NR], v[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) { u[r] = h0[NR * j + r]; v[r] = h1[NR * j + r]; }
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          a0[s] = fmaf(u[r], Rw[r][s][j], a0[s]);
          a1[s] = fmaf(v[r], Rw[r][s][j], a1[s]);
        }
    }
    const float E0 = lane < ES ? Ebuf[p0 * ES + lane] : 0.0f;
    const float E1 = lane < ES ? Ebuf[p1 * ES + lane] : 0.0f;
    emit_window<NS, LP>(a0, E0, d0, t0, A.nticks, lane, 1.0f, act, fast);
    if (mode == 2) emit_window<NS, LP>(a0, E0, row0 + (t0 - 1) + lane, t0, A.nticks, lane, -1.0f, act, fast);
    if (two) {
      emit_window<NS, LP>(a1, E1, d1, t1, A.nticks, lane, 1.0f, act, fast);
      if (mode == 2) emit_window<NS, LP>(a1, E1, row0 + (t1 - 1) + lane, t1, A.nticks, lane, -1.0f, act, fast);
    }
  }
}

template <int NS, int NR, bool LP, int KP, bool TWO>
__device__ __forceinline__ void consume_pairs_npos(const SortArgs& A, const TileSmem& sm, const float (&Rw)[3][NS][KP], unsigned todo, int row,
                                                   const float* hbuf, int hstride, const float* Ebuf, float* row0, int mode, int lane, int npos) {
  // warp-uniform: a short compare chain per unit; only the position counts this variant can meet (2 .. KP) are instantiated
  constexpr int N3 = KP >= 3 ? 3 : KP, N4 = KP >= 4 ? 4 : KP, N5 = KP >= 5 ? 5 : KP;
  if (KP >= 6 && npos >= 6) consume_pairs<NS, NR, KP, LP, KP, TWO>(A, sm, Rw, todo, row, hbuf, hstride, Ebuf, row0, mode, lane);
  else if (KP >= 5 && npos == 5) consume_pairs<NS, NR, N5, LP, KP, TWO>(A, sm, Rw, todo, row, hbuf, hstride, Ebuf, row0, mode, lane);
  else if (KP >= 4 && npos == 4) consume_pairs<NS, NR, N4, LP, KP, TWO>(A, sm, Rw, todo, row, hbuf, hstride, Ebuf, row0, mode, lane);
  else if (KP >= 3 && npos == 3) consume_pairs<NS, NR, N3, LP, KP, TWO>(A, sm, Rw, todo, row, hbuf, hstride, Ebuf, row0, mode, lane);
  else consume_pairs<NS, NR, 2, LP, KP, TWO>(A, sm, Rw, todo, row, hbuf, hstride, Ebuf, row0, mode, lane);
}

// KP = impulse positions whose response samples are held in registers.  The KP = KPT kernel can serve every tile; with
// KP = 4 (3) the response needs 48 (36) instead of 72 registers and three (four) CTAs fit on an SM — 24 (32) instead of 16
// warps to hide the latencies the kernel is bound by.  A launch serves the tiles of spans span_lo .. span_hi (KP >= span_hi
// + 2), a contiguous range of the tile table, and pulls them through its own counter gcnt[GC_FWD + launch].
template <int NS, bool LP, int KP, bool TWO>
__global__ void __launch_bounds__(TILE_THREADS, NS <= 4 ? (KP <= 3 ? 4 : (KP <= 4 ? 3 : 2)) : (NS == 5 && KP <= 4 ? 2 : 1))
k_acc_tiles(const __grid_constant__ SortArgs A, const int span_lo, const int span_hi, const int launch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem& sm = *reinterpret_cast<TileSmem*>(smem_raw);
  if (A.counts[2] != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = A.nb, L = A.L;
  const int* irec = reinterpret_cast<const int*>(A.rec);
  const int64_t n = A.n;
  RowLookup lk = A.lk;
  lk.n_unique = A.counts[0];
  lk.n_neg = A.counts[1];
  // merged transverse-diffusion groups per in-pixel bin (same tables as accumulate.cu)
  if (threadIdx.x < nb && threadIdx.x < 16) build_bin_groups(threadIdx.x, nb, A.half2, sm.g_n[threadIdx.x], sm.g_ox[threadIdx.x], sm.g_ci[threadIdx.x], sm.g_mask[threadIdx.x]);
  for (int u = threadIdx.x; u < A.P * A.P; u += TILE_THREADS) {
    sm.udx[u] = (signed char)(u / A.P - A.n_neigh);
    sm.udy[u] = (signed char)(u % A.P - A.n_neigh);
  }
  const int tile_lo = A.gcnt[GC_SPAN + span_lo];
  const int ntiles = A.gcnt[GC_SPAN + span_hi + 1];
  int* const tile_counter = A.gcnt + GC_FWD + launch;
  const int n_neigh_units = A.P * A.P;
  const int n_units = 25 + 1 + n_neigh_units;  // merged diffusion groups, neighbourhood-sum row, neighbour pixels
  float* myh = sm.ph[warp];
  float* myE = sm.pE[warp];
  // Waveform row 0 collects the neighbourhood sum of EVERY run (sim_jax.py:724-725): ~13 windows per run from all CTAs
  // onto the same 8 KB would serialise in the L2 atomic units, so each CTA reduces into a private copy (summed by
  // k_reduce_row0 afterwards).
  float* row0 = A.row0 + (int64_t)blockIdx.x * A.nticks;

  for (;;) {
    __syncthreads();  // everybody is done with the previous tile
    if (threadIdx.x == 0) { sm.tile = tile_lo + atomicAdd(tile_counter, 1); sm.next_unit = 0; }
    __syncthreads();
    const int tile = sm.tile;
    if (tile >= ntiles) break;
    const int4 ti = A.tile_info[tile];
    const int cls = ti.x, count = ti.z;
    const int ncb = A.ncls / (SPAN_MAX_S + 1);
    const int cls_b = cls % ncb;     // class = span * (ntpl * nb * nb) + (idx * nb + bxm) * nb + bym
    const int npos = cls / ncb + 2;  // impulse positions of every run of this tile
    const int bym = cls_b % nb, bxm = (cls_b / nb) % nb, idx = cls_b / (nb * nb);
    // ---- stage the runs (warp 0: lane <-> run) ---------------------------------------------------------------
    if (warp == 0) {
      int len = 0;
      if (lane < count) {
        const int4 e = A.runs[ti.y + lane];
        sm.run[lane] = e;
        const int64_t s0 = e.x;
        sm.ep[lane] = irec[(int64_t)LARND_I_EP * n + s0];
        sm.mpx[lane] = floordiv_i(irec[(int64_t)LARND_I_BX * n + s0], nb);
        sm.mpy[lane] = floordiv_i(irec[(int64_t)LARND_I_BY * n + s0], nb);
        len = e.y & 0xffff;
      }
      const unsigned low = __ballot_sync(0xffffffffu, lane < count && sm.run[lane].z < 2);
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
    // ---- stage the segments (thread <-> segment): products, Lagrange weights, group-merged transverse weights ---
    for (int i = threadIdx.x; i < sm.nseg; i += TILE_THREADS) {
      const int r = sm.owner[i];
      const int4 e = sm.run[r];
      const int64_t s = (int64_t)e.x + (i - sm.soff[r]);
      const float q = A.rec[(int64_t)LARND_F_Q * n + s], f = A.rec[(int64_t)LARND_F_FRAC * n + s];
      sm.qf[i] = q * f;
      sm.qo[i] = q * (1.0f - f);
      sm.fr[i] = f;
      sm.ca[i] = A.rec[(int64_t)LARND_F_A * n + s];
      sm.cb[i] = A.rec[(int64_t)LARND_F_B * n + s];
      sm.cc[i] = A.rec[(int64_t)LARND_F_C * n + s];
      sm.m[i] = irec[(int64_t)LARND_I_T0 * n + s] - e.z;
      float vx[LARND_NB_TRAN_BINS], vy[LARND_NB_TRAN_BINS];
#pragma unroll
      for (int k = 0; k < LARND_NB_TRAN_BINS; ++k) {
        vx[k] = A.rec[(int64_t)(LARND_F_WX0 + k) * n + s];
        vy[k] = A.rec[(int64_t)(LARND_F_WY0 + k) * n + s];
      }
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        const int mx = g < sm.g_n[bxm] ? sm.g_mask[bxm][g] : 0, my = g < sm.g_n[bym] ? sm.g_mask[bym][g] : 0;
        float sx = 0.0f, sy = 0.0f;
#pragma unroll
        for (int k = 0; k < LARND_NB_TRAN_BINS; ++k) { sx += (mx >> k & 1) ? vx[k] : 0.0f; sy += (my >> k & 1) ? vy[k] : 0.0f; }
        sm.wxg[g][i] = sx;
        sm.wyg[g][i] = sy;
      }
    }
    __syncthreads();
    // ---- neighbour impulse train + correction moments (thread <-> run) -------------------------------------------
    if (threadIdx.x < count) {
      const int r = threadIdx.x;
#pragma unroll
      for (int k = 0; k < KPT; ++k) sm.hN[r][k] = 0.0f;
      for (int k = 0; k < 5 * KPT; ++k) sm.mom[r][k] = 0.0f;
      const int len = sm.run[r].y & 0xffff, so = sm.soff[r];
      for (int t = 0; t < len; ++t) {
        const int i = so + t;
        const float qf = sm.qf[i], qo = sm.qo[i], f = sm.fr[i], o = 1.0f - f;
        const int m = sm.m[i];
        sm.hN[r][m] += qf;
        sm.hN[r][m + 1] += qo;
        float* mo = sm.mom[r] + 5 * m;
        mo[0] += qo; mo[1] += qo * o; mo[2] += qf * o; mo[3] += qf; mo[4] += qf * f;
      }
    }
    __syncthreads();

    for (;;) {
      int unit = 0;
      if (lane == 0) unit = atomicAdd(&sm.next_unit, 1);
      unit = __shfl_sync(0xffffffffu, unit, 0);
      if (unit >= n_units) break;
      float Rw[3][NS][KP];
      if (unit < 25) {
        // ---------------- merged diffusion-bin group (gi, gj): 3-template blend on a main pixel ----------------
        const int gi = unit / LARND_NB_TRAN_BINS, gj = unit % LARND_NB_TRAN_BINS;
        if (gi >= sm.g_n[bxm] || gj >= sm.g_n[bym]) continue;
        const int bin = (int)sm.g_ci[bxm][gi] * 5 + (int)sm.g_ci[bym][gj];
        const int ox = (int)sm.g_ox[bxm][gi] - 1, oy = (int)sm.g_ox[bym][gj] - 1;
        int row = -1;
        if (lane < count) {
          const int pid = pixel2id_dev(sm.mpx[lane] + ox, sm.mpy[lane] + oy, sm.ep[lane], A.nxp, A.nyp);
          row = lookup_row(lk, pid);  // not in the list -> dropped (sim_jax.py:152-154)
          if (A.skip_garbage && pid < 0) row = -1;
        }
        if (__ballot_sync(0xffffffffu, row >= 0) == 0u) continue;
        // build: lane <-> run
        if (row >= 0) {
          float* h = myh + lane * HS;
          float* E = myE + lane * ES;
#pragma unroll
          for (int k = 0; k < 3 * KPT; ++k) h[k] = 0.0f;
#pragma unroll
          for (int k = 0; k < ES; ++k) E[k] = 0.0f;
          const int4 e = sm.run[lane];
          const int len = e.y & 0xffff, tmin = e.z, so = sm.soff[lane];
          const float* crow = A.cm + (int64_t)(idx * 25 + bin) * A.nt;
          const float Cl = __ldg(crow + A.nt - L);
          const float* wxp = sm.wxg[gi];
          const float* wyp = sm.wyg[gj];
          for (int t = 0; t < len; ++t) {
            const int i = so + t;
            const int m = sm.m[i];
            // boundary correction of this segment (sim_jax.py:236-247): D lands on tick T0 (weight 1-f) and T0-1 (f)
            int ct = A.nt - L - (tmin + m);
            ct = max(0, min(ct, A.nt - 1));
            const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, A.nt - 1));
            const float w = wxp[i] * wyp[i];
            const float qf = w * sm.qf[i], qo = w * sm.qo[i], f = sm.fr[i];
            const float ca = sm.ca[i], cb = sm.cb[i], cc = sm.cc[i];
            float* hm = h + 3 * m;
            hm[0] = fmaf(qf, ca, hm[0]); hm[1] = fmaf(qf, cb, hm[1]); hm[2] = fmaf(qf, cc, hm[2]);
            hm[3] = fmaf(qo, ca, hm[3]); hm[4] = fmaf(qo, cb, hm[4]); hm[5] = fmaf(qo, cc, hm[5]);
            const float D = Cl - (Ca * (1.0f - f) + Cb * f);
            E[m] = fmaf(qf, D, E[m]);
            E[m + 1] = fmaf(qo, D, E[m + 1]);
          }
        }
        __syncwarp();
        const float* const rows[3] = {A.rm + (int64_t)((idx - 1) * 25 + bin) * A.Lp, A.rm + (int64_t)(idx * 25 + bin) * A.Lp,
                                      A.rm + (int64_t)((idx + 1) * 25 + bin) * A.Lp};
        load_response<NS, 3, KP>(Rw, rows, A.Lp, lane);
        // consume: lane <-> tick
        consume_pairs_npos<NS, 3, LP, KP, TWO>(A, sm, Rw, __ballot_sync(0xffffffffu, row >= 0), row, myh, HS, myE, row0, 0, lane, npos);
        __syncwarp();
      } else {
        // ---------------- neighbour pixels: template 0, full segment charge (sim_jax.py:197-225,250-261) ----------
        // unit 25 = the neighbourhood-sum row into waveform row 0; the others add to the neighbour's own row (if it
        // is a main pixel of the batch) and take the same deposit back out of row 0 (see accumulate.cu)
        const bool sum_unit = unit == 25;
        if (sum_unit && A.skip_garbage) continue;
        const int u = unit - 26;
        const int dx = sum_unit ? 0 : (int)sm.udx[u], dy = sum_unit ? 0 : (int)sm.udy[u];
        if (!sum_unit && dx == 0 && dy == 0) continue;  // the centre id is -999: row 0 only (inside the sum row)
        int row = -1;
        if (lane < count) {
          if (sum_unit) row = 0;
          else {
            const int pid = pixel2id_dev(sm.mpx[lane] + dx, sm.mpy[lane] + dy, sm.ep[lane], A.nxp, A.nyp);
            row = lookup_row(lk, pid);
            if (row <= 0) row = -1;
            else if (A.skip_garbage && pid < 0) row = -1;
          }
        }
        unsigned owned = __ballot_sync(0xffffffffu, row >= 0);
        if (owned == 0u) continue;
        const float* rowp0;
        const float* crow;
        if (sum_unit) {
          const int sb = bxm * nb + bym;
          rowp0 = A.sr + (int64_t)sb * A.Lp;
          crow = A.sc + (int64_t)sb * A.nt;
        } else {
          const int vx = 2 * bxm - A.half2 - 2 * nb * dx, vy = 2 * bym - A.half2 - 2 * nb * dy;
          const int bin = (abs(vx) >> 1) * A.ny_lut + (abs(vy) >> 1);
          rowp0 = A.r0 + (int64_t)bin * A.Lp;
          crow = A.c0 + (int64_t)bin * A.nt;
        }
        if (row >= 0) {  // merged corrections from the run's moments: E_j = e0_j + e1_{j-1}
          const int4 e = sm.run[lane];
          const int span = e.y >> 16, tmin = e.z;
          const float Cl = __ldg(crow + A.nt - L);
          float e1prev = 0.0f;
          float* E = myE + lane * ES;
#pragma unroll
          for (int j = 0; j < ES; ++j) {
            float Ej = 0.0f;
            if (j <= span + 1) {
              float e1 = 0.0f, e0 = 0.0f;
              if (j <= span) {
                int ct = A.nt - L - (tmin + j);
                ct = max(0, min(ct, A.nt - 1));
                const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, A.nt - 1));
                const float* mo = sm.mom[lane] + 5 * j;
                e1 = Cl * mo[0] - Ca * mo[1] - Cb * mo[2];
                e0 = Cl * mo[3] - Ca * mo[2] - Cb * mo[4];
              }
              Ej = e0 + e1prev;
              e1prev = e1;
            }
            E[j] = Ej;
          }
        }
        __syncwarp();
        const float* const rows[1] = {rowp0};
        load_response<NS, 1, KP>(Rw, rows, A.Lp, lane);
        const bool dual = !sum_unit && !A.skip_garbage;
        consume_pairs_npos<NS, 1, LP, KP, TWO>(A, sm, Rw, owned, row, &sm.hN[0][0], KPT, myE, row0, sum_unit ? 1 : (dual ? 2 : 0), lane, npos);
        __syncwarp();
      }
    }
  }
}

__global__ void k_reduce_row0(const float* __restrict__ row0, int ncopies, int nticks, float* __restrict__ wfs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nticks) return;
  float acc = 0.0f;
  for (int k = 0; k < ncopies; ++k) acc += row0[(int64_t)k * nticks + c];
  if (acc != 0.0f) atomicAdd(wfs + c, acc);
}

}  // namespace

size_t larnd_sorted_workspace_bytes(int64_t n) {
  size_t b = 0;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  b += align_up(nn * sizeof(int4), 256) * 2;                                       // runs_tmp, runs
  b += align_up((size_t)LARND_NCLS_MAX * sizeof(int), 256) * 3;                    // class_count, class_start, cursor
  b += align_up((nn / TR + LARND_NCLS_MAX + 1) * sizeof(int4), 256);               // tile_info
  b += 256;                                                                        // counters
  b += align_up((size_t)LARND_ROW0_COPIES * LARND_ROW0_TICKS_MAX * sizeof(float), 256);  // private garbage rows
  return b;
}

void larnd_carve_sorted(char* p, int64_t n, Workspace* ws) {
  const size_t nn = (size_t)(n > 0 ? n : 1);
  ws->runs_tmp = p; p += align_up(nn * sizeof(int4), 256);
  ws->runs = p; p += align_up(nn * sizeof(int4), 256);
  ws->class_count = reinterpret_cast<int*>(p); p += align_up((size_t)LARND_NCLS_MAX * sizeof(int), 256);
  ws->class_start = reinterpret_cast<int*>(p); p += align_up((size_t)LARND_NCLS_MAX * sizeof(int), 256);
  ws->cursor = reinterpret_cast<int*>(p); p += align_up((size_t)LARND_NCLS_MAX * sizeof(int), 256);
  ws->tile_info = p; p += align_up((nn / TR + LARND_NCLS_MAX + 1) * sizeof(int4), 256);
  ws->gcnt = reinterpret_cast<int*>(p); p += 256;
  ws->row0 = reinterpret_cast<float*>(p);
}

int larnd_sorted_supported(const larnd_params_t& p, const larnd_lut* lut) {
  const int nb = p.nb_sampling_bins_per_pixel;
  if (lut->ntpl * nb * nb * (SPAN_MAX_S + 1) > LARND_NCLS_MAX) return 0;
  if (lut->L + 2 + SPAN_MAX_S > 32 * 6) return 0;
  if (p.n_ticks > LARND_ROW0_TICKS_MAX) return 0;
  return 1;
}


int larnd_launch_accumulate_sorted(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                   int32_t npix_capacity, int32_t flags, float* wfs, const int32_t* counts, cudaStream_t st) {
  if (n == 0) return LARND_OK;
  SortArgs A;
  {
    int rc0 = larnd_lut_ensure_neighbour_sums(const_cast<larnd_lut*>(lut), p.nb_sampling_bins_per_pixel, p.number_pix_neighbors, st);
    if (rc0) return rc0;
  }
  A.wfs = wfs;
  A.skip_garbage = flags & 1;
  prof_begin(1, st);
  {
    int rc0 = sorted_fill_and_build(A, n, p, lut, ws, npix_capacity, counts, st);
    if (rc0) return rc0;
  }
  const int need = lut->L + 2 + SPAN_MAX_S;
  const int ns = need <= 32 * 4 ? 4 : (need <= 32 * 5 ? 5 : 6);
  const int split = (ns == 4) ? sorted_split_mode() : 0;
  const bool split5 = ns == 5 && sorted_split_mode() >= 1;  // 150-tick windows: 60 response registers, 2 CTAs/SM instead of 1
  const int grid = sorted_grid(2, LARND_ROW0_COPIES);
  const int grid3 = split ? sorted_grid(3, LARND_ROW0_COPIES) : 0;
  const int grid4 = split >= 2 ? sorted_grid(4, LARND_ROW0_COPIES) : 0;
  const int ncopies = max(grid, max(grid3, grid4));
  if (!A.skip_garbage) LARND_CUDA(cudaMemsetAsync(ws.row0, 0, (size_t)ncopies * p.n_ticks * sizeof(float), st));
  const size_t smem = sizeof(TileSmem);
  static bool attr_done = false;
  if (!attr_done) {
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<4, true, KPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<5, true, KPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<6, true, KPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<4, false, KPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<5, false, KPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<6, false, KPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<4, true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<4, false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<4, true, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<5, true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<5, false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<4, false, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  // LP: for every number of impulse positions (2 .. KPT) the run window L + NPOS ends inside the last 32-tick slot, so
  // all other slots are flushed without a predicate
  const bool lp = lut->L + 2 >= 32 * (ns - 1) && lut->L + KPT < 32 * ns;
  int big_lo = 0;  // first span left to the KPT kernel
  if (split >= 2) {
    if (lp) k_acc_tiles<4, true, 3, true><<<grid4, TILE_THREADS, smem, st>>>(A, 0, 1, 0); else k_acc_tiles<4, false, 3, true><<<grid4, TILE_THREADS, smem, st>>>(A, 0, 1, 0);
    LARND_LAUNCH_CHECK("k_acc_tiles<3>");
    big_lo = 2;
  }
  if (split >= 1) {
    if (lp) k_acc_tiles<4, true, 4, true><<<grid3, TILE_THREADS, smem, st>>>(A, big_lo, 2, 1); else k_acc_tiles<4, false, 4, true><<<grid3, TILE_THREADS, smem, st>>>(A, big_lo, 2, 1);
    LARND_LAUNCH_CHECK("k_acc_tiles<4>");
    big_lo = 3;
  }
  if (ns == 4) { if (lp) k_acc_tiles<4, true, KPT, true><<<grid, TILE_THREADS, smem, st>>>(A, big_lo, SPAN_MAX_S, 2); else k_acc_tiles<4, false, KPT, true><<<grid, TILE_THREADS, smem, st>>>(A, big_lo, SPAN_MAX_S, 2); }
  else if (ns == 5) {
    if (split5) {
      if (lp) k_acc_tiles<5, true, 4, true><<<grid, TILE_THREADS, smem, st>>>(A, 0, 2, 1); else k_acc_tiles<5, false, 4, true><<<grid, TILE_THREADS, smem, st>>>(A, 0, 2, 1);
      LARND_LAUNCH_CHECK("k_acc_tiles<5,4>");
      big_lo = 3;
    }
    if (lp) k_acc_tiles<5, true, KPT, true><<<grid, TILE_THREADS, smem, st>>>(A, big_lo, SPAN_MAX_S, 2); else k_acc_tiles<5, false, KPT, true><<<grid, TILE_THREADS, smem, st>>>(A, big_lo, SPAN_MAX_S, 2);
  }
  else { if (lp) k_acc_tiles<6, true, KPT, true><<<grid, TILE_THREADS, smem, st>>>(A, 0, SPAN_MAX_S, 2); else k_acc_tiles<6, false, KPT, true><<<grid, TILE_THREADS, smem, st>>>(A, 0, SPAN_MAX_S, 2); }
  LARND_LAUNCH_CHECK("k_acc_tiles");
  if (!A.skip_garbage) {
    k_reduce_row0<<<(p.n_ticks + 255) / 256, 256, 0, st>>>(ws.row0, ncopies, p.n_ticks, wfs);
    LARND_LAUNCH_CHECK("k_reduce_row0");
  }
  // segments whose window touches the ends of the readout: per-segment path of accumulate.cu
  int rc = larnd_launch_accumulate(n, p, lut, ws, npix_capacity, flags | LARND_ACC_SLOW_ONLY, wfs, counts, st);
  prof_end(1, st);
  return rc;
}
