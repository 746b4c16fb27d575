// K3s class-sorted lut_accumulate (forward), sm_100a — the large-batch path of simulate_signals
// (reference sim_jax.py:142-286; same arithmetic as accumulate.cu, different work decomposition).
//
// The runs of the whole batch are SORTED BY RESPONSE CLASS (longitudinal template index, sub-pixel bin inside the pixel,
// tick span; sorted_runs.cuh) and cut into tiles of <= 32 runs.  A persistent CTA pulls a tile, stages its segments in
// shared memory and then works in two phases that overlap freely between its 8 warps:
//
//   phase A  main pixels (3-template blend on the <= 3x3 pixels the diffusion stencil reaches): warp <-> merged
//            diffusion-bin group.  Every run of the tile reads the same three response rows for a group, so the warp
//            holds them in REGISTERS and streams the runs through pure FFMAs.
//   phase B  neighbour pixels (template 0, full segment charge): warp <-> run.  A run owns a waveform row for only
//            ~15 % of its (2n+1)^2 neighbours; everything else lands in the garbage row 0 (sim_jax.py:724-725), which by
//            linearity is (neighbourhood-sum response) - (owned neighbours).  The warp keeps that row-0 frame in registers
//            while it walks the owned neighbours of its run and flushes it ONCE (the round-1 kernel flushed every owned
//            window twice: +own row, -row 0).
//
// Output layout: lane <-> 4 consecutive ticks.  A run's frame starts at B = (tmin - 1) & ~3, so with a waveform row
// stride that is a multiple of 4 every lane's float4 is 16-byte aligned and a 128-tick frame leaves as ONE
// red.global.add.v4.f32 per lane (round 1: lane <-> tick, four scalar reductions).  The alignment shift s = (tmin-1) & 3
// is absorbed by four pre-shifted copies of the response tables (lut_tables.cu::k_shift_rows): a lane reads the
// NPOS + 3 samples its four ticks need with two or three aligned 128-bit loads, whatever s is.  Runs are ordered by s
// inside a tile, so phase A reloads its registers at most four times per (tile, group).
//
// Segments whose window ends beyond the readout are left to accumulate.cu's per-segment path
// (larnd_launch_accumulate with LARND_ACC_SLOW_ONLY); windows sticking out at the LOW end (garbage column 0,
// sim_jax.py:177-178,243-244) are handled here by the generic flush.
#include "sorted_runs.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ tile consumer
constexpr int SEGMAX = TR * MAXLEN;  // segments per tile
#ifndef LARND_KP4_CTAS
#define LARND_KP4_CTAS 3
#endif
#ifndef LARND_KP6_CTAS
#define LARND_KP6_CTAS 3
#endif
// Boundary corrections of a (run, unit) frame live at frame positions j + s <= KP + 2; a lane reads the four positions of
// its ticks, lanes whose ticks lie beyond read the always-zero rows KP + 3 .. KP + 6 (no branch in the consume loops).
// Row stride TRE = 33: a lane-dependent ROW with a uniform column must not land in one bank.
constexpr int TRE = TR + 1;
// (per warp only the rows a lane below USED / 4 can touch, rounded up to whole groups of four; every other lane reads the
// CTA-wide zero block TileSmem::zE — shared memory given back to the L1, which the kernel is sensitive to)
template <int KP> struct ECfg { static constexpr int USED = KP + 3, EK = (KP + 3 + 3) / 4 * 4; };
// per-warp train buffer: built as [run][3 j + template] (stride HS, conflict-free read-modify-write), then re-laid in place
// as float4 [j][run] = (3 templates, 0) so that the consume loop fetches a position with ONE broadcast 128-bit load
template <int KP> struct HCfg {
  static constexpr int HS = 3 * KP + 1;   // per-lane stride of the build layout (odd: conflict-free)
  static constexpr int FLOATS = (TR * HS > KP * TR * 4) ? TR * HS : KP * TR * 4;
};

template <int KP>
struct TileSmem {
  int4 run[TR];                 // start, len | span << 16, tmin, class
  int ep[TR], mpx[TR], mpy[TR], soff[TR];
  int prow[9][TR];              // waveform row of the 3x3 pixels around every run's main pixel (-1: not in the list)
  // per-segment data of the tile, staged once (thread <-> segment) and read by every group's build phase
  float qf[SEGMAX], qo[SEGMAX], ca[SEGMAX], cb[SEGMAX], cc[SEGMAX], fr[SEGMAX];
  int m[SEGMAX];                // T0 - tmin of the run
  float wxg[5][SEGMAX], wyg[5][SEGMAX];  // transverse weights merged per group of the class
  unsigned char owner[SEGMAX];
  float hN[TR][KPT];            // neighbour impulse train (full segment charge)
  float mom[TR][MS];            // neighbour correction moments A1,A2,A3,B1,B3 per position
  __align__(16) float ph[NW][HCfg<KP>::FLOATS];  // per-warp trains of the current group (see HCfg)
  float pE[NW][ECfg<KP>::EK * TRE];               // per-warp boundary corrections: [frame position][run or unit]
  float zE[4 * TRE];                              // always zero: the corrections of lanes whose ticks lie beyond position USED
  unsigned char g_n[16], g_ox[16][5], g_ci[16][5], g_mask[16][5];
  signed char udx[225], udy[225];  // relative pixel of every neighbour unit (P <= 15)
  int ubin[96];                 // neighbour unit -> response row (template 0 bin; entry P*P: the neighbourhood-sum row)
  float ucl[96];                // ... and its running sum at Nt - L
  unsigned smask[4];            // runs of the tile per alignment shift
  int tile, next_unit, next_run, nseg;
  int low_end;  // some run of the tile starts below tick 2 (its frames need the garbage-column handling)
};

// KP = impulse positions whose response samples are held in registers: a lane's four ticks need samples
// 4x + c + 7 - j (c < 4, j < KP) of the shifted row, i.e. floats [4x + WOFF, 4x + WOFF + WN)
template <int KP> struct WCfg { static constexpr int WN = KP <= 4 ? 8 : 12, WOFF = KP <= 4 ? 4 : 0; };

// [region: load_w]
template <int NSV, int NR, int KP>
__device__ __forceinline__ void load_w(float (&W)[3][NSV][WCfg<KP>::WN], const float* const (&rows)[NR], int lane) {
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int v = 0; v < NSV; ++v)
#pragma unroll
      for (int q = 0; q < WCfg<KP>::WN / 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rows[r] + 128 * v + 4 * lane + WCfg<KP>::WOFF + 4 * q));
        W[r][v][4 * q + 0] = t.x; W[r][v][4 * q + 1] = t.y; W[r][v][4 * q + 2] = t.z; W[r][v][4 * q + 3] = t.w;
      }
}

// [region: emit]
__device__ __forceinline__ void red_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Frame of one (run, target row): lane holds columns B + 128 v + 4 lane + c; acc = window part (boundary corrections of
// slot 0 already added by the caller on the fast paths).  d = address of this lane's first column.  v4: aligned rows, one
// vector reduction per active slot; otherwise scalar reductions.
template <int NSV>
__device__ __forceinline__ void emit_fast(const float (&acc)[NSV][4], float* d, const bool (&act)[NSV], bool v4) {
  if (v4) {
#pragma unroll
    for (int v = 0; v < NSV; ++v)
      if (act[v]) red_v4(d + 128 * v, acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
  } else {
#pragma unroll
    for (int v = 0; v < NSV; ++v)
      if (act[v]) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (acc[v][c] != 0.0f) atomicAdd(d + 128 * v + c, acc[v][c]);
      }
  }
}

// Frames of runs that start below tick 2: window samples are valid on columns >= 2, corrections (e: this lane's four
// ticks of slot 0) on columns >= 1, what falls below goes to the garbage column 0 (sim_jax.py:177-178,243-244).
template <int NSV>
__device__ __forceinline__ void emit_generic(const float (&acc)[NSV][4], const float (&e)[4], float* rowbase, int B, int nticks,
                                             int lane, float sign) {
  float g = 0.0f;
#pragma unroll
  for (int v = 0; v < NSV; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = B + 128 * v + 4 * lane + c;
      const float w = acc[v][c], ee = (v == 0) ? e[c] : 0.0f;
      const float val = (col >= 2 ? w : 0.0f) + (col >= 1 ? ee : 0.0f);
      g += (col < 2 ? w : 0.0f) + (col < 1 ? ee : 0.0f);
      if (val != 0.0f && col < nticks) atomicAdd(rowbase + col, sign * val);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
  if (lane == 0 && g != 0.0f) atomicAdd(rowbase, sign * g);
}

// [region: conv_frame]
// window part of a frame: acc[v][c] += sum_j sum_r h[NR j + r] W[r][v][c + 7 - j - WOFF]
template <int NSV, int NR, int NPOS, int KP>
__device__ __forceinline__ void conv_frame(float (&acc)[NSV][4], const float (&W)[3][NSV][WCfg<KP>::WN], const float* __restrict__ h) {
#pragma unroll
  for (int j = 0; j < NPOS; ++j) {
    float u[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) u[r] = h[NR * j + r];
#pragma unroll
    for (int v = 0; v < NSV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[v][c] = fmaf(u[r], W[r][v][c + 7 - j - WCfg<KP>::WOFF], acc[v][c]);
  }
}

// [region: consume_main]
// Phase A consume loop of one (group, shift): every run of `todo` gets its frame computed from the register-resident
// response and flushed.  Two runs are in flight per iteration (independent FFMA chains) and the number of impulse
// positions is a compile-time constant (uniform per tile: the class key contains the tick span).
// myoff (lane <-> run) = element offset of the run's frame start (row * stride + B) inside the waveform buffer.
template <int NSV, int NPOS, int KP>
__device__ __forceinline__ void consume_main(const SortArgs& A, const TileSmem<KP>& sm, const float (&W)[3][NSV][WCfg<KP>::WN], unsigned todo,
                                             int myoff, const float* __restrict__ hbuf, const float* __restrict__ Ebuf, int s, int lane,
                                             int path) {
  bool act[NSV];
#pragma unroll
  for (int v = 0; v < NSV; ++v) act[v] = 128 * v + 4 * lane - s < A.L + NPOS;
  const float* const Elane = 4 * lane < ECfg<KP>::USED ? Ebuf + 4 * lane * TRE : sm.zE;
  const float4* const h4 = reinterpret_cast<const float4*>(hbuf);
  float* const wl = A.wfs + 4 * lane;
  while (todo) {
    const int p0 = __ffs(todo) - 1;
    todo &= todo - 1;
    const bool two = todo != 0u;
    const int p1 = two ? __ffs(todo) - 1 : p0;
    if (two) todo &= todo - 1;
    const int off0 = __shfl_sync(0xffffffffu, myoff, p0), off1 = __shfl_sync(0xffffffffu, myoff, p1);
    float a0[NSV][4], a1[NSV][4], e0[4], e1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { e0[c] = Elane[c * TRE + p0]; e1[c] = Elane[c * TRE + p1]; }
#pragma unroll
    for (int v = 0; v < NSV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) { a0[v][c] = v == 0 ? e0[c] : 0.0f; a1[v][c] = v == 0 ? e1[c] : 0.0f; }
#pragma unroll
    for (int j = 0; j < NPOS; ++j) {
      const float4 u4 = h4[j * TR + p0], w4 = h4[j * TR + p1];
      const float u[3] = {u4.x, u4.y, u4.z}, w[3] = {w4.x, w4.y, w4.z};
#pragma unroll
      for (int v = 0; v < NSV; ++v)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            a0[v][c] = fmaf(u[r], W[r][v][c + 7 - j - WCfg<KP>::WOFF], a0[v][c]);
            a1[v][c] = fmaf(w[r], W[r][v][c + 7 - j - WCfg<KP>::WOFF], a1[v][c]);
          }
    }
    if (path < 2) {
      emit_fast<NSV>(a0, wl + off0, act, path == 0);
      if (two) emit_fast<NSV>(a1, wl + off1, act, path == 0);
    } else {  // myoff = row * stride here (the frame start may lie below the row); window and corrections go separately
#pragma unroll
      for (int c = 0; c < 4; ++c) { a0[0][c] -= e0[c]; a1[0][c] -= e1[c]; }
      emit_generic<NSV>(a0, e0, A.wfs + off0, (sm.run[p0].z - 1) & ~3, A.nticks, lane, 1.0f);
      if (two) emit_generic<NSV>(a1, e1, A.wfs + off1, (sm.run[p1].z - 1) & ~3, A.nticks, lane, 1.0f);
    }
  }
}

// [region: dispatch]
template <int NSV, int KP>
__device__ __forceinline__ void consume_main_npos(const SortArgs& A, const TileSmem<KP>& sm, const float (&W)[3][NSV][WCfg<KP>::WN], unsigned todo,
                                                  int myoff, const float* hbuf, const float* Ebuf, int s, int lane, int path, int npos) {
  // warp-uniform: only the position counts this variant can meet (2 .. KP) are instantiated
  constexpr int N3 = KP >= 3 ? 3 : KP, N4 = KP >= 4 ? 4 : KP, N5 = KP >= 5 ? 5 : KP;
  if (KP >= 6 && npos >= 6) consume_main<NSV, KP, KP>(A, sm, W, todo, myoff, hbuf, Ebuf, s, lane, path);
  else if (KP >= 5 && npos == 5) consume_main<NSV, N5, KP>(A, sm, W, todo, myoff, hbuf, Ebuf, s, lane, path);
  else if (KP >= 4 && npos == 4) consume_main<NSV, N4, KP>(A, sm, W, todo, myoff, hbuf, Ebuf, s, lane, path);
  else if (KP >= 3 && npos == 3) consume_main<NSV, N3, KP>(A, sm, W, todo, myoff, hbuf, Ebuf, s, lane, path);
  else consume_main<NSV, 2, KP>(A, sm, W, todo, myoff, hbuf, Ebuf, s, lane, path);
}

// [region: neigh_frame]
// Phase B helpers.
template <int NSV, int NPOS, int KP>
__device__ __forceinline__ void neigh_frame(float (&acc)[NSV][4], const float (&e4)[4], const float* __restrict__ trow4 /* row + 4 lane */,
                                            const float (&hreg)[KPT]) {
  float W[NSV][WCfg<KP>::WN];
#pragma unroll
  for (int v = 0; v < NSV; ++v)
#pragma unroll
    for (int q = 0; q < WCfg<KP>::WN / 4; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(trow4 + 128 * v + WCfg<KP>::WOFF + 4 * q));
      W[v][4 * q + 0] = t.x; W[v][4 * q + 1] = t.y; W[v][4 * q + 2] = t.z; W[v][4 * q + 3] = t.w;
    }
#pragma unroll
  for (int v = 0; v < NSV; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[v][c] = v == 0 ? e4[c] : 0.0f;
#pragma unroll
  for (int j = 0; j < NPOS; ++j)
#pragma unroll
    for (int v = 0; v < NSV; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[v][c] = fmaf(hreg[j], W[v][c + 7 - j - WCfg<KP>::WOFF], acc[v][c]);
}

// [region: neigh_run setup]
// One run of phase B (NPOS = span + 2 impulse positions, compile-time).  Units = the (2n+1)^2 neighbour pixels plus, as
// unit P*P, the neighbourhood-sum row (row n0 + in-pixel bin of the shifted template-0 table); lane <-> unit (three words
// of 32) for the row lookups and for the boundary corrections (every lane walks the <= NPOS positions of ITS unit: full
// SIMD width), then the warp walks the owned units of the word with lane <-> 4 ticks.
template <int NSV, int NPOS, int KP>
__device__ __forceinline__ void neigh_run(const SortArgs& A, const TileSmem<KP>& sm, const RowLookup& lk, int p, float* __restrict__ myE,
                                          float* row0, int lane, int basepath) {
  const int4 e = sm.run[p];
  const int tmin = e.z;
  const int s = (tmin - 1) & 3, B = (tmin - 1) & ~3;
  const int path = tmin < 2 ? 2 : basepath;
  const int nu = A.P * A.P;
  const int mpx = sm.mpx[p], mpy = sm.mpy[p], ep = sm.ep[p];
  float hreg[KPT];
#pragma unroll
  for (int j = 0; j < KPT; ++j) hreg[j] = j < NPOS ? sm.hN[p][j] : 0.0f;
  bool act[NSV];
#pragma unroll
  for (int v = 0; v < NSV; ++v) act[v] = 128 * v + 4 * lane - s < A.L + NPOS;
  const float* const t4 = A.t0s + ((int64_t)s * A.n0rows * A.lps + 4 * lane);  // shifted copy s, this lane's ticks
  float* const wl = A.wfs + 4 * lane;
  const float* const Elane = 4 * lane < ECfg<KP>::USED ? myE + 4 * lane * TRE : sm.zE;
  const float* const mom = sm.mom[p];
  const int ctbase = A.nt - A.L - tmin;
  float acc0[NSV][4];  // frame of the garbage row: neighbourhood sum minus the neighbours that own a row
#pragma unroll
  for (int v = 0; v < NSV; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc0[v][c] = 0.0f;
  // [region: neigh lookups]
  // ---- which units own a waveform row: the lookups of all (<= 3) words are issued together, so their two dependent L2
  // loads overlap instead of heading every word of the loop below ----
  int rows3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int u = lane + 32 * k;
    int row = -1;
    if (u < nu) {
      const int dx = sm.udx[u], dy = sm.udy[u];
      if (dx != 0 || dy != 0) {  // the centre id is -999: row 0 only (inside the sum row)
        const int pid = pixel2id_dev(mpx + dx, mpy + dy, ep, A.nxp, A.nyp);
        row = lookup_row(lk, pid);
        if (row <= 0) row = -1;
        else if (A.skip_garbage && pid < 0) row = -1;
      }
    } else if (u == nu && !A.skip_garbage) row = 0;  // the neighbourhood-sum unit
    rows3[k] = row;
  }
#pragma unroll 1
  for (int k = 0; 32 * k <= nu; ++k) {
    const int u = lane + 32 * k;
    const int row = k == 0 ? rows3[0] : (k == 1 ? rows3[1] : rows3[2]);
    unsigned m = __ballot_sync(0xffffffffu, row >= 0);
    if (m == 0u) continue;
    // [region: neigh corrections]
    // ---- boundary corrections of the owned units (lane <-> unit): E_j = e0_j + e1_{j-1} at frame position j + s ----
    int woff = 0, toff = 0;  // element offsets of the unit's frame start in wfs and of its response row in the table
    if (row >= 0) {
      const int bin = sm.ubin[u];
      woff = (int)((int64_t)row * A.wstride) + (path < 2 ? B : 0);
      toff = bin * A.lps;
#pragma unroll
      for (int q = 0; q < ECfg<KP>::USED; ++q) myE[q * TRE + lane] = 0.0f;
      const float* crow = u == nu ? A.sc + (int64_t)(bin - A.n0) * A.nt : A.c0 + (int64_t)bin * A.nt;
      const float Cl = sm.ucl[u];
      float e1prev = 0.0f;
      float* E = myE + s * TRE + lane;
#pragma unroll
      for (int j = 0; j < NPOS - 1; ++j) {  // positions 0 .. span
        int ct = ctbase - j;
        ct = max(0, min(ct, A.nt - 1));
        const float Ca = __ldg(crow + ct), Cb = __ldg(crow + min(ct + 1, A.nt - 1));
        const float* mo = mom + 5 * j;
        const float e1 = Cl * mo[0] - Ca * mo[1] - Cb * mo[2];
        const float e0 = Cl * mo[3] - Ca * mo[2] - Cb * mo[4];
        E[j * TRE] = e0 + e1prev;
        e1prev = e1;
      }
      E[(NPOS - 1) * TRE] = e1prev;
    }
    __syncwarp();
    // [region: neigh owned loop]
    // ---- frames of the owned units ----
    const int sumbit = (nu >> 5) == k ? (nu & 31) : -1;  // lane of the neighbourhood-sum unit in this word
    if (path == 2) {  // low end: window and corrections follow different garbage rules, no register merging
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int wo = __shfl_sync(0xffffffffu, woff, b), to = __shfl_sync(0xffffffffu, toff, b);
        float acc[NSV][4], e4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) e4[c] = Elane[c * TRE + b];
        neigh_frame<NSV, NPOS, KP>(acc, e4, t4 + to, hreg);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[0][c] -= e4[c];
        if (b == sumbit) emit_generic<NSV>(acc, e4, row0, B, A.nticks, lane, 1.0f);
        else {
          emit_generic<NSV>(acc, e4, A.wfs + wo, B, A.nticks, lane, 1.0f);
          if (!A.skip_garbage) emit_generic<NSV>(acc, e4, row0, B, A.nticks, lane, -1.0f);
        }
      }
    } else {
      if (sumbit >= 0 && (m >> sumbit & 1u)) {  // the neighbourhood-sum frame seeds the garbage-row accumulator
        m &= ~(1u << sumbit);
        const int to = __shfl_sync(0xffffffffu, toff, sumbit);
        float acc[NSV][4], e4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) e4[c] = Elane[c * TRE + sumbit];
        neigh_frame<NSV, NPOS, KP>(acc, e4, t4 + to, hreg);
#pragma unroll
        for (int v = 0; v < NSV; ++v)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc0[v][c] += acc[v][c];
      }
      // two owned units in flight: their response / correction loads are issued together (the kernel is bound by the
      // latency of these L2 loads, not by issue slots)
      while (m) {
        const int b0 = __ffs(m) - 1;
        m &= m - 1;
        const bool two = m != 0u;
        const int b1 = two ? __ffs(m) - 1 : b0;
        if (two) m &= m - 1;
        const int wo0 = __shfl_sync(0xffffffffu, woff, b0), to0 = __shfl_sync(0xffffffffu, toff, b0);
        const int wo1 = __shfl_sync(0xffffffffu, woff, b1), to1 = __shfl_sync(0xffffffffu, toff, b1);
        float W0[NSV][WCfg<KP>::WN], W1[NSV][WCfg<KP>::WN];
#pragma unroll
        for (int v = 0; v < NSV; ++v)
#pragma unroll
          for (int q = 0; q < WCfg<KP>::WN / 4; ++q) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(t4 + to0 + 128 * v + WCfg<KP>::WOFF + 4 * q));
            const float4 y = __ldg(reinterpret_cast<const float4*>(t4 + to1 + 128 * v + WCfg<KP>::WOFF + 4 * q));
            W0[v][4 * q + 0] = x.x; W0[v][4 * q + 1] = x.y; W0[v][4 * q + 2] = x.z; W0[v][4 * q + 3] = x.w;
            W1[v][4 * q + 0] = y.x; W1[v][4 * q + 1] = y.y; W1[v][4 * q + 2] = y.z; W1[v][4 * q + 3] = y.w;
          }
        float a0[NSV][4], a1[NSV][4];
#pragma unroll
        for (int v = 0; v < NSV; ++v)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            a0[v][c] = v == 0 ? Elane[c * TRE + b0] : 0.0f;
            a1[v][c] = v == 0 ? Elane[c * TRE + b1] : 0.0f;
          }
#pragma unroll
        for (int j = 0; j < NPOS; ++j)
#pragma unroll
          for (int v = 0; v < NSV; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              a0[v][c] = fmaf(hreg[j], W0[v][c + 7 - j - WCfg<KP>::WOFF], a0[v][c]);
              a1[v][c] = fmaf(hreg[j], W1[v][c + 7 - j - WCfg<KP>::WOFF], a1[v][c]);
            }
        emit_fast<NSV>(a0, wl + wo0, act, path == 0);
        if (two) emit_fast<NSV>(a1, wl + wo1, act, path == 0);
#pragma unroll
        for (int v = 0; v < NSV; ++v)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc0[v][c] -= two ? a0[v][c] + a1[v][c] : a0[v][c];
      }
    }
    __syncwarp();
  }
  if (!A.skip_garbage && path != 2) emit_fast<NSV>(acc0, row0 + B + 4 * lane, act, true);  // private copies are aligned
}

// [region: dispatch]
template <int NSV, int KP>
__device__ __forceinline__ void neigh_run_npos(const SortArgs& A, const TileSmem<KP>& sm, const RowLookup& lk, int p, float* myE, float* row0,
                                               int lane, int basepath, int npos) {
  constexpr int N3 = KP >= 3 ? 3 : KP, N4 = KP >= 4 ? 4 : KP, N5 = KP >= 5 ? 5 : KP;
  if (KP >= 6 && npos >= 6) neigh_run<NSV, KP, KP>(A, sm, lk, p, myE, row0, lane, basepath);
  else if (KP >= 5 && npos == 5) neigh_run<NSV, N5, KP>(A, sm, lk, p, myE, row0, lane, basepath);
  else if (KP >= 4 && npos == 4) neigh_run<NSV, N4, KP>(A, sm, lk, p, myE, row0, lane, basepath);
  else if (KP >= 3 && npos == 3) neigh_run<NSV, N3, KP>(A, sm, lk, p, myE, row0, lane, basepath);
  else neigh_run<NSV, 2, KP>(A, sm, lk, p, myE, row0, lane, basepath);
}

// [region: kernel prologue]
// A launch serves the tiles of spans span_lo .. span_hi (KP >= span_hi + 2), a contiguous range of the tile table, and
// pulls them through its own counter gcnt[GC_FWD + launch].
template <int NSV, int KP>
__global__ void __launch_bounds__(TILE_THREADS, NSV == 1 ? (KP <= 4 ? LARND_KP4_CTAS : LARND_KP6_CTAS) : 2)
k_acc_tiles(const __grid_constant__ SortArgs A, const int span_lo, const int span_hi, const int launch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<KP>& sm = *reinterpret_cast<TileSmem<KP>*>(smem_raw);
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
  for (int i = threadIdx.x; i < NW * ECfg<KP>::EK * TRE; i += TILE_THREADS) (&sm.pE[0][0])[i] = 0.0f;  // incl. the always-zero rows
  for (int i = threadIdx.x; i < 4 * TRE; i += TILE_THREADS) sm.zE[i] = 0.0f;
  const int tile_lo = A.gcnt[GC_SPAN + span_lo];
  const int ntiles = A.gcnt[GC_SPAN + span_hi + 1];
  int* const tile_counter = A.gcnt + GC_FWD + launch;
  float* myh = sm.ph[warp];
  float* myE = sm.pE[warp];
  // Waveform row 0 collects the garbage of EVERY run (sim_jax.py:724-725): one frame per run from all CTAs onto the same
  // 8 KB would serialise in the L2 atomic units, so each CTA reduces into a private copy (summed by k_reduce_row0).
  float* row0 = A.row0 + (int64_t)blockIdx.x * A.r0stride;
  const int basepath = A.v4ok ? 0 : 1;

  for (;;) {
    // [region: tile pull]
    __syncthreads();  // everybody is done with the previous tile
    if (threadIdx.x == 0) { sm.tile = tile_lo + atomicAdd(tile_counter, 1); sm.next_unit = 0; sm.next_run = 0; }
    __syncthreads();
    const int tile = sm.tile;
    if (tile >= ntiles) break;
    const int4 ti = A.tile_info[tile];
    const int cls = ti.x, count = ti.z;
    const int ncb = A.ncls / (SPAN_MAX_S + 1);
    const int cls_b = cls % ncb;     // class = span * (ntpl * nb * nb) + (idx * nb + bxm) * nb + bym
    const int npos = cls / ncb + 2;  // impulse positions of every run of this tile
    const int bym = cls_b % nb, bxm = (cls_b / nb) % nb, idx = cls_b / (nb * nb);
    // [region: stage runs]
    // ---- stage the runs (warp 0: lane <-> run) ---------------------------------------------------------------
    if (warp == 0) {
      int len = 0, tm = 2, sh = -1;
      if (lane < count) {
        const int4 e = A.runs[ti.y + lane];
        sm.run[lane] = e;
        const int64_t s0 = e.x;
        sm.ep[lane] = irec[(int64_t)LARND_I_EP * n + s0];
        sm.mpx[lane] = floordiv_i(irec[(int64_t)LARND_I_BX * n + s0], nb);
        sm.mpy[lane] = floordiv_i(irec[(int64_t)LARND_I_BY * n + s0], nb);
        len = e.y & 0xffff;
        tm = e.z;
        sh = (e.z - 1) & 3;
      }
      const unsigned low = __ballot_sync(0xffffffffu, tm < 2);
      if (lane == 0) sm.low_end = low != 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned mq = __ballot_sync(0xffffffffu, sh == q);
        if (lane == 0) sm.smask[q] = mq;
      }
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
    if (threadIdx.x >= 32 && threadIdx.x < 32 + 96) {  // per-unit response rows of this class (warps 1-3)
      const int u = threadIdx.x - 32, nu = A.P * A.P;
      int bin = 0;
      float cl = 0.0f;
      if (u < nu) {
        const int dx = u / A.P - A.n_neigh, dy = u % A.P - A.n_neigh;
        const int vx = 2 * bxm - A.half2 - 2 * nb * dx, vy = 2 * bym - A.half2 - 2 * nb * dy;
        bin = (abs(vx) >> 1) * A.ny_lut + (abs(vy) >> 1);
        cl = __ldg(A.c0 + (int64_t)bin * A.nt + A.nt - L);
      } else if (u == nu) {  // the neighbourhood-sum rows follow the nx * ny bins in the shifted table
        const int sb = bxm * nb + bym;
        bin = A.n0 + sb;
        cl = __ldg(A.sc + (int64_t)sb * A.nt + A.nt - L);
      }
      sm.ubin[u] = bin;
      sm.ucl[u] = cl;
    }
    __syncthreads();
    // rows of the 3x3 main pixels, looked up ONCE per (run, pixel) for all diffusion-bin groups of phase A: the two dependent
    // L2 loads of a lookup used to head every (group, tile) unit
    for (int i = threadIdx.x; i < 9 * TR; i += TILE_THREADS) {
      const int o = i / TR, r = i - o * TR;
      int row = -1;
      if (r < count) {
        const int pid = pixel2id_dev(sm.mpx[r] + o / 3 - 1, sm.mpy[r] + o % 3 - 1, sm.ep[r], A.nxp, A.nyp);
        row = lookup_row(lk, pid);  // not in the list -> dropped (sim_jax.py:152-154)
        if (A.skip_garbage && pid < 0) row = -1;
      }
      sm.prow[o][r] = row;
    }
    // [region: stage segments]
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
    // [region: moments]
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
    const int pathA = sm.low_end ? 2 : basepath;

    // [region: phaseA unit setup]
    // ================= phase A: merged diffusion-bin groups (gi, gj), 3-template blend on a main pixel =================
    for (;;) {
      int unit = 0;
      if (lane == 0) unit = atomicAdd(&sm.next_unit, 1);
      unit = __shfl_sync(0xffffffffu, unit, 0);
      if (unit >= 25) break;
      const int gi = unit / LARND_NB_TRAN_BINS, gj = unit % LARND_NB_TRAN_BINS;
      if (gi >= sm.g_n[bxm] || gj >= sm.g_n[bym]) continue;
      const int bin = (int)sm.g_ci[bxm][gi] * 5 + (int)sm.g_ci[bym][gj];
      const int ox = (int)sm.g_ox[bxm][gi] - 1, oy = (int)sm.g_ox[bym][gj] - 1;
      const int row = sm.prow[(ox + 1) * 3 + (oy + 1)][lane];
      const unsigned owned = __ballot_sync(0xffffffffu, row >= 0);
      if (owned == 0u) continue;
      // [region: phaseA build]
      // build: lane <-> run
      if (row >= 0) {
        float* h = myh + lane * HCfg<KP>::HS;
#pragma unroll
        for (int k = 0; k < 3 * KP; ++k) h[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < ECfg<KP>::USED; ++k) myE[k * TRE + lane] = 0.0f;
        const int4 e = sm.run[lane];
        const int len = e.y & 0xffff, tmin = e.z, so = sm.soff[lane];
        float* E = myE + ((tmin - 1) & 3) * TRE + lane;  // correction of position j sits at frame position j + s
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
          E[m * TRE] = fmaf(qf, D, E[m * TRE]);
          E[(m + 1) * TRE] = fmaf(qo, D, E[(m + 1) * TRE]);
        }
      }
      {  // re-lay the trains in place: [run][3 j + template] -> float4 [j][run]
        __syncwarp();
        float hv[3 * KP];
        if (row >= 0) {
#pragma unroll
          for (int k = 0; k < 3 * KP; ++k) hv[k] = myh[lane * HCfg<KP>::HS + k];
        }
        __syncwarp();
        if (row >= 0) {
#pragma unroll
          for (int j = 0; j < KP; ++j) reinterpret_cast<float4*>(myh)[j * TR + lane] = make_float4(hv[3 * j], hv[3 * j + 1], hv[3 * j + 2], 0.0f);
        }
      }
      __syncwarp();
      // [region: phaseA shift loop]
      // element offset of the run's frame start (fast paths) / of its row (low-end tiles)
      const int myoff = row >= 0 ? (int)((int64_t)row * A.wstride) + (pathA < 2 ? ((sm.run[lane].z - 1) & ~3) : 0) : 0;
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const unsigned todo = owned & sm.smask[s];
        if (todo == 0u) continue;
        const float* const tb = A.tms + (int64_t)s * A.nmrows * A.lps;
        const float* const rows[3] = {tb + (int64_t)((idx - 1) * 25 + bin) * A.lps, tb + (int64_t)(idx * 25 + bin) * A.lps,
                                      tb + (int64_t)((idx + 1) * 25 + bin) * A.lps};
        float W[3][NSV][WCfg<KP>::WN];
        load_w<NSV, 3, KP>(W, rows, lane);
        consume_main_npos<NSV, KP>(A, sm, W, todo, myoff, myh, myE, s, lane, pathA, npos);
      }
      __syncwarp();
    }
    // [region: phaseB pull]
    // ================= phase B: neighbour pixels, template 0, full segment charge (sim_jax.py:197-225,250-261) =========
    for (;;) {
      int p = 0;
      if (lane == 0) p = atomicAdd(&sm.next_run, 1);
      p = __shfl_sync(0xffffffffu, p, 0);
      if (p >= count) break;
      neigh_run_npos<NSV, KP>(A, sm, lk, p, myE, row0, lane, basepath, npos);
    }
  }
}

__global__ void k_reduce_row0(const float* __restrict__ row0, int ncopies, int r0stride, int nticks, float* __restrict__ wfs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nticks) return;
  float acc = 0.0f;
  for (int k = 0; k < ncopies; ++k) acc += row0[(int64_t)k * r0stride + c];
  if (acc != 0.0f) atomicAdd(wfs + c, acc);
}

static_assert(sizeof(TileSmem<4>) <= (227 / LARND_KP4_CTAS - 1) * 1024, "the 4-position tile kernel must fit LARND_KP4_CTAS CTAs per SM");

}  // namespace

size_t larnd_sorted_workspace_bytes(int64_t n) {
  size_t b = 0;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  b += align_up(nn * sizeof(int4), 256) * 2;                                       // runs_tmp, runs
  b += align_up((size_t)LARND_NCLS_MAX * 4 * sizeof(int), 256) * 2;                // class_count, cursor (4 shift sub-buckets)
  b += align_up((size_t)LARND_NCLS_MAX * sizeof(int), 256);                        // class_start
  b += align_up((nn / TR + LARND_NCLS_MAX + 1) * sizeof(int4), 256);               // tile_info
  b += 256 + 1024;                                                                 // counters + per-class-block totals (sorted_runs.cuh)
  b += align_up((size_t)LARND_ROW0_COPIES * LARND_ROW0_TICKS_MAX * sizeof(float), 256);  // private garbage rows
  return b;
}

void larnd_carve_sorted(char* p, int64_t n, Workspace* ws) {
  const size_t nn = (size_t)(n > 0 ? n : 1);
  ws->runs_tmp = p; p += align_up(nn * sizeof(int4), 256);
  ws->runs = p; p += align_up(nn * sizeof(int4), 256);
  ws->class_count = reinterpret_cast<int*>(p); p += align_up((size_t)LARND_NCLS_MAX * 4 * sizeof(int), 256);
  ws->cursor = reinterpret_cast<int*>(p); p += align_up((size_t)LARND_NCLS_MAX * 4 * sizeof(int), 256);
  ws->class_start = reinterpret_cast<int*>(p); p += align_up((size_t)LARND_NCLS_MAX * sizeof(int), 256);
  ws->tile_info = p; p += align_up((nn / TR + LARND_NCLS_MAX + 1) * sizeof(int4), 256);
  ws->gcnt = reinterpret_cast<int*>(p); p += 256 + 1024;
  ws->row0 = reinterpret_cast<float*>(p);
}

int larnd_sorted_supported(const larnd_params_t& p, const larnd_lut* lut) {
  const int nb = p.nb_sampling_bins_per_pixel;
  if (lut->ntpl * nb * nb * (SPAN_MAX_S + 1) > LARND_NCLS_MAX) return 0;
  if (lut->nsv == 0) return 0;                       // forward: frames of <= 256 ticks
  if (lut->L + 2 + SPAN_MAX_S > 32 * 6) return 0;    // backward kernel: 6 slots of 32 ticks
  if (p.n_ticks + 3 > LARND_ROW0_TICKS_MAX) return 0;
  if (p.number_pix_neighbors > 4) return 0;          // phase B: three neighbour units per lane
  return 1;
}

int larnd_launch_accumulate_sorted(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                   int32_t npix_capacity, int32_t flags, float* wfs, int64_t wfs_stride, const int32_t* counts,
                                   cudaStream_t st) {
  if (n == 0) return LARND_OK;
  SortArgs A;
  A.wfs = wfs;
  A.wstride = wfs_stride;
  A.v4ok = (reinterpret_cast<uintptr_t>(wfs) % 16 == 0) && wfs_stride % 4 == 0 && wfs_stride >= p.n_ticks + 3;
  A.r0stride = (p.n_ticks + 3 + 3) & ~3;
  A.skip_garbage = flags & LARND_FLAG_SKIP_GARBAGE;
  prof_begin(1, st);
  {
    int rc0 = sorted_fill_and_build(A, n, p, lut, ws, npix_capacity, counts, st);
    if (rc0) return rc0;
  }
  const int nsv = lut->nsv;
  const int split = sorted_split_mode(flags);
  const size_t smem4 = sizeof(TileSmem<4>), smem6 = sizeof(TileSmem<KPT>);
  static bool attr_done = false;
  if (!attr_done) {
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<1, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<2, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6));
#ifdef LARND_CARVEOUT   // experiment: force the shared-memory carveout (percent) to see how much the tile kernels depend on L1 capacity
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<1, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, LARND_CARVEOUT));
    LARND_CUDA(cudaFuncSetAttribute(k_acc_tiles<1, KPT>, cudaFuncAttributePreferredSharedMemoryCarveout, LARND_CARVEOUT));
#endif
    attr_done = true;
  }
  const int grid_small = sorted_grid(nsv == 1 ? LARND_KP4_CTAS : 2, LARND_ROW0_COPIES);  // KP = 4 variant
  const int grid_big = sorted_grid(nsv == 1 ? LARND_KP6_CTAS : 2, LARND_ROW0_COPIES);
  const int ncopies = max(grid_small, grid_big);
  if (!A.skip_garbage) LARND_CUDA(cudaMemsetAsync(ws.row0, 0, (size_t)ncopies * A.r0stride * sizeof(float), st));
  int big_lo = 0;  // first span left to the KPT kernel
  if (split >= 1) {
    if (nsv == 1) k_acc_tiles<1, 4><<<grid_small, TILE_THREADS, smem4, st>>>(A, 0, 2, 1);
    else k_acc_tiles<2, 4><<<grid_small, TILE_THREADS, smem4, st>>>(A, 0, 2, 1);
    LARND_LAUNCH_CHECK("k_acc_tiles<4>");
    big_lo = 3;
  }
  if (nsv == 1) k_acc_tiles<1, KPT><<<grid_big, TILE_THREADS, smem6, st>>>(A, big_lo, SPAN_MAX_S, 2);
  else k_acc_tiles<2, KPT><<<grid_big, TILE_THREADS, smem6, st>>>(A, big_lo, SPAN_MAX_S, 2);
  LARND_LAUNCH_CHECK("k_acc_tiles");
  if (!A.skip_garbage) {
    k_reduce_row0<<<(p.n_ticks + 255) / 256, 256, 0, st>>>(ws.row0, ncopies, A.r0stride, p.n_ticks, wfs);
    LARND_LAUNCH_CHECK("k_reduce_row0");
  }
  // segments whose window ends beyond the readout: per-segment path of accumulate.cu.  A window starts at T0 = nt - L - ct
  // (ct >= 0) and ends at T0 + L <= nt: with a response no longer than the readout (nt <= n_ticks - 2, seg_is_fast) there is
  // no such segment and the pass — 78 k CTAs that each find nothing to do at spill size — is not launched at all.
  int rc = LARND_OK;
  if (lut->nt > p.n_ticks - 2)
    rc = larnd_launch_accumulate(n, p, lut, ws, npix_capacity, flags | LARND_ACC_SLOW_ONLY, wfs, wfs_stride, counts, st);
  prof_end(1, st);
  return rc;
}
