// K4 MC-current ("parametrized") mode, forward and backward, sm_100a.
//
// Replaces the XLA lowering of simulate_parametrized's front half for number_pix_neighbors = 0, mc_diff = True
// (reference sim_jax.py:339-372): simulate_drift (:122-139) with generate_electrons (detsim_jax.py:376-400) and
// get_pixels (:477-491), the unique/pad/sort block (sim_jax.py:363-366), current_mc (detsim_jax.py:618-639) with
// current_model / current_model_diff (:545-615), integrated_expon[_diff] (:461-474,536-542), emg_pdf (:440-458),
// and accumulate_signals_parametrized (:207-228).  The FEE half is larnd_fee_forward.
//
// One warp per segment: lanes evaluate the 51 analytic current samples (SFU-bound erf/erfc/exp), the row
// normalisation of the diffusion variant is a 2-value warp broadcast, and the deposit is 51 consecutive
// red.global.add.f32 into the pixel's waveform row.  The backward kernel re-evaluates the current with
// forward-mode dual numbers in (t0, |dx|, |dy|, sigma_L) and contracts with the upstream gradient row.
#include "larnd_common.cuh"
#include "segment_physics.cuh"

namespace {

// record slots (reusing the generic per-segment record area of the workspace)
enum { M_Q = 0, M_T0F, M_XD, M_YD, M_SIG, M_TD, M_ST, M_SLCM, M_REC, M_XI, M_COS2, M_R0, M_R1, M_R2, M_T0FULL,
       M_TICK, M_PID, M_FLAGS, M_NF };
static_assert(M_NF <= LARND_NFIELDS, "record area too small");

constexpr int MC_NT = 51;  // int(5 / t_sampling) + 1 for t_sampling = 0.1
constexpr int MCP_THREADS = 128;

__global__ void __launch_bounds__(MCP_THREADS)
k_mc_prepare(const float* __restrict__ tracks, int64_t n, const __grid_constant__ larnd_columns_t cols,
             const __grid_constant__ larnd_params_t p, const float* __restrict__ rnd, float* __restrict__ rec,
             uint32_t* __restrict__ bitmap, int64_t n_words, int pid_offset, int32_t* __restrict__ counts) {
  extern __shared__ float srow[];
  const int ncols = cols.ncols;
  const int stride = ncols | 1;
  const int64_t base = (int64_t)blockIdx.x * MCP_THREADS;
  const int rows_here = (int)min((int64_t)MCP_THREADS, n - base);
  const int total = rows_here * ncols;
  const float* src = tracks + base * ncols;
  {
    const int dr = MCP_THREADS / ncols, dc = MCP_THREADS - dr * ncols;   // incremental (row, column): no division per element
    int r = threadIdx.x / ncols, c = threadIdx.x - r * ncols;
    for (int i = threadIdx.x; i < total; i += MCP_THREADS) {
      srow[r * stride + c] = __ldg(src + i);
      r += dr;
      c += dc;
      if (c >= ncols) { c -= ncols; ++r; }
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  int pid = 0;
  bool ok = false;
  if (t < rows_here) {
    const float* tr = srow + t * stride;
    const int64_t s = base + t;
    const SegPhys ph = segment_physics(tr, cols, p);
    const float r0 = rnd[s * 3 + 0], r1 = rnd[s * 3 + 1], r2 = rnd[s * 3 + 2];
    // generate_electrons: x += N(0,1)*sigma_T, y likewise, z += N(0,1)*sigma_L only if diffusion is NOT in the current model
    const float xe = fadd(ph.x, fmul(r0, ph.sT));
    const float ye = fadd(ph.y, fmul(r1, ph.sT));
    const float ze = p.diffusion_in_current_sim ? ph.z : fadd(ph.z, fmul(r2, ph.sl_cm));
    // get_pixels with n = 0
    const int px = (int)floor_divide_f(fsub(xe, p.tpc_borders[ph.plane][0][0]), p.pixel_pitch);
    const int py = (int)floor_divide_f(fsub(ye, p.tpc_borders[ph.plane][1][0]), p.pixel_pitch);
    const int ev = (int)tr[cols.eventID];
    pid = pixel2id_dev(px, py, ev * p.n_tpc + ph.plane, p.n_pixels_x, p.n_pixels_y);
    ok = true;
    // pixel centre from the id (id2pixel + get_pixel_coordinates; Python floor semantics also for id -1)
    const int nx = p.n_pixels_x, ny = p.n_pixels_y;
    const int xp = pid - floordiv_i(pid, nx) * nx;
    const int q1 = floordiv_i(pid, nx);
    const int yp = q1 - floordiv_i(q1, ny) * ny;
    const int q2 = floordiv_i(pid, nx * ny);
    const int pl = q2 - floordiv_i(q2, p.n_tpc) * p.n_tpc;
    const float xpix = fadd(__fmaf_rn((float)xp, p.pixel_pitch, p.tpc_borders[pl][0][0]), p.half_pitch);
    const float ypix = fadd(__fmaf_rn((float)yp, p.pixel_pitch, p.tpc_borders[pl][1][0]), p.half_pitch);
    const float dxs = fsub(xe, xpix), dys = fsub(ye, ypix);
    // current_mc timing
    const float dza = fsub(ze, ph.z_anode);
    const float t0 = fdiv(fabsf(dza), p.vdrift);
    const int tick = (int)fadd(fdiv(t0, p.t_sampling), 0.5f);
    const float t0f = fsub(t0, fmul((float)tick, p.t_sampling));
    int flags = (ph.inside ? 1 : 0) | (fsub(ph.z, ph.z_anode) > 0.f ? 2 : 0) | (dza > 0.f ? 4 : 0) | (dxs > 0.f ? 8 : 0) |
                (dys > 0.f ? 16 : 0);
    rec[(int64_t)M_Q * n + s] = ph.q;
    rec[(int64_t)M_T0F * n + s] = t0f;
    rec[(int64_t)M_XD * n + s] = fabsf(dxs);
    rec[(int64_t)M_YD * n + s] = fabsf(dys);
    rec[(int64_t)M_SIG * n + s] = fdiv(ph.sl_cm, p.vdrift);
    rec[(int64_t)M_TD * n + s] = ph.td;
    rec[(int64_t)M_ST * n + s] = ph.sT;
    rec[(int64_t)M_SLCM * n + s] = ph.sl_cm;
    rec[(int64_t)M_REC * n + s] = ph.recomb;
    rec[(int64_t)M_XI * n + s] = ph.xi;
    rec[(int64_t)M_COS2 * n + s] = ph.cos2;
    rec[(int64_t)M_R0 * n + s] = r0;
    rec[(int64_t)M_R1 * n + s] = r1;
    rec[(int64_t)M_R2 * n + s] = r2;
    rec[(int64_t)M_T0FULL * n + s] = t0;
    int* irec = reinterpret_cast<int*>(rec);
    irec[(int64_t)M_TICK * n + s] = tick;
    irec[(int64_t)M_PID * n + s] = pid;
    irec[(int64_t)M_FLAGS * n + s] = flags;
  }
  unsigned live = __ballot_sync(0xffffffffu, ok);
  if (ok) {
    unsigned peers = __match_any_sync(live, pid);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
      long long bidx = (long long)pid + pid_offset;
      if (bidx < 0 || (bidx >> 5) >= n_words) atomicOr(counts + 2, 2);
      else {
        uint32_t bit = 1u << (bidx & 31);
        if (!(__ldg(bitmap + (bidx >> 5)) & bit)) atomicOr(bitmap + (bidx >> 5), bit);
      }
    }
  }
}

// ---- the analytic current model, generic over plain floats and forward-mode duals ---------------------
struct D4 {  // value + derivatives w.r.t. (t0, xd, yd, sigma)
  float v, d[4];
};
__device__ __forceinline__ D4 mk(float v) { return D4{v, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ D4 var(float v, int i) { D4 r = mk(v); r.d[i] = 1.f; return r; }
#define D4_BIN(op, expr_v, expr_d)                                            \
  __device__ __forceinline__ D4 operator op(const D4& a, const D4& b) {       \
    D4 r; r.v = expr_v;                                                       \
    _Pragma("unroll") for (int i = 0; i < 4; ++i) r.d[i] = expr_d;            \
    return r; }
D4_BIN(+, a.v + b.v, a.d[i] + b.d[i])
D4_BIN(-, a.v - b.v, a.d[i] - b.d[i])
D4_BIN(*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
D4_BIN(/, a.v / b.v, (a.d[i] * b.v - a.v * b.d[i]) / (b.v * b.v))
__device__ __forceinline__ D4 operator*(float s, const D4& a) { D4 r; r.v = s * a.v; for (int i = 0; i < 4; ++i) r.d[i] = s * a.d[i]; return r; }
__device__ __forceinline__ D4 operator+(float s, const D4& a) { D4 r = a; r.v += s; return r; }
__device__ __forceinline__ D4 operator-(float s, const D4& a) { D4 r; r.v = s - a.v; for (int i = 0; i < 4; ++i) r.d[i] = -a.d[i]; return r; }
__device__ __forceinline__ D4 operator/(float s, const D4& a) { D4 r; r.v = s / a.v; for (int i = 0; i < 4; ++i) r.d[i] = -s * a.d[i] / (a.v * a.v); return r; }
__device__ __forceinline__ D4 operator-(const D4& a) { D4 r; r.v = -a.v; for (int i = 0; i < 4; ++i) r.d[i] = -a.d[i]; return r; }
__device__ __forceinline__ D4 chain(const D4& a, float f, float df) { D4 r; r.v = f; for (int i = 0; i < 4; ++i) r.d[i] = df * a.d[i]; return r; }
__device__ __forceinline__ D4 g_exp(const D4& a) { float e = expf(a.v); return chain(a, e, e); }
__device__ __forceinline__ D4 g_erf(const D4& a) { return chain(a, erff(a.v), 1.12837917f * expf(-a.v * a.v)); }
__device__ __forceinline__ D4 g_erfc(const D4& a) { return chain(a, erfcf(a.v), -1.12837917f * expf(-a.v * a.v)); }
__device__ __forceinline__ D4 g_min0(const D4& a) { return a.v < 0.f ? a : mk(0.f); }  // jnp.minimum(0., a)
__device__ __forceinline__ D4 g_min1(const D4& a) { return a.v < 1.f ? a : mk(1.f); }  // jnp.minimum(a, 1)
__device__ __forceinline__ float g_exp(float a) { return expf(a); }
__device__ __forceinline__ float g_erf(float a) { return erff(a); }
__device__ __forceinline__ float g_erfc(float a) { return erfcf(a); }
__device__ __forceinline__ float g_min0(float a) { return fminf(0.f, a); }
__device__ __forceinline__ float g_min1(float a) { return fminf(a, 1.f); }

__device__ __forceinline__ float lift(float v, const float&) { return v; }
__device__ __forceinline__ D4 lift(float v, const D4&) { return mk(v); }

template <typename T>
__device__ __forceinline__ T quad(const float (&c)[6], const T& x, const T& y) {
  return c[0] + c[1] * x + c[2] * y + c[3] * (x * y) + c[4] * (x * x) + c[5] * (y * y);
}

// 0.5*erf((e-loc)/(sqrt2 diff)) - emg_pdf(e, loc, diff, lam)/lam at a bin edge e (detsim_jax.py:440-458,467-468)
template <typename T>
__device__ __forceinline__ T expon_diff_edge(float e, const T& loc, const T& lam, const T& diff) {
  const T ls2 = lam * diff * diff;
  const T er = g_erf((e - loc) / (1.41421354f * diff));
  const T expo = (0.5f * lam) * (2.0f * loc + ls2 - lift(2.0f * e, loc));
  const T ec = g_erfc((loc + ls2 - lift(e, loc)) / (1.41421354f * diff));
  // emg/lam = 0.5 * exp(expo) * erfc(...)
  return 0.5f * er - 0.5f * (g_exp(expo) * ec);
}

// integrated_expon_diff(-t, loc, scale, diff, dt)[tick] (detsim_jax.py:461-474), rows normalised over the 51 ticks
template <typename T>
__device__ __forceinline__ T expon_diff(float t, const T& loc, const T& scale, const T& diff, float dtk) {
  const T lam = 1.0f / scale;
  const float x = -t, half = 0.5f * dtk;
  const T up = expon_diff_edge(x + half, loc, lam, diff);
  const T lo = expon_diff_edge(x - half, loc, lam, diff);
  const T lo_first = expon_diff_edge(0.0f - half, loc, lam, diff);                    // lower_values[..., 0]   (t = 0)
  const T up_last = expon_diff_edge(-(dtk * (MC_NT - 1)) + half, loc, lam, diff);     // upper_values[..., -1]  (t = 5)
  return ((up - lo) / (lo_first - up_last)) / lift(dtk, loc);
}

// integrated_expon(-t, loc, scale, dt)[tick] (detsim_jax.py:536-542)
template <typename T>
__device__ __forceinline__ T expon_plain(float t, const T& loc, const T& scale, float dtk) {
  const float x = -t, half = 0.5f * dtk;
  const T e1 = g_exp(g_min0((loc - lift(x - half, loc)) / scale));
  const T e2 = g_exp(g_min0((loc - lift(x + half, loc)) / scale));
  const T e3 = g_exp(g_min0((loc - lift(half, loc)) / scale));
  return (e1 - e2 + e3 / lift((float)MC_NT, loc)) / lift(dtk, loc);
}

// current_model / current_model_diff (detsim_jax.py:545-615) at tick time t, per unit charge
template <typename T>
__device__ __forceinline__ T current_sample(float t, const T& t0f, const T& xd, const T& yd, const T& sig, bool diffusion, float dtk) {
  const float Bp[6] = {1.060f, -0.909f, -0.909f, 5.856f, 0.207f, 0.207f};
  const float Cp[6] = {0.679f, -1.083f, -1.083f, 8.772f, -5.521f, -5.521f};
  const float Dp[6] = {2.644f, -9.174f, -9.174f, 13.483f, 45.887f, 45.887f};
  const float Tp[6] = {2.948f, -2.705f, -2.705f, 4.825f, 20.814f, 20.814f};
  const T a = g_min1(quad(Bp, xd, yd));
  const T b = quad(Cp, xd, yd);
  const T c = quad(Dp, xd, yd);
  const T loc = -(t0f + quad(Tp, xd, yd));   // -shifted_t0
  if (diffusion) return a * expon_diff(t, loc, b, sig, dtk) + (1.0f - a) * expon_diff(t, loc, c, sig, dtk);
  return a * expon_plain(t, loc, b, dtk) + (1.0f - a) * expon_plain(t, loc, c, dtk);
}

struct McArgs {
  const float* rec;
  int64_t n;
  RowLookup lk;
  const int32_t* counts;
  int nticks;
  float* wfs;
  const float* g;
  int64_t g_stride;
  float* partials;
};

constexpr int MC_WARPS = 8;
constexpr int MC_EDGES = MC_NT + 1;         // bin edges of the 51 samples
constexpr int MC_SEGS = 4;                  // segments per CTA of k_mc_accumulate: 4 x 52 = 208 of 224 threads busy
constexpr int MC_ACC_THREADS = 224;

// current_mc + accumulate_signals_parametrized (detsim_jax.py:618-639, 207-228): thread <-> (segment, bin edge).  tick =
// t0_tick - 51 + k; < 0 or >= Nticks-1 -> column 0, else +1.  Every sample is a difference of ONE edge function at the two
// edges of its bin (integrated_expon_diff :461-474: upper_k - lower_k, rows normalised by lower_0 - upper_50;
// integrated_expon :536-542), and lower_k is upper_(k+1): the 52 edges e_j = dt/2 - t_j of a segment are evaluated once —
// thread j, both exponential components — and handed over in shared memory; sample k then takes edges k and k + 1 and the
// normalisation edges 1 and 50.  The first versions evaluated upper and lower edge per sample (4 edge evaluations per
// sample, 955 instructions per sample at 90 % issue utilisation); the shared form evaluates 52 / 51 per sample.  The lower
// edge of sample k thereby becomes fl(dt/2 - t_(k+1)) instead of fl(-t_k - dt/2): an ulp of the argument, well inside the
// waveform tolerance.
__global__ void __launch_bounds__(MC_ACC_THREADS)
k_mc_accumulate(const __grid_constant__ McArgs A, const __grid_constant__ larnd_params_t p) {
  __shared__ float s_E[MC_SEGS][MC_EDGES][2];
  __shared__ float s_garbage[MC_SEGS];
  const int sl = threadIdx.x / MC_EDGES, k = threadIdx.x - sl * MC_EDGES;
  const int64_t s = (int64_t)blockIdx.x * MC_SEGS + sl;
  const bool diffusion = p.diffusion_in_current_sim != 0;
  const float dtk = 5.0f / (MC_NT - 1);
  bool live = A.counts[2] == 0 && sl < MC_SEGS && s < A.n;
  int row = -1, start = 0;
  float q = 0.f, t0f = 0.f, xd = 0.f, yd = 0.f, sig = 1.f;
  if (live) {
    RowLookup lk = A.lk;
    lk.n_unique = A.counts[0];
    lk.n_neg = A.counts[1];
    const int* irec = reinterpret_cast<const int*>(A.rec);
    const int64_t n = A.n;
    row = lookup_row(lk, irec[(int64_t)M_PID * n + s]);
    live = row >= 0;  // cannot fail: every id was inserted by k_mc_prepare
    q = A.rec[(int64_t)M_Q * n + s];
    t0f = A.rec[(int64_t)M_T0F * n + s]; xd = A.rec[(int64_t)M_XD * n + s]; yd = A.rec[(int64_t)M_YD * n + s];
    sig = A.rec[(int64_t)M_SIG * n + s];
    start = irec[(int64_t)M_TICK * n + s] - MC_NT;
  }
  if (threadIdx.x < MC_SEGS) s_garbage[threadIdx.x] = 0.f;
  const float Bp[6] = {1.060f, -0.909f, -0.909f, 5.856f, 0.207f, 0.207f};
  const float Cp[6] = {0.679f, -1.083f, -1.083f, 8.772f, -5.521f, -5.521f};
  const float Dp[6] = {2.644f, -9.174f, -9.174f, 13.483f, 45.887f, 45.887f};
  const float Tp[6] = {2.948f, -2.705f, -2.705f, 4.825f, 20.814f, 20.814f};
  const float a = g_min1(quad(Bp, xd, yd));
  const float b = quad(Cp, xd, yd), c = quad(Dp, xd, yd);
  const float loc = -(t0f + quad(Tp, xd, yd));
  const float half = 0.5f * dtk;
  if (live) {
    // jnp.linspace(0, 5, 51)[j] = 0*(1-s) + 5*s with s = j/50 (exact end point); edge 51 lies one bin beyond
    const float sfrac = __fdiv_rn((float)k, (float)(MC_NT - 1));
    const float t = (k == MC_NT - 1) ? 5.0f : __fmul_rn(5.0f, sfrac);
    const float e = -t + half;   // upper edge of sample k = lower edge of sample k - 1
    if (diffusion) {
      s_E[sl][k][0] = expon_diff_edge(e, loc, 1.0f / b, sig);
      s_E[sl][k][1] = expon_diff_edge(e, loc, 1.0f / c, sig);
    } else {
      s_E[sl][k][0] = g_exp(g_min0((loc - e) / b));
      s_E[sl][k][1] = g_exp(g_min0((loc - e) / c));
    }
  }
  __syncthreads();
  if (live && k < MC_NT) {
    const float ub = s_E[sl][k][0], lb = s_E[sl][k + 1][0], uc = s_E[sl][k][1], lc = s_E[sl][k + 1][1];
    float cur;
    if (diffusion) {
      const float den_b = s_E[sl][1][0] - s_E[sl][MC_NT - 1][0], den_c = s_E[sl][1][1] - s_E[sl][MC_NT - 1][1];
      cur = a * (((ub - lb) / den_b) / dtk) + (1.0f - a) * (((uc - lc) / den_c) / dtk);
    } else {
      const float inv = (float)MC_NT;
      cur = a * ((lb - ub + s_E[sl][0][0] / inv) / dtk) + (1.0f - a) * ((lc - uc + s_E[sl][0][1] / inv) / dtk);
    }
    cur *= q;
    float* base = A.wfs + (int64_t)row * A.nticks;
    const int tick = start + k;
    if (tick < 0 || tick >= A.nticks - 1) atomicAdd(&s_garbage[sl], cur);  // rare: windows sticking out of the readout
    else atomicAdd(base + tick + 1, cur);
  }
  __syncthreads();
  if (live && k == 0 && s_garbage[sl] != 0.f) atomicAdd(A.wfs + (int64_t)row * A.nticks, s_garbage[sl]);
}

__global__ void __launch_bounds__(MC_WARPS * 32)
k_mc_backward(const __grid_constant__ McArgs A, const __grid_constant__ larnd_params_t p) {
  __shared__ float s_grad[MC_WARPS][16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * MC_WARPS + wid;
  float gacc[LARND_NPARAMS];
#pragma unroll
  for (int k = 0; k < LARND_NPARAMS; ++k) gacc[k] = 0.f;
  const bool live = (A.counts[2] == 0) && s < A.n;
  if (live) {
    RowLookup lk = A.lk;
    lk.n_unique = A.counts[0];
    lk.n_neg = A.counts[1];
    const int* irec = reinterpret_cast<const int*>(A.rec);
    const int64_t n = A.n;
    const float q = A.rec[(int64_t)M_Q * n + s];
    const int pid = irec[(int64_t)M_PID * n + s];
    const int flags = irec[(int64_t)M_FLAGS * n + s];
    const int row = lookup_row(lk, pid);
    if (row >= 0 && (flags & 1) && pid >= 0) {
      const float t0f = A.rec[(int64_t)M_T0F * n + s], xd = A.rec[(int64_t)M_XD * n + s], yd = A.rec[(int64_t)M_YD * n + s];
      const float sig = A.rec[(int64_t)M_SIG * n + s];
      const int start = irec[(int64_t)M_TICK * n + s] - MC_NT;
      const float dtk = 5.0f / (MC_NT - 1);
      const float* grow = A.g + (int64_t)row * A.g_stride;
      float dq = 0.f, din[4] = {0.f, 0.f, 0.f, 0.f};
      const bool use_sig = p.diffusion_in_current_sim != 0;
      // Every sample of the current is a difference of ONE edge function at consecutive bin edges
      //   e_j = dt/2 - t_j, j = 0 .. 51:   up(k) = G_k,  lo(k) = G_(k+1),  row normalisation D = G_1 - G_50     (diffusion variant,
      //   detsim_jax.py:461-474);          e1(k) = G_(k+1),  e2(k) = G_k,  e3 = G_0                             (plain, :536-542),
      // so the loss L = sum_k g_k q cur_k is a LINEAR form in the 2 x 52 edge values (times 1 / D): the warp evaluates every edge
      // once (lane <-> edge, two passes) together with its three closed-form partials (d/dloc, d/dscale, d/dsigma), forms the
      // scalar coefficient dL/dG_j of its edges from the neighbouring ticks' upstream gradients, and reduces five sums.  (The
      // first version pushed 4-direction dual numbers through every operation of every sample: 25 ms per 2 M segments; duals
      // on shared edges: 5.1 ms.)
      const float Bp[6] = {1.060f, -0.909f, -0.909f, 5.856f, 0.207f, 0.207f};
      const float Cp[6] = {0.679f, -1.083f, -1.083f, 8.772f, -5.521f, -5.521f};
      const float Dp[6] = {2.644f, -9.174f, -9.174f, 13.483f, 45.887f, 45.887f};
      const float Tp[6] = {2.948f, -2.705f, -2.705f, 4.825f, 20.814f, 20.814f};
      auto quad_dx = [&](const float (&c)[6]) { return c[1] + c[3] * yd + 2.0f * c[4] * xd; };
      auto quad_dy = [&](const float (&c)[6]) { return c[2] + c[3] * xd + 2.0f * c[5] * yd; };
      const float qa = quad(Bp, xd, yd);
      const float ca = fminf(qa, 1.0f);
      const float scale[2] = {quad(Cp, xd, yd), quad(Dp, xd, yd)};
      const float loc = -(t0f + quad(Tp, xd, yd));
      const float half = 0.5f * dtk;
      float gvk[2], G[2][2], Gl[2][2], Gs[2][2], Gg[2][2];   // [pass][component]: edge value, d/dloc, d/dscale, d/dsigma
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const int j = lane + 32 * ps;
        const int tick = start + j;
        gvk[ps] = (j < MC_NT && tick >= 0 && tick < A.nticks - 1) ? __ldg(grow + tick + 1) : 0.0f;   // else: garbage column
#pragma unroll
        for (int cpt = 0; cpt < 2; ++cpt) { G[ps][cpt] = 0.f; Gl[ps][cpt] = 0.f; Gs[ps][cpt] = 0.f; Gg[ps][cpt] = 0.f; }
        if (j <= MC_NT) {
          const float e = half - __fmul_rn(5.0f, __fdiv_rn((float)j, (float)(MC_NT - 1)));
          const float d = loc - e;
#pragma unroll
          for (int cpt = 0; cpt < 2; ++cpt) {
            const float sc = scale[cpt];
            if (use_sig) {
              const float lam = 1.0f / sc, is2 = 1.0f / (1.41421354f * sig), ls2 = lam * sig * sig;
              const float u = -d * is2, E = lam * (d + 0.5f * ls2), v = (d + ls2) * is2;
              const float X = expf(E) * erfcf(v);
              const float Au = 0.56418958f * expf(-u * u), Bv = 0.56418958f * expf(E - v * v);
              G[ps][cpt] = 0.5f * erff(u) - 0.5f * X;
              Gl[ps][cpt] = (Bv - Au) * is2 - 0.5f * X * lam;
              const float Glam = -0.5f * X * (d + ls2) + Bv * sig * 0.70710678f;
              Gs[ps][cpt] = -lam * lam * Glam;
              Gg[ps][cpt] = -Au * u / sig - 0.5f * X * lam * lam * sig + Bv * (lam * 0.70710678f - d * is2 / sig);
            } else {
              const float arg = d / sc;
              if (arg < 0.0f) {
                const float F = expf(arg);
                G[ps][cpt] = F; Gl[ps][cpt] = F / sc; Gs[ps][cpt] = -F * arg / sc;
              } else {
                G[ps][cpt] = 1.0f;
              }
            }
          }
        }
      }
      auto wsum = [](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
      };
      // neighbours: edge j + 1 (values) for the samples, upstream gradient of tick j - 1 for the edge coefficients
      float Gn[2][2], gprev[2];
#pragma unroll
      for (int cpt = 0; cpt < 2; ++cpt) {
        const float first1 = __shfl_sync(0xffffffffu, G[1][cpt], 0);
        Gn[0][cpt] = __shfl_down_sync(0xffffffffu, G[0][cpt], 1);
        if (lane == 31) Gn[0][cpt] = first1;
        Gn[1][cpt] = __shfl_down_sync(0xffffffffu, G[1][cpt], 1);
      }
      {
        const float last0 = __shfl_sync(0xffffffffu, gvk[0], 31);
        gprev[0] = __shfl_up_sync(0xffffffffu, gvk[0], 1);
        if (lane == 0) gprev[0] = 0.0f;
        gprev[1] = __shfl_up_sync(0xffffffffu, gvk[1], 1);
        if (lane == 0) gprev[1] = last0;
      }
      float Dn[2], G0[2], S[2];   // per component: normalisation, edge 0, sum_k g_k C_k
#pragma unroll
      for (int cpt = 0; cpt < 2; ++cpt) {
        Dn[cpt] = __shfl_sync(0xffffffffu, G[0][cpt], 1) - __shfl_sync(0xffffffffu, G[1][cpt], MC_NT - 1 - 32);
        G0[cpt] = __shfl_sync(0xffffffffu, G[0][cpt], 0);
        float part = 0.0f;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const float C = use_sig ? ((G[ps][cpt] - Gn[ps][cpt]) / Dn[cpt]) / dtk : (Gn[ps][cpt] - G[ps][cpt] + G0[cpt] / (float)MC_NT) / dtk;
          part = fmaf(gvk[ps], C, part);   // gvk is zero for the lanes without a tick
        }
        S[cpt] = wsum(part);
      }
      const float gsum = use_sig ? 0.0f : wsum(gvk[0] + gvk[1]);
      dq = ca * S[0] + (1.0f - ca) * S[1];
      // coefficient of edge j in L, then the four parameter sums
      float s_loc = 0.f, s_sc[2] = {0.f, 0.f}, s_sig = 0.f;
#pragma unroll
      for (int cpt = 0; cpt < 2; ++cpt) {
        const float wa = (cpt == 0 ? ca : 1.0f - ca) * q;   // weight of the component times the charge
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int j = lane + 32 * ps;
          float coef;
          if (use_sig) {
            coef = wa * (gvk[ps] - gprev[ps]) / (Dn[cpt] * dtk);
            if (j == 1) coef -= wa * S[cpt] / Dn[cpt];
            if (j == MC_NT - 1) coef += wa * S[cpt] / Dn[cpt];
          } else {
            coef = wa * (gprev[ps] - gvk[ps]) / dtk;
            if (j == 0) coef += wa * gsum / ((float)MC_NT * dtk);
          }
          if (j > MC_NT) coef = 0.0f;
          s_loc = fmaf(coef, Gl[ps][cpt], s_loc);
          s_sc[cpt] = fmaf(coef, Gs[ps][cpt], s_sc[cpt]);
          s_sig = fmaf(coef, Gg[ps][cpt], s_sig);
        }
      }
      s_loc = wsum(s_loc); s_sc[0] = wsum(s_sc[0]); s_sc[1] = wsum(s_sc[1]); s_sig = wsum(s_sig);
      const float dLda = (qa < 1.0f) ? q * (S[0] - S[1]) : 0.0f;   // a = min(quad, 1)
      din[0] = -s_loc;
      din[1] = -s_loc * quad_dx(Tp) + s_sc[0] * quad_dx(Cp) + s_sc[1] * quad_dx(Dp) + dLda * quad_dx(Bp);
      din[2] = -s_loc * quad_dy(Tp) + s_sc[0] * quad_dy(Cp) + s_sc[1] * quad_dy(Dp) + dLda * quad_dy(Bp);
      din[3] = use_sig ? s_sig : 0.0f;
      if (lane == 0) {
        const float td = A.rec[(int64_t)M_TD * n + s], sT = A.rec[(int64_t)M_ST * n + s], slcm = A.rec[(int64_t)M_SLCM * n + s];
        const float recb = A.rec[(int64_t)M_REC * n + s], xi = A.rec[(int64_t)M_XI * n + s], cos2 = A.rec[(int64_t)M_COS2 * n + s];
        const float r0 = A.rec[(int64_t)M_R0 * n + s], r1 = A.rec[(int64_t)M_R1 * n + s], r2 = A.rec[(int64_t)M_R2 * n + s];
        const float t0 = A.rec[(int64_t)M_T0FULL * n + s];
        const float v = p.vdrift, tau = p.lifetime;
        const float sgn_a = (flags & 2) ? 1.f : -1.f;   // sign(z - z_anode) of the segment (drift time)
        const float sgn_e = (flags & 4) ? 1.f : -1.f;   // sign(z_electron - z_anode)
        const float sgn_x = (flags & 8) ? 1.f : -1.f, sgn_y = (flags & 16) ? 1.f : -1.f;
        // inputs of the current: t0f = |ze - za|/v - tick*ts ; xd = |xe - xpix| ; yd ; sig = slcm / v
        const float g_ze = din[0] * sgn_e / v;
        const float g_xe = din[1] * sgn_x, g_ye = din[2] * sgn_y;
        float g_slcm = din[3] / v + (p.diffusion_in_current_sim ? 0.f : g_ze * r2);
        const float g_sT = g_xe * r0 + g_ye * r1;
        const float g_td = dq * (-q / tau) + (td > 0.f ? (g_slcm * slcm + g_sT * sT) / (2.f * td) : 0.f);
        const float g_v = g_td * (-td / v) + din[0] * (-t0 / v) + din[3] * (-(slcm / v) / v);
        gacc[LARND_P_SHIFT_Z] += g_td * (-sgn_a / v) - g_ze;
        gacc[LARND_P_SHIFT_X] += -g_xe;
        gacc[LARND_P_SHIFT_Y] += -g_ye;
        gacc[LARND_P_LIFETIME] += dq * q * td / (tau * tau);
        if (p.long_diff > 0.f) gacc[LARND_P_LONG_DIFF] += g_slcm * slcm / (2.f * p.long_diff);
        if (p.tran_diff > 0.f) gacc[LARND_P_TRAN_DIFF] += g_sT * sT / (2.f * p.tran_diff);
        gacc[LARND_P_MEV_TO_ELECTRONS] += dq * q / p.MeVToElectrons;
        float g_rec = (recb != 0.f) ? dq * q / recb : 0.f;
        float g_E = g_v * p.dvdrift_dEfield;
        if (p.recombination_mode == 2) {
          const float dn = 1.f + xi;
          gacc[LARND_P_AB] += g_rec * recb / p.Ab;
          const float g_xi = g_rec * (-recb / dn);
          if (p.kb != 0.f) gacc[LARND_P_KB] += g_xi * xi / p.kb;
          g_E += g_xi * (-xi / p.eField);
          gacc[LARND_P_LAR_DENSITY] += g_xi * (-xi / p.lArDensity);
        } else if (recb > 0.f) {
          const float den = (p.recombination_mode == 3) ? xi + 1e-10f : xi;
          const float lg = logf(p.alpha + xi);
          gacc[LARND_P_ALPHA] += g_rec / ((p.alpha + xi) * den);
          const float g_xi = g_rec * (1.f / ((p.alpha + xi) * den) - lg / (den * den));
          gacc[LARND_P_BETA] += g_xi * xi / p.beta;
          g_E += g_xi * (-xi / p.eField);
          gacc[LARND_P_LAR_DENSITY] += g_xi * (-xi / p.lArDensity);
          if (p.recombination_mode == 3) {
            const float gg = 1.f - cos2 + p.inv_R2 * cos2;
            gacc[LARND_P_R_PARAM] += g_xi * xi * cos2 / (p.R_param * p.R_param * p.R_param * gg);
          }
        }
        gacc[LARND_P_EFIELD] += g_E;
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < LARND_NPARAMS; ++k) s_grad[wid][k] = gacc[k];
  }
  __syncthreads();
  if (threadIdx.x < LARND_NPARAMS) {
    float v = 0.f;
    for (int w = 0; w < MC_WARPS; ++w) v += s_grad[w][threadIdx.x];
    A.partials[(int64_t)blockIdx.x * 16 + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) k_mc_reduce_partials(const float* __restrict__ partials, int64_t n_blocks,
                                                            float* __restrict__ grad) {
  __shared__ double sm[256];
  const int pidx = blockIdx.x;
  double acc = 0.0;
  for (int64_t c = threadIdx.x; c < n_blocks; c += 256) acc += (double)partials[c * 16 + pidx];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) grad[pidx] += (float)sm[0];
}


// ---- stream-form operators: current_mc (detsim_jax.py:619-639) and accumulate_signals_parametrized (:209-228) with the
// reference's own argument lists (electrons (N, ncols) + pixel centres in, (N, 51) currents out), for code that calls
// the stages one by one.  The fused kernels above never materialise the (N, 51) array.
struct CurArgs {
  const float* electrons; int64_t n; int ncols;
  int cx, cy, cz, cld, cq, cplane;
  const float* pixc;       // (n, 2)
  int32_t* t0_tick;        // (n)
  float* signals;          // (n, 51)
  const float* g_signals;  // backward
  float* g_electrons;      // (n, ncols), columns x, y, z, long_diff, n_electrons written
  float* g_pixc;           // (n, 2)
};

template <bool BWD>
__global__ void __launch_bounds__(MC_WARPS * 32)
k_current_mc(const __grid_constant__ CurArgs A, const __grid_constant__ larnd_params_t p) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * MC_WARPS + (threadIdx.x >> 5);
  if (s >= A.n) return;
  const float* el = A.electrons + s * A.ncols;
  const float x = __ldg(el + A.cx), y = __ldg(el + A.cy), z = __ldg(el + A.cz);
  const float q = __ldg(el + A.cq);
  const int plane = max(0, min((int)__ldg(el + A.cplane), p.n_tpc - 1));
  const float dxs = fsub(x, __ldg(A.pixc + 2 * s)), dys = fsub(y, __ldg(A.pixc + 2 * s + 1));
  const float dza = fsub(z, p.tpc_borders[plane][2][0]);
  const float t0 = fdiv(fabsf(dza), p.vdrift);
  const int tick = (int)fadd(fdiv(t0, p.t_sampling), 0.5f);
  const float t0f = fsub(t0, fmul((float)tick, p.t_sampling));
  const float xd = fabsf(dxs), yd = fabsf(dys);
  const float sig = fdiv(__ldg(el + A.cld), p.vdrift);
  const float dtk = 5.0f / (MC_NT - 1);
  const bool use_sig = p.diffusion_in_current_sim != 0;
  if (!BWD) {
    if (lane == 0) A.t0_tick[s] = tick;
    for (int k = lane; k < MC_NT; k += 32) {
      const float sfrac = __fdiv_rn((float)k, (float)(MC_NT - 1));
      const float t = (k == MC_NT - 1) ? 5.0f : __fmul_rn(5.0f, sfrac);
      A.signals[s * MC_NT + k] = current_sample<float>(t, t0f, xd, yd, sig, use_sig, dtk) * q;
    }
  } else {
    float dq = 0.f, din[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < MC_NT; k += 32) {
      const float gv = __ldg(A.g_signals + s * MC_NT + k);
      const float sfrac = __fdiv_rn((float)k, (float)(MC_NT - 1));
      const float t = (k == MC_NT - 1) ? 5.0f : __fmul_rn(5.0f, sfrac);
      const D4 cur = current_sample<D4>(t, var(t0f, 0), var(xd, 1), var(yd, 2), use_sig ? var(sig, 3) : mk(sig), use_sig, dtk);
      dq = fmaf(gv, cur.v, dq);
#pragma unroll
      for (int i = 0; i < 4; ++i) din[i] = fmaf(gv * q, cur.d[i], din[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dq += __shfl_xor_sync(0xffffffffu, dq, o);
#pragma unroll
      for (int i = 0; i < 4; ++i) din[i] += __shfl_xor_sync(0xffffffffu, din[i], o);
    }
    if (lane == 0) {
      float* ge = A.g_electrons + s * A.ncols;
      const float gx = din[1] * (dxs > 0.f ? 1.f : (dxs < 0.f ? -1.f : 0.f));
      const float gy = din[2] * (dys > 0.f ? 1.f : (dys < 0.f ? -1.f : 0.f));
      ge[A.cx] = gx;
      ge[A.cy] = gy;
      ge[A.cz] = din[0] * (dza > 0.f ? 1.f : (dza < 0.f ? -1.f : 0.f)) / p.vdrift;
      ge[A.cld] = din[3] / p.vdrift;
      ge[A.cq] = dq;
      A.g_pixc[2 * s] = -gx;
      A.g_pixc[2 * s + 1] = -gy;
    }
  }
}

// wfs.at[flat].add(signals): flat = pixID*Nticks + tick rule; negative flat ids wrap like numpy, ids past the end are dropped
template <bool BWD>
__global__ void __launch_bounds__(256)
k_accumulate_parametrized(float* __restrict__ wfs, const float* __restrict__ g_wfs, float* __restrict__ sig,
                          const int32_t* __restrict__ pix, const int32_t* __restrict__ start, int64_t n, int nsig, int npix, int nticks) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n * nsig) return;
  const int64_t s = i / nsig;
  const int k = (int)(i - s * nsig);
  int tt = __ldg(start + s) + k;
  tt = (tt < 0 || tt >= nticks - 1) ? 0 : tt + 1;
  const int64_t total = (int64_t)npix * nticks;
  int64_t flat = (int64_t)__ldg(pix + s) * nticks + tt;
  if (flat < 0) flat += total;
  const bool ok = flat >= 0 && flat < total;
  if (!BWD) { if (ok) atomicAdd(wfs + flat, sig[i]); }
  else sig[i] = ok ? __ldg(g_wfs + flat) : 0.f;
}

int mc_check(const larnd_params_t* p, const larnd_columns_t* cols, const float* rnd) {
  if (!p || !cols || !rnd) { larnd_set_error("larnd_mc: null argument"); return LARND_E_ARG; }
  if (p->number_pix_neighbors != 0) {
    larnd_set_error("the MC-current path only supports number_pix_neighbors = 0 (reference current_mc broadcasts (N,) electrons "
                    "against (N*P^2,) pixels, detsim_jax.py:623-624)");
    return LARND_E_ARG;
  }
  if (p->n_tpc < 1 || p->n_tpc > LARND_MAX_TPC) { larnd_set_error("n_tpc unsupported"); return LARND_E_ARG; }
  int nt = (int)(5.0 / (double)p->t_sampling + 1e-4) + 1;  // Python evaluates int(5/0.1)+1 in double
  if (nt != MC_NT) { larnd_set_error("MC-current mode is built for t_sampling = 0.1 us (51 ticks), got %d ticks", nt); return LARND_E_ARG; }
  return LARND_OK;
}

}  // namespace

extern "C" int larnd_mc_forward(const float* tracks_d, int64_t n, const larnd_columns_t* cols, const larnd_params_t* p,
                                const float* rnd_d, int32_t n_events, int32_t npix_capacity, void* workspace_d,
                                size_t workspace_bytes, int32_t* unique_pixels_d, float* wfs_d, int32_t* counts_d, void* stream) {
  int rc = mc_check(p, cols, rnd_d);
  if (rc) return rc;
  if ((!tracks_d && n > 0) || !unique_pixels_d || !wfs_d || !counts_d || npix_capacity < 1) { larnd_set_error("larnd_mc_forward: bad argument"); return LARND_E_ARG; }
  Workspace ws;
  if (!larnd_carve_workspace(workspace_d, workspace_bytes, n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y, &ws)) {
    larnd_set_error("workspace too small");
    return LARND_E_CAPACITY;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LARND_CUDA(cudaMemsetAsync(ws.bitmap, 0, ws.n_words * sizeof(uint32_t), st));
  LARND_CUDA(cudaMemsetAsync(counts_d, 0, 4 * sizeof(int32_t), st));
  LARND_CUDA(cudaMemsetAsync(wfs_d, 0, (size_t)npix_capacity * p->n_ticks * sizeof(float), st));
  if (n > 0) {
    size_t smem = (size_t)MCP_THREADS * (cols->ncols | 1) * sizeof(float);
    larnd_runs_cache_drop(ws.rec);
    k_mc_prepare<<<(unsigned)((n + MCP_THREADS - 1) / MCP_THREADS), MCP_THREADS, smem, st>>>(
        tracks_d, n, *cols, *p, rnd_d, ws.rec, ws.bitmap, ws.n_words, ws.pid_offset, counts_d);
    LARND_LAUNCH_CHECK("k_mc_prepare");
  }
  if ((rc = larnd_launch_scan(ws, *p, counts_d, st))) return rc;
  if ((rc = larnd_launch_unique(ws, *p, npix_capacity, /*extra=*/0, unique_pixels_d, counts_d, st))) return rc;
  if (n > 0) {
    McArgs A;
    A.rec = ws.rec; A.n = n;
    A.lk.bitmap = ws.bitmap; A.lk.wprefix = ws.wprefix; A.lk.n_words = ws.n_words; A.lk.pid_offset = ws.pid_offset;
    A.lk.n_unique = 0; A.lk.n_neg = 0; A.lk.npix = npix_capacity;
    A.counts = counts_d; A.nticks = p->n_ticks; A.wfs = wfs_d; A.g = nullptr; A.g_stride = 0; A.partials = nullptr;
    k_mc_accumulate<<<(unsigned)((n + MC_SEGS - 1) / MC_SEGS), MC_ACC_THREADS, 0, st>>>(A, *p);
    LARND_LAUNCH_CHECK("k_mc_accumulate");
  }
  return LARND_OK;
}

extern "C" int larnd_mc_backward(const float* tracks_d, int64_t n, const larnd_columns_t* cols, const larnd_params_t* p,
                                 const float* rnd_d, int32_t n_events, int32_t npix_capacity, void* workspace_d,
                                 size_t workspace_bytes, const int32_t* counts_d, const float* g_wfs_d, int64_t g_row_stride,
                                 float* grad_params_d, void* stream) {
  (void)tracks_d;
  int rc = mc_check(p, cols, rnd_d);
  if (rc) return rc;
  if (!g_wfs_d || !grad_params_d || !counts_d) { larnd_set_error("larnd_mc_backward: null argument"); return LARND_E_ARG; }
  if (n == 0) return LARND_OK;
  Workspace ws;
  if (!larnd_carve_workspace(workspace_d, workspace_bytes, n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y, &ws)) {
    larnd_set_error("workspace too small");
    return LARND_E_CAPACITY;
  }
  const int64_t blocks = (n + MC_WARPS - 1) / MC_WARPS;
  if (blocks > ws.n_chunks_max * (LARND_CHUNK / MC_WARPS)) { larnd_set_error("internal: partial buffer too small"); return LARND_E_CAPACITY; }
  cudaStream_t st = (cudaStream_t)stream;
  McArgs A;
  A.rec = ws.rec; A.n = n;
  A.lk.bitmap = ws.bitmap; A.lk.wprefix = ws.wprefix; A.lk.n_words = ws.n_words; A.lk.pid_offset = ws.pid_offset;
  A.lk.n_unique = 0; A.lk.n_neg = 0; A.lk.npix = npix_capacity;
  A.counts = counts_d;
  A.nticks = p->n_ticks; A.wfs = nullptr; A.g = g_wfs_d; A.g_stride = g_row_stride; A.partials = ws.partials;
  k_mc_backward<<<(unsigned)blocks, MC_WARPS * 32, 0, st>>>(A, *p);
  LARND_LAUNCH_CHECK("k_mc_backward");
  k_mc_reduce_partials<<<LARND_NPARAMS, 256, 0, st>>>(ws.partials, blocks, grad_params_d);
  LARND_LAUNCH_CHECK("k_mc_reduce_partials");
  return LARND_OK;
}

static int cur_args(CurArgs& A, const float* electrons_d, int64_t n, const larnd_current_columns_t* c, const float* pixels_coord_d,
                    const larnd_params_t* p) {
  if (!c || !p || n < 0 || (n > 0 && (!electrons_d || !pixels_coord_d))) { larnd_set_error("larnd_current_mc: bad argument"); return LARND_E_ARG; }
  if (p->n_tpc < 1 || p->n_tpc > LARND_MAX_TPC) { larnd_set_error("n_tpc unsupported"); return LARND_E_ARG; }
  int nt = (int)(5.0 / (double)p->t_sampling + 1e-4) + 1;
  if (nt != MC_NT) { larnd_set_error("current_mc is built for t_sampling = 0.1 us (51 ticks), got %d ticks", nt); return LARND_E_ARG; }
  A.electrons = electrons_d; A.n = n; A.ncols = c->ncols;
  A.cx = c->x; A.cy = c->y; A.cz = c->z; A.cld = c->long_diff; A.cq = c->n_electrons; A.cplane = c->pixel_plane;
  A.pixc = pixels_coord_d; A.t0_tick = nullptr; A.signals = nullptr; A.g_signals = nullptr; A.g_electrons = nullptr; A.g_pixc = nullptr;
  return LARND_OK;
}

extern "C" int larnd_current_mc(const float* electrons_d, int64_t n, const larnd_current_columns_t* c, const float* pixels_coord_d,
                                const larnd_params_t* p, int32_t* t0_tick_d, float* signals_d, void* stream) {
  CurArgs A;
  int rc = cur_args(A, electrons_d, n, c, pixels_coord_d, p);
  if (rc) return rc;
  if (n == 0) return LARND_OK;
  if (!t0_tick_d || !signals_d) { larnd_set_error("larnd_current_mc: null output"); return LARND_E_ARG; }
  A.t0_tick = t0_tick_d; A.signals = signals_d;
  k_current_mc<false><<<(unsigned)((n + MC_WARPS - 1) / MC_WARPS), MC_WARPS * 32, 0, (cudaStream_t)stream>>>(A, *p);
  LARND_LAUNCH_CHECK("k_current_mc");
  return LARND_OK;
}

extern "C" int larnd_current_mc_backward(const float* electrons_d, int64_t n, const larnd_current_columns_t* c,
                                         const float* pixels_coord_d, const larnd_params_t* p, const float* g_signals_d,
                                         float* g_electrons_d, float* g_pixels_coord_d, void* stream) {
  CurArgs A;
  int rc = cur_args(A, electrons_d, n, c, pixels_coord_d, p);
  if (rc) return rc;
  if (n == 0) return LARND_OK;
  if (!g_signals_d || !g_electrons_d || !g_pixels_coord_d) { larnd_set_error("larnd_current_mc_backward: null argument"); return LARND_E_ARG; }
  A.g_signals = g_signals_d; A.g_electrons = g_electrons_d; A.g_pixc = g_pixels_coord_d;
  cudaStream_t st = (cudaStream_t)stream;
  LARND_CUDA(cudaMemsetAsync(g_electrons_d, 0, (size_t)n * c->ncols * sizeof(float), st));
  k_current_mc<true><<<(unsigned)((n + MC_WARPS - 1) / MC_WARPS), MC_WARPS * 32, 0, st>>>(A, *p);
  LARND_LAUNCH_CHECK("k_current_mc<bwd>");
  return LARND_OK;
}

extern "C" int larnd_accumulate_parametrized(float* wfs_d, int32_t npix, int32_t n_ticks, const float* signals_d, int32_t n_signal_ticks,
                                             const int32_t* pix_id_d, const int32_t* start_ticks_d, int64_t n, void* stream) {
  if (!wfs_d || npix < 1 || n_ticks < 2 || n_signal_ticks < 1 || n < 0 || (n > 0 && (!signals_d || !pix_id_d || !start_ticks_d))) {
    larnd_set_error("larnd_accumulate_parametrized: bad argument");
    return LARND_E_ARG;
  }
  if (n == 0) return LARND_OK;
  const int64_t tot = n * n_signal_ticks;
  k_accumulate_parametrized<false><<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      wfs_d, nullptr, const_cast<float*>(signals_d), pix_id_d, start_ticks_d, n, n_signal_ticks, npix, n_ticks);
  LARND_LAUNCH_CHECK("k_accumulate_parametrized");
  return LARND_OK;
}

extern "C" int larnd_accumulate_parametrized_backward(const float* g_wfs_d, int32_t npix, int32_t n_ticks, float* g_signals_d,
                                                      int32_t n_signal_ticks, const int32_t* pix_id_d, const int32_t* start_ticks_d,
                                                      int64_t n, void* stream) {
  if (!g_wfs_d || npix < 1 || n_ticks < 2 || n_signal_ticks < 1 || n < 0 || (n > 0 && (!g_signals_d || !pix_id_d || !start_ticks_d))) {
    larnd_set_error("larnd_accumulate_parametrized_backward: bad argument");
    return LARND_E_ARG;
  }
  if (n == 0) return LARND_OK;
  const int64_t tot = n * n_signal_ticks;
  k_accumulate_parametrized<true><<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      nullptr, g_wfs_d, g_signals_d, pix_id_d, start_ticks_d, n, n_signal_ticks, npix, n_ticks);
  LARND_LAUNCH_CHECK("k_accumulate_parametrized<bwd>");
  return LARND_OK;
}
