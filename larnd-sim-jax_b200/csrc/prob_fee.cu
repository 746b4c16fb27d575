// Noise-averaged ("probabilistic") front end, forward — replaces the XLA lowering of
// get_adc_values_average_noise_vmap / _find_one_hit_step (reference fee_jax.py:334-461) and the helpers
// _soft_max / _soft_where / log_diff_ndtr (:12-53).
//
// Per pixel the reference runs a 20-path beam search over 10 hit steps; each step evaluates, for every
// (path, tick), three log-probabilities built from log_ndtr of shifted cumulative charges, reduces them over the
// paths with logsumexp, and keeps the 20 ticks with the largest selection probability as the next paths.
// Design: one CTA per pixel and step (the steps are separate launches because the reference stops ALL pixels as soon
// as no pixel has probability left — a global flag, here an atomicOr read by the next launch).  The running sum of
// the pixel's charge and its forward / reverse running maxima are computed once (k_prob_setup): lax.cummax of
// (q_sum - c_path) equals cummax(q_sum) - c_path bit for bit, so the per-path scans of the reference disappear.
// Each thread walks its ticks, keeps the 3 x 20 path terms of one tick in registers, reduces them (max-shifted
// logsumexp like jax.nn.logsumexp) and writes the two per-tick results; the top-20 selection is 20 block-wide
// arg-max rounds (value descending, lower tick first among equals, like lax.top_k).
#include "larnd_common.cuh"

namespace {

constexpr int PF_THREADS = 256;
constexpr int PF_MAXPATHS = 32;

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }  // logaddexp(x, 0)
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float soft_max_f(float x, float lo, float sharp) { return lo + softplus_f((x - lo) * sharp) / sharp; }
__device__ __forceinline__ float soft_where_f(float c, float tv, float fv, float sharp) {
  const float w = sigmoid_f(c * sharp);
  return w * tv + (1.0f - w) * fv;
}

__device__ __forceinline__ float ndtr_f(float x) {
  const float hs2 = 0.70710678f;
  const float w = x * hs2, z = fabsf(w);
  const float y = z < hs2 ? 1.0f + erff(w) : (w > 0.0f ? 2.0f - erfcf(z) : erfcf(z));
  return 0.5f * y;
}

// jax.scipy.special.log_ndtr, float32 segments (-10, 5), asymptotic series of order 3
__device__ __forceinline__ float log_ndtr_f(float x) {
  if (x > 5.0f) return -ndtr_f(-x);
  if (x > -10.0f) return logf(ndtr_f(x));
  const float x2 = x * x;
  const float log_scale = -0.5f * x2 - logf(-x) - 0.918938533f;
  const float even = 3.0f / (x2 * x2), odd = 1.0f / x2 + 15.0f / (x2 * x2 * x2);
  return log_scale + logf(1.0f + even - odd);
}

__device__ __forceinline__ float log_diff_ndtr_f(float a, float b) {
  const float la = log_ndtr_f(a), lb = log_ndtr_f(b);
  const float safe_diff = a > b ? lb - la : -1.0f;
  const float neg_expm1 = -expm1f(safe_diff);
  const float neg_safe = soft_max_f(neg_expm1, 1e-30f, 1e10f);
  const float log_term_safe = soft_max_f(logf(neg_safe), -100.0f, 10.0f);
  return soft_where_f(a - b, la + log_term_safe, -1000.0f, 1000.0f);
}

// q_sum = left-to-right float32 running sum of wfs * t_sampling; cmf / cmr = forward / reverse running maxima
__global__ void k_prob_setup(const float* __restrict__ wfs, int64_t stride, int npix, int nt, float t_sampling,
                             float* __restrict__ qsum, float* __restrict__ cmf, float* __restrict__ cmr, float* __restrict__ charges,
                             float* __restrict__ lps, int npaths, int* __restrict__ flags, int nsteps) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x <= nsteps) flags[threadIdx.x] = threadIdx.x == 0 ? 1 : 0;
  if (pix >= npix) return;
  const float* w = wfs + (int64_t)pix * stride;
  float* q = qsum + (int64_t)pix * nt;
  float* f = cmf + (int64_t)pix * nt;
  float* r = cmr + (int64_t)pix * nt;
  float acc = 0.0f, mx = -INFINITY;
  for (int t = 0; t < nt; ++t) {
    acc = __fadd_rn(acc, __fmul_rn(w[t], t_sampling));
    q[t] = acc;
    mx = fmaxf(mx, acc);
    f[t] = mx;
  }
  mx = -INFINITY;
  for (int t = nt - 1; t >= 0; --t) {
    mx = fmaxf(mx, q[t]);
    r[t] = mx;
  }
  for (int k = 0; k < npaths; ++k) {
    charges[(int64_t)pix * npaths + k] = 0.0f;
    lps[(int64_t)pix * npaths + k] = k == 0 ? 0.0f : -1000.0f;
  }
}

template <int NP>
__global__ void __launch_bounds__(PF_THREADS)
k_prob_step(const float* __restrict__ qsum, const float* __restrict__ cmf, const float* __restrict__ cmr, int npix, int nt,
            float* __restrict__ charges, float* __restrict__ lps, int* __restrict__ flags, int step, int nsteps, float zscale, float thr,
            int interval, float log_stop, float* __restrict__ out_lp, float* __restrict__ out_q, int* __restrict__ out_top,
            float* __restrict__ out_state) {
  extern __shared__ float smf[];
  float* s_q = smf;              // [nt]
  float* s_f = s_q + nt;         // [nt]
  float* s_r = s_f + nt;         // [nt]
  float* s_tot = s_r + nt;       // [nt]
  float* s_sel = s_tot + nt;     // [nt]
  __shared__ float s_c[PF_MAXPATHS], s_lp[PF_MAXPATHS];
  __shared__ float s_rv[PF_THREADS / 32];
  __shared__ int s_ri[PF_THREADS / 32];
  __shared__ int s_top[PF_MAXPATHS];
  const int pix = blockIdx.x;
  const int ntm = nt - 1;
  float* olp = out_lp + ((int64_t)pix * nsteps + step) * ntm;
  float* oq = out_q + ((int64_t)pix * nsteps + step) * ntm;
  if (flags[step] == 0) {  // globally stopped (fee_jax.py:431-437): cheap pass-through
    for (int t = threadIdx.x; t < ntm; t += PF_THREADS) { olp[t] = -1000.0f; oq[t] = 0.0f; }
    return;
  }
  for (int t = threadIdx.x; t < nt; t += PF_THREADS) {
    s_q[t] = qsum[(int64_t)pix * nt + t];
    s_f[t] = cmf[(int64_t)pix * nt + t];
    s_r[t] = cmr[(int64_t)pix * nt + t];
  }
  if (threadIdx.x < NP) {
    s_c[threadIdx.x] = charges[(int64_t)pix * NP + threadIdx.x];
    s_lp[threadIdx.x] = lps[(int64_t)pix * NP + threadIdx.x];
    if (out_state) {  // path state BEFORE this step (needed by the backward pass)
      out_state[((int64_t)pix * nsteps + step) * 2 * NP + threadIdx.x] = s_c[threadIdx.x];
      out_state[((int64_t)pix * nsteps + step) * 2 * NP + NP + threadIdx.x] = s_lp[threadIdx.x];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ntm; t += PF_THREADS) {
    const int sh = min(t + interval + 1, nt - 1);
    const int shn = min(sh + 1, nt - 1), fend = min(sh + interval + 1, nt - 1);
    const float q_t = s_q[t], q_t1 = s_q[t + 1], q_sh = s_q[sh], q_shn = s_q[shn];
    const float f_t = s_f[t], f_t1 = s_f[t + 1], r_fe = s_r[fend];
    const float esp = q_sh + thr - 0.5f * (q_t1 + q_t);
    float va[NP], vb[NP], vc[NP];
    float ma = -INFINITY, mb = -INFINITY, mc = -INFINITY;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const float c = s_c[p], lp = s_lp[p];
      const float lg = log_diff_ndtr_f(((f_t1 - c) - thr) * zscale, ((f_t - c) - thr) * zscale);
      float le = log_diff_ndtr_f(((q_sh - c) - thr) * zscale, ((q_t - c) - thr) * zscale);
      le = fminf(le, lg);
      le = soft_max_f(le, -1000.0f, 1.0f);
      le = soft_where_f(esp - thr, le, -1000.0f, 10.0f);
      const float lf = log_ndtr_f((((r_fe - c) - (q_shn - c)) - thr) * zscale);
      va[p] = le + lp;
      vb[p] = lg + lp;
      vc[p] = (lg + lf) + lp;
      ma = fmaxf(ma, va[p]); mb = fmaxf(mb, vb[p]); mc = fmaxf(mc, vc[p]);
    }
    float sa = 0.0f, sb = 0.0f, sc = 0.0f;
#pragma unroll
    for (int p = 0; p < NP; ++p) { sa += expf(va[p] - ma); sb += expf(vb[p] - mb); sc += expf(vc[p] - mc); }
    olp[t] = logf(sa) + ma;
    oq[t] = esp;
    s_tot[t] = logf(sb) + mb;
    s_sel[t] = logf(sc) + mc;
  }
  __syncthreads();
  // top-NP ticks of the selection probability: value descending, lower tick first among equals (lax.top_k)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < NP; ++k) {
    float bv = -INFINITY;
    int bi = INT32_MAX;
    for (int t = threadIdx.x; t < ntm; t += PF_THREADS) {
      const float v = s_sel[t];
      if (v > bv || (v == bv && t < bi)) { bv = v; bi = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_rv[wid] = bv; s_ri[wid] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = s_rv[0];
      int i = s_ri[0];
      for (int w = 1; w < PF_THREADS / 32; ++w)
        if (s_rv[w] > v || (s_rv[w] == v && s_ri[w] < i)) { v = s_rv[w]; i = s_ri[w]; }
      s_top[k] = i;
      s_sel[i] = -INFINITY;  // NaN-free inputs: a selected tick never wins again
    }
    __syncthreads();
  }
  if (threadIdx.x < NP) {
    const int tk = s_top[threadIdx.x];
    const int sh = min(tk + interval + 1, nt - 1);
    lps[(int64_t)pix * NP + threadIdx.x] = s_tot[tk];
    charges[(int64_t)pix * NP + threadIdx.x] = s_q[min(sh + 1, nt - 1)];
    if (out_top) out_top[((int64_t)pix * nsteps + step) * NP + threadIdx.x] = tk;
    s_lp[threadIdx.x] = s_tot[tk];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int p = 0; p < NP; ++p) m = fmaxf(m, s_lp[p]);
    float s = 0.0f;
    for (int p = 0; p < NP; ++p) s += expf(s_lp[p] - m);
    if (logf(s) + m > log_stop) atomicOr(flags + step + 1, 1);
  }
}

}  // namespace

extern "C" size_t larnd_prob_fee_scratch_bytes(int32_t npix, int32_t n_ticks, int32_t n_paths, int32_t n_steps) {
  if (npix < 0 || n_ticks < 2 || n_paths < 1 || n_steps < 1) return 0;
  size_t b = 0;
  b += align_up((size_t)npix * n_ticks * sizeof(float), 256) * 3;   // q_sum, forward / reverse running maxima
  b += align_up((size_t)npix * n_paths * sizeof(float), 256) * 2;   // path charges, path log-probabilities
  b += align_up((size_t)(n_steps + 1) * sizeof(int), 256);          // global "still active" flags per step
  return b;
}

extern "C" int larnd_prob_fee_forward(const float* wfs_d, int64_t wfs_row_stride, int32_t npix, int32_t n_ticks,
                                      const larnd_params_t* p, int32_t n_paths, float stop_threshold, float* log_prob_d,
                                      float* charge_d, int32_t* top_ticks_d, float* state_d, int32_t* flags_d, void* scratch_d,
                                      size_t scratch_bytes, void* stream) {
  if (!p || (!wfs_d && npix > 0) || !log_prob_d || !charge_d || !scratch_d || npix < 0 || n_ticks < 2) {
    larnd_set_error("larnd_prob_fee_forward: bad argument");
    return LARND_E_ARG;
  }
  if (n_paths != 20) { larnd_set_error("larnd_prob_fee_forward: fee_paths_scaling must be 20 (got %d)", n_paths); return LARND_E_ARG; }
  const int nsteps = p->max_adc_values;
  if (scratch_bytes < larnd_prob_fee_scratch_bytes(npix, n_ticks, n_paths, nsteps)) { larnd_set_error("prob_fee scratch too small"); return LARND_E_CAPACITY; }
  if (p->reset_noise_charge <= 0.0f) { larnd_set_error("larnd_prob_fee_forward: RESET_NOISE_CHARGE must be > 0 (it is the noise sigma)"); return LARND_E_ARG; }
  if (npix == 0) return LARND_OK;
  cudaStream_t st = (cudaStream_t)stream;
  char* b = reinterpret_cast<char*>(scratch_d);
  float* qsum = reinterpret_cast<float*>(b); b += align_up((size_t)npix * n_ticks * sizeof(float), 256);
  float* cmf = reinterpret_cast<float*>(b); b += align_up((size_t)npix * n_ticks * sizeof(float), 256);
  float* cmr = reinterpret_cast<float*>(b); b += align_up((size_t)npix * n_ticks * sizeof(float), 256);
  float* charges = reinterpret_cast<float*>(b); b += align_up((size_t)npix * n_paths * sizeof(float), 256);
  float* lps = reinterpret_cast<float*>(b); b += align_up((size_t)npix * n_paths * sizeof(float), 256);
  int* flags = reinterpret_cast<int*>(b);
  k_prob_setup<<<(npix + 63) / 64, 64, 0, st>>>(wfs_d, wfs_row_stride, npix, n_ticks, p->t_sampling, qsum, cmf, cmr, charges, lps,
                                                n_paths, flags, nsteps);
  LARND_LAUNCH_CHECK("k_prob_setup");
  const size_t smem = (size_t)5 * n_ticks * sizeof(float);
  if (smem > 200 * 1024) { larnd_set_error("larnd_prob_fee_forward: too many ticks for shared memory"); return LARND_E_ARG; }
  LARND_CUDA(cudaFuncSetAttribute(k_prob_step<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float zscale = 1.0f / p->reset_noise_charge;
  for (int s = 0; s < nsteps; ++s) {
    k_prob_step<20><<<npix, PF_THREADS, smem, st>>>(qsum, cmf, cmr, npix, n_ticks, charges, lps, flags, s, nsteps, zscale,
                                                     p->discrimination_threshold, p->hold_interval, logf(stop_threshold), log_prob_d,
                                                     charge_d, top_ticks_d, state_d);
    LARND_LAUNCH_CHECK("k_prob_step");
  }
  if (flags_d) LARND_CUDA(cudaMemcpyAsync(flags_d, flags, (size_t)(nsteps + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return LARND_OK;
}

// =============================================================================================== backward
// VJP of the beam search w.r.t. the waveforms.  The discrete choices (beam ticks, global stop flags) are the forward's
// (lax.stop_gradient in the reference, fee_jax.py:380).  Steps are walked in reverse with the adjoint of the path state
// (d charges, d log-probabilities) as carry; every step recomputes its log-probability terms and their partials.
namespace {

__device__ __forceinline__ float log_ndtr_prime(float x, float l) { return 0.39894228f * expf(-0.5f * x * x - l); }  // pdf / cdf

// value and partials of log_diff_ndtr(a, b)
__device__ __forceinline__ float ldn_grad(float a, float b, float& da, float& db) {
  const float la = log_ndtr_f(a), lb = log_ndtr_f(b);
  const bool gt = a > b;
  const float sd = gt ? lb - la : -1.0f;
  const float esd = expf(sd);
  const float ne = -expm1f(sd);
  const float arg = (ne - 1e-30f) * 1e10f;
  const float ns = 1e-30f + softplus_f(arg) / 1e10f;
  const float lt = logf(ns);
  const float lts = soft_max_f(lt, -100.0f, 10.0f);
  const float lpv = la + lts;
  const float w = sigmoid_f(1000.0f * (a - b));
  const float k = gt ? sigmoid_f((lt + 100.0f) * 10.0f) * (1.0f / ns) * sigmoid_f(arg) * (-esd) : 0.0f;  // d lts / d sd
  const float dab = 1000.0f * w * (1.0f - w) * (lpv + 1000.0f);
  da = w * (1.0f - k) * log_ndtr_prime(a, la) + dab;
  db = w * k * log_ndtr_prime(b, lb) - dab;
  return w * lpv + (1.0f - w) * (-1000.0f);
}

__global__ void k_prob_bwd_setup(const float* __restrict__ wfs, int64_t stride, int npix, int nt, float t_sampling,
                                 float* __restrict__ qsum, float* __restrict__ cmf, int* __restrict__ amax, float* __restrict__ dq,
                                 float* __restrict__ dcmf, float* __restrict__ carry, int npaths) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const float* w = wfs + (int64_t)pix * stride;
  float acc = 0.0f, mx = -INFINITY;
  int im = 0;
  for (int t = 0; t < nt; ++t) {
    acc = __fadd_rn(acc, __fmul_rn(w[t], t_sampling));
    qsum[(int64_t)pix * nt + t] = acc;
    if (acc > mx) { mx = acc; im = t; }  // first index attaining the running maximum
    cmf[(int64_t)pix * nt + t] = mx;
    amax[(int64_t)pix * nt + t] = im;
    dq[(int64_t)pix * nt + t] = 0.0f;
    dcmf[(int64_t)pix * nt + t] = 0.0f;
  }
  for (int k = 0; k < 2 * npaths; ++k) carry[(int64_t)pix * 2 * npaths + k] = 0.0f;
}

template <int NP>
__global__ void __launch_bounds__(PF_THREADS)
k_prob_bwd_step(const float* __restrict__ qsum, const float* __restrict__ cmf, int npix, int nt, const float* __restrict__ state,
                const int* __restrict__ top, const int* __restrict__ flags, int step, int nsteps, float zscale, float thr, int interval,
                const float* __restrict__ lp_fwd, const float* __restrict__ g_lp, const float* __restrict__ g_q,
                float* __restrict__ dq, float* __restrict__ dcmf, float* __restrict__ carry) {
  if (flags[step] == 0) return;  // inactive step: constants out, carry passes through
  extern __shared__ float smf[];
  float* s_q = smf;            // [nt]
  float* s_f = s_q + nt;       // [nt]
  float* s_gt = s_f + nt;      // [nt] adjoint of log_total_dist_tick (non-zero at the beam ticks only)
  float* s_dqt = s_gt + nt;    // [nt] contributions to d q_sum[t]
  float* s_dqs = s_dqt + nt;   // [nt] contributions destined to d q_sum[shifted(t)]
  float* s_dft = s_dqs + nt;   // [nt] contributions to d cmf[t]
  float* s_dft1 = s_dft + nt;  // [nt] contributions destined to d cmf[t + 1]
  float* s_dqt1 = s_dft1 + nt; // [nt] contributions destined to d q_sum[t + 1]
  __shared__ float s_c[PF_MAXPATHS], s_lp[PF_MAXPATHS], s_dc1[PF_MAXPATHS], s_dlp1[PF_MAXPATHS];
  __shared__ int s_top[PF_MAXPATHS];
  __shared__ float s_red[PF_THREADS / 32][2 * PF_MAXPATHS];
  const int pix = blockIdx.x, ntm = nt - 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < nt; t += PF_THREADS) {
    s_q[t] = qsum[(int64_t)pix * nt + t];
    s_f[t] = cmf[(int64_t)pix * nt + t];
    s_gt[t] = 0.0f; s_dqt[t] = 0.0f; s_dqs[t] = 0.0f; s_dft[t] = 0.0f; s_dft1[t] = 0.0f; s_dqt1[t] = 0.0f;
  }
  if (threadIdx.x < NP) {
    s_c[threadIdx.x] = state[((int64_t)pix * nsteps + step) * 2 * NP + threadIdx.x];
    s_lp[threadIdx.x] = state[((int64_t)pix * nsteps + step) * 2 * NP + NP + threadIdx.x];
    s_top[threadIdx.x] = top[((int64_t)pix * nsteps + step) * NP + threadIdx.x];
    s_dc1[threadIdx.x] = carry[(int64_t)pix * 2 * NP + threadIdx.x];
    s_dlp1[threadIdx.x] = carry[(int64_t)pix * 2 * NP + NP + threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x < NP) {
    const int tk = s_top[threadIdx.x];
    s_gt[tk] = s_dlp1[threadIdx.x];                                   // new_log_prob[k] = log_total_dist_tick[top_k]  (distinct ticks)
    const int bn = min(min(tk + interval + 1, nt - 1) + 1, nt - 1);   // charges_new[k] = q_sum[best_next]
    atomicAdd(dq + (int64_t)pix * nt + bn, s_dc1[threadIdx.x]);
  }
  __syncthreads();
  float dc[NP], dlp[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { dc[p] = 0.0f; dlp[p] = 0.0f; }
  const float* glp = g_lp + ((int64_t)pix * nsteps + step) * ntm;
  const float* gq = g_q + ((int64_t)pix * nsteps + step) * ntm;
  const float* lpf = lp_fwd + ((int64_t)pix * nsteps + step) * ntm;
  for (int t = threadIdx.x; t < ntm; t += PF_THREADS) {
    const int sh = min(t + interval + 1, nt - 1);
    const float q_t = s_q[t], q_t1 = s_q[t + 1], q_sh = s_q[sh], f_t = s_f[t], f_t1 = s_f[t + 1];
    const float esp = q_sh + thr - 0.5f * (q_t1 + q_t);
    const float w2 = sigmoid_f(10.0f * (esp - thr));
    const float g_hit = glp[t], lse_hit = lpf[t], g_tot = s_gt[t];
    float lse_tot = 0.0f;
    if (g_tot != 0.0f) {  // beam ticks only: logsumexp over the paths of (log_guess + previous log-probability)
      float vb[NP], mb = -INFINITY;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const float c = s_c[p];
        vb[p] = log_diff_ndtr_f(((f_t1 - c) - thr) * zscale, ((f_t - c) - thr) * zscale) + s_lp[p];
        mb = fmaxf(mb, vb[p]);
      }
      float sb = 0.0f;
#pragma unroll
      for (int p = 0; p < NP; ++p) sb += expf(vb[p] - mb);
      lse_tot = logf(sb) + mb;
    }
    float d_esp = 0.0f, a_qs = 0.0f, a_qt = 0.0f, a_ft1 = 0.0f, a_ft = 0.0f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const float c = s_c[p], lp = s_lp[p];
      float dga, dgb, dea, deb;
      const float vg = ldn_grad(((f_t1 - c) - thr) * zscale, ((f_t - c) - thr) * zscale, dga, dgb);
      const float ve = ldn_grad(((q_sh - c) - thr) * zscale, ((q_t - c) - thr) * zscale, dea, deb);
      const float m = fminf(ve, vg);
      const float dme = ve < vg ? 1.0f : (ve == vg ? 0.5f : 0.0f), dmg = 1.0f - dme;
      const float sm1 = soft_max_f(m, -1000.0f, 1.0f);
      const float le = w2 * sm1 + (1.0f - w2) * (-1000.0f);
      const float G_le = expf((le + lp) - lse_hit) * g_hit;
      const float G_vt = g_tot != 0.0f ? expf((vg + lp) - lse_tot) * g_tot : 0.0f;
      dlp[p] += G_le + G_vt;
      d_esp += G_le * 10.0f * w2 * (1.0f - w2) * (sm1 + 1000.0f);
      const float G_m = G_le * w2 * sigmoid_f(m + 1000.0f);
      const float G_ve = G_m * dme, G_vg = G_m * dmg + G_vt;
      const float xa_e = G_ve * dea * zscale, xb_e = G_ve * deb * zscale, xa_g = G_vg * dga * zscale, xb_g = G_vg * dgb * zscale;
      a_qs += xa_e; a_qt += xb_e; a_ft1 += xa_g; a_ft += xb_g;
      dc[p] -= (xa_e + xb_e) + (xa_g + xb_g);
    }
    d_esp += gq[t];
    s_dqs[t] = a_qs + d_esp;
    s_dqt[t] = a_qt - 0.5f * d_esp;
    s_dqt1[t] = -0.5f * d_esp;
    s_dft[t] = a_ft;
    s_dft1[t] = a_ft1;
  }
  __syncthreads();
  // scatter the per-tick contributions to their destinations (each destination index is owned by one thread)
  for (int u = threadIdx.x; u < nt; u += PF_THREADS) {
    float a = (u < ntm ? s_dqt[u] : 0.0f) + (u >= 1 ? s_dqt1[u - 1] : 0.0f);
    float b = (u < ntm ? s_dft[u] : 0.0f) + (u >= 1 ? s_dft1[u - 1] : 0.0f);
    if (u < nt - 1) {
      const int t = u - interval - 1;               // shifted(t) = u, unclipped
      if (t >= 0 && t < ntm) a += s_dqs[t];
    } else {
      for (int t = max(0, nt - 2 - interval); t < ntm; ++t) a += s_dqs[t];  // every t whose shifted tick is clipped to nt-1
    }
    dq[(int64_t)pix * nt + u] += a;
    dcmf[(int64_t)pix * nt + u] += b;
  }
  // block reduction of the path adjoints -> carry for the previous step
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float v0 = dc[p], v1 = dlp[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
    if (lane == 0) { s_red[wid][p] = v0; s_red[wid][NP + p] = v1; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * NP) {
    float v = 0.0f;
    for (int w = 0; w < PF_THREADS / 32; ++w) v += s_red[w][threadIdx.x];
    carry[(int64_t)pix * 2 * NP + threadIdx.x] = v;
  }
}

// d q_sum[argmax] += d cmf, then d wfs = t_sampling * reverse cumsum of d q_sum
__global__ void k_prob_bwd_finish(const float* __restrict__ dq, const float* __restrict__ dcmf, const int* __restrict__ amax, int npix,
                                  int nt, float t_sampling, float* __restrict__ dq_tmp, float* __restrict__ g_wfs, int64_t g_stride) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  float* tmp = dq_tmp + (int64_t)pix * nt;
  for (int t = 0; t < nt; ++t) tmp[t] = dq[(int64_t)pix * nt + t];
  for (int t = 0; t < nt; ++t) tmp[amax[(int64_t)pix * nt + t]] += dcmf[(int64_t)pix * nt + t];
  float acc = 0.0f;
  for (int t = nt - 1; t >= 0; --t) {
    acc += tmp[t];
    g_wfs[(int64_t)pix * g_stride + t] = acc * t_sampling;
  }
}

}  // namespace

extern "C" size_t larnd_prob_fee_bwd_scratch_bytes(int32_t npix, int32_t n_ticks, int32_t n_paths) {
  if (npix < 0 || n_ticks < 2 || n_paths < 1) return 0;
  return align_up((size_t)npix * n_ticks * 4, 256) * 6 + align_up((size_t)npix * 2 * n_paths * 4, 256);
}

extern "C" int larnd_prob_fee_backward(const float* wfs_d, int64_t wfs_row_stride, int32_t npix, int32_t n_ticks,
                                       const larnd_params_t* p, int32_t n_paths, const float* g_log_prob_d, const float* g_charge_d,
                                       const float* log_prob_d, const float* state_d, const int32_t* top_ticks_d, const int32_t* flags_d,
                                       float* g_wfs_d, int64_t g_row_stride, void* scratch_d, size_t scratch_bytes, void* stream) {
  if (!p || !g_log_prob_d || !g_charge_d || !log_prob_d || !state_d || !top_ticks_d || !flags_d || !g_wfs_d || !scratch_d || npix < 0 ||
      n_ticks < 2) {
    larnd_set_error("larnd_prob_fee_backward: bad argument");
    return LARND_E_ARG;
  }
  if (n_paths != 20) { larnd_set_error("larnd_prob_fee_backward: fee_paths_scaling must be 20"); return LARND_E_ARG; }
  if (scratch_bytes < larnd_prob_fee_bwd_scratch_bytes(npix, n_ticks, n_paths)) { larnd_set_error("prob_fee backward scratch too small"); return LARND_E_CAPACITY; }
  if (npix == 0) return LARND_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int nsteps = p->max_adc_values;
  char* b = reinterpret_cast<char*>(scratch_d);
  const size_t row = align_up((size_t)npix * n_ticks * 4, 256);
  float* qsum = reinterpret_cast<float*>(b); b += row;
  float* cmf = reinterpret_cast<float*>(b); b += row;
  int* amax = reinterpret_cast<int*>(b); b += row;
  float* dq = reinterpret_cast<float*>(b); b += row;
  float* dcmf = reinterpret_cast<float*>(b); b += row;
  float* tmp = reinterpret_cast<float*>(b); b += row;
  float* carry = reinterpret_cast<float*>(b);
  k_prob_bwd_setup<<<(npix + 63) / 64, 64, 0, st>>>(wfs_d, wfs_row_stride, npix, n_ticks, p->t_sampling, qsum, cmf, amax, dq, dcmf, carry, n_paths);
  LARND_LAUNCH_CHECK("k_prob_bwd_setup");
  const size_t smem = (size_t)8 * n_ticks * sizeof(float);
  if (smem > 200 * 1024) { larnd_set_error("larnd_prob_fee_backward: too many ticks for shared memory"); return LARND_E_ARG; }
  LARND_CUDA(cudaFuncSetAttribute(k_prob_bwd_step<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float zscale = 1.0f / p->reset_noise_charge;
  for (int s = nsteps - 1; s >= 0; --s) {
    k_prob_bwd_step<20><<<npix, PF_THREADS, smem, st>>>(qsum, cmf, npix, n_ticks, state_d, top_ticks_d, flags_d, s, nsteps, zscale,
                                                         p->discrimination_threshold, p->hold_interval, log_prob_d, g_log_prob_d,
                                                         g_charge_d, dq, dcmf, carry);
    LARND_LAUNCH_CHECK("k_prob_bwd_step");
  }
  k_prob_bwd_finish<<<(npix + 63) / 64, 64, 0, st>>>(dq, dcmf, amax, npix, n_ticks, p->t_sampling, tmp, g_wfs_d, g_row_stride);
  LARND_LAUNCH_CHECK("k_prob_bwd_finish");
  return LARND_OK;
}
