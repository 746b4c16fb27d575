// K5 front-end electronics: fused discriminator / ADC / digitiser / coordinates / hit compaction, sm_100a.
//
// Replaces the XLA lowering of simulate_stochastic (reference sim_jax.py:738-769):
//   get_adc_values (fee_jax.py:170-279: cumsum, 10x {first threshold crossing, hold-delay sample, reject,
//   subtract, clamp}), digitize (fee_jax.py:57-71), id2pixel / get_pixel_coordinates / get_hit_z
//   (detsim_jax.py:265-319) and parse_output (sim_jax.py:620-647).
//
// One warp owns one pixel row (n_ticks-1 samples, 8 KB in shared memory).  The running sum is formed
// strictly left to right in float32 by lane 0 (bit-identical to a sequential cumsum, so threshold
// crossings match the oracle exactly); the ten discriminator passes are warp-cooperative ballot scans
// over the row in shared memory.  HBM traffic = one read of the waveform + O(10) words per row.
#include "larnd_common.cuh"

namespace {

#ifndef LARND_FEE_WARPS
#define LARND_FEE_WARPS 4
#endif
#define FEE_ROW_STRIDE(nt) ((((nt) + 3) & ~3) + 4)  // floats per warp: the row, rounded up to whole vectors, + the alignment shift

struct FeeArgs {
  const float* wfs;
  int64_t stride;
  const int32_t* unique_pixels;
  int npix;
  int ntw;  // n_ticks - 1
  const float* noise;
  float* adc; float* ticks; float* pixel_z; float* pixel_x; float* pixel_y; int32_t* event; float* saved;
  int32_t* row_counts;
  int bulk;  // rows are loaded with cp.async.bulk (layout contract checked by larnd_fee_forward)
  int clear; // LARND_FEE_CLEAR_WFS: every non-zero sample of the waveform buffer is zeroed once it has been read (bulk layout only)
};

__device__ __forceinline__ int floordiv_pos(int a, int b) { return floordiv_i(a, b); }

// Sequential float32 running sum (fee_jax.py:193) of one row's window [t_first, t_last], strictly left to right, in place.
// Scalar head (t < head), whole vectors of the body (the window bounds sit on vector boundaries there; samples of a vector
// outside the window are zeros), scalar tail: one 128-bit load and store per four dependent adds.  Returns the row total.
__device__ __forceinline__ float fee_serial_sum(float* c, int head, int nvec, int t_first, int t_last) {
  float4* c4 = reinterpret_cast<float4*>(c + head);
  float acc = 0.0f;
  int t = t_first;
  for (; t <= t_last && t < head; ++t) {
    acc = __fadd_rn(acc, c[t]);
    c[t] = acc;
  }
  if (t <= t_last) {
    const int vend = min(nvec - 1, (t_last - head) >> 2);
#pragma unroll 4
    for (int v = (t - head) >> 2; v <= vend; ++v) {
      float4 q = c4[v];
      acc = __fadd_rn(acc, q.x); q.x = acc;
      acc = __fadd_rn(acc, q.y); q.y = acc;
      acc = __fadd_rn(acc, q.z); q.z = acc;
      acc = __fadd_rn(acc, q.w); q.w = acc;
      c4[v] = q;
    }
    for (int t2 = max(t, head + 4 * nvec); t2 <= t_last; ++t2) {
      acc = __fadd_rn(acc, c[t2]);
      c[t2] = acc;
    }
  }
  return acc;
}

// Three phases per CTA of FEE_WARPS rows:
//   1. warp <-> row: the row is fetched into shared memory (TMA), the bounds of its non-zero samples are found with an
//      integer OR over 128-bit vectors (no arithmetic on the ~1850 zero samples of a track's pixel) and only the window is
//      scaled by t_sampling;
//   2. lane <-> row: lanes 0 .. FEE_WARPS-1 of warp 0 form the sequential running sums of ALL rows of the CTA at once — the
//      dependent-add chain is the one part of the algorithm that has no parallelism inside a row (it was 21 % of the
//      kernel's instructions with one useful lane of 32 per row);
//   3. warp <-> row: the ten discriminator passes as ballot scans over the window.
template <int FEE_WARPS>
__global__ void __launch_bounds__(FEE_WARPS * 32)
k_fee_forward(const __grid_constant__ FeeArgs F, const __grid_constant__ larnd_params_t p) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * FEE_WARPS + wid;
  const bool live = row < F.npix;
  const int Nt = F.ntw;
  const int rstride = FEE_ROW_STRIDE(Nt);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + (size_t)FEE_WARPS * rstride);
  int* const s_first = reinterpret_cast<int*>(bars + FEE_WARPS);
  int* const s_last = s_first + FEE_WARPS;
  float* const s_total = reinterpret_cast<float*>(s_last + FEE_WARPS);
  const float* w = F.wfs + (int64_t)(live ? row : 0) * F.stride;
  // The row base is only 4-byte aligned (simulate_wfs hands out wfs[:, 1:] of a 2001-float row): `head` floats up to the
  // next 16-byte boundary and `tail` floats after the last whole vector are handled by six lanes, the body as 128-bit
  // vectors.  The copy in shared memory is shifted by `shift` floats so that the vectors are aligned there as well.
  const int head = min((int)(((16u - (unsigned)((uintptr_t)w & 15u)) & 15u) >> 2), Nt);
  const int nvec = (Nt - head) >> 2, tail = Nt - head - 4 * nvec;
  const int shift = (4 - head) & 3;
  float* c = smem + (size_t)wid * rstride + shift;
  int t_first = Nt, t_last = -1;  // bounds of the non-zero samples (vector granularity: a superset is enough, zeros inside
                                  // the window do not change any running sum)
  const float4* w4 = reinterpret_cast<const float4*>(w + head);
  float4* c4 = reinterpret_cast<float4*>(c + head);
  if (live) {
    if (F.bulk) {
      // Row load by the TMA engine: ONE cp.async.bulk per row (the whole 8 KB in flight per warp instead of four 512-byte
      // warp loads at a time), completion on a per-warp mbarrier.  Layout contract (checked on the host): rows start `shift`
      // (0 or 1) floats past a 16-byte boundary and the stride is a multiple of four floats — simulate_wfs' view [:, 1:] of
      // the padded waveform buffer — so the copy starts at the aligned address below the row (its garbage column) and lands at
      // the start of this warp's buffer; c = buffer + shift is the same shifted copy the vector path builds.
      const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bars + wid);
      const unsigned dst_s = (unsigned)__cvta_generic_to_shared(c - shift);
      const unsigned bytes = (unsigned)(((shift + Nt + 3) >> 2) << 4);
      if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_s), "l"(w - shift), "r"(bytes), "r"(bar_s) : "memory");
      }
      __syncwarp();
      unsigned done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_s) : "memory");
      }
    } else {
      if (lane < head + tail) {
        const int t = lane < head ? lane : head + 4 * nvec + (lane - head);
        c[t] = __ldg(w + t);
      }
      for (int v0 = 0; v0 < nvec; v0 += 128) {
        float4 x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // four 512-byte warp loads in flight
          const int v = v0 + 32 * k + lane;
          x[k] = v < nvec ? __ldg(w4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int v = v0 + 32 * k + lane;
          if (v < nvec) c4[v] = x[k];
        }
      }
      __syncwarp();
    }
    // bounds of the non-zero samples: bitwise OR of whole vectors (-0.0 counts as non-zero: a superset), two vectors per trip
    if (lane < head + tail) {
      const int t = lane < head ? lane : head + 4 * nvec + (lane - head);
      if (__float_as_uint(c[t]) != 0u) { t_first = t; t_last = t; }
    }
    {
      const uint4* u4 = reinterpret_cast<const uint4*>(c4);
      int vf = 0x3fffffff, vl = -1;
      for (int v = lane; v < nvec; v += 64) {
        const uint4 a = u4[v];
        const bool has_b = v + 32 < nvec;
        const uint4 b = has_b ? u4[v + 32] : make_uint4(0u, 0u, 0u, 0u);
        const unsigned oa = a.x | a.y | a.z | a.w, ob = b.x | b.y | b.z | b.w;
        if (oa) { vf = min(vf, v); vl = v; }
        if (ob) { vf = min(vf, v + 32); vl = v + 32; }
      }
      if (vl >= 0) {
        t_first = min(t_first, head + 4 * vf);
        t_last = max(t_last, head + 4 * vl + 3);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t_first = min(t_first, __shfl_xor_sync(0xffffffffu, t_first, o));
      t_last = max(t_last, __shfl_xor_sync(0xffffffffu, t_last, o));
    }
    // Self-cleaning waveform buffer: the row now lives in shared memory, and everything non-zero in global memory lies inside
    // [t_first, t_last] (plus, possibly, the garbage column in front of the row, which came along with the TMA copy): zero
    // exactly that — ~150 floats per row instead of the 2 GB memset the next accumulate call would otherwise need.
    if (F.clear) {
      float* wz = const_cast<float*>(w);
      if (lane == 0 && shift > 0) {
        bool nz0 = false;
        for (int k = 1; k <= shift; ++k) nz0 |= __float_as_uint(c[-k]) != 0u;
        if (nz0) for (int k = 1; k <= shift; ++k) wz[-k] = 0.0f;
      }
      if (t_last >= t_first) {
        if (lane < head + tail) {
          const int t = lane < head ? lane : head + 4 * nvec + (lane - head);
          wz[t] = 0.0f;
        }
        float4* wz4 = reinterpret_cast<float4*>(wz + head);
        const int v0 = max(0, (t_first - head) >> 2), v1 = min(nvec - 1, (t_last - head) >> 2);
        for (int v = v0 + lane; v <= v1; v += 32) wz4[v] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    }
    // q = wfs * t_sampling, on the window only (0 * t_sampling = 0 elsewhere)
    if (t_last >= t_first) {
      if (lane < head + tail) {
        const int t = lane < head ? lane : head + 4 * nvec + (lane - head);
        c[t] = __fmul_rn(c[t], p.t_sampling);
      }
      const int v0 = max(0, (t_first - head) >> 2), v1 = min(nvec - 1, (t_last - head) >> 2);
      for (int v = v0 + lane; v <= v1; v += 32) {
        float4 q = c4[v];
        q.x = __fmul_rn(q.x, p.t_sampling); q.y = __fmul_rn(q.y, p.t_sampling);
        q.z = __fmul_rn(q.z, p.t_sampling); q.w = __fmul_rn(q.w, p.t_sampling);
        c4[v] = q;
      }
    }
  }
  if (lane == 0) { s_first[wid] = t_first; s_last[wid] = t_last; }
  __syncthreads();
  // sequential running sums of the CTA's rows, lane <-> row.  Adding the zero samples before the first and after the last
  // non-zero tick does not change a float32 sum, so only [t_first, t_last] is walked.
  if (wid == 0 && lane < FEE_WARPS) {
    float tot = 0.0f;
    const int rl = blockIdx.x * FEE_WARPS + lane;
    if (rl < F.npix) {
      const float* wl = F.wfs + (int64_t)rl * F.stride;
      const int hl = min((int)(((16u - (unsigned)((uintptr_t)wl & 15u)) & 15u) >> 2), Nt);
      tot = fee_serial_sum(smem + (size_t)lane * rstride + ((4 - hl) & 3), hl, (Nt - hl) >> 2, s_first[lane], s_last[lane]);
    }
    s_total[lane] = tot;
  }
  __syncthreads();
  if (!live) return;
  const float total = s_total[wid];
  // Outside [lo, hi] the running sum is constant (0 before the first non-zero sample, the total after the last one) and
  // stays constant under the subtract-and-clamp of every pass: those two regions are carried as the scalars hv / tv and
  // only the window is kept in shared memory and walked by the ten passes (~150 ticks of 2000 on a track's pixel).
  const int lo = min(t_first, Nt), hi = t_last;  // empty row: lo = Nt, hi = -1
  float hv = 0.0f, tv = total;
  auto cget = [&](int t) { return t < lo ? hv : (t > hi ? tv : c[t]); };
  const float thr = p.discrimination_threshold;
  const int interval = p.hold_interval;
  const int nmax = p.max_adc_values;
  const float* nz = F.noise;
  const int64_t np = F.npix;
  float base = nz ? __fmul_rn(nz[row], p.reset_noise_charge) : 0.0f;
  // pixel geometry (id2pixel with Python floor semantics; negative ids give event < 0)
  const int pid = F.unique_pixels[row];
  const int nx = p.n_pixels_x, ny = p.n_pixels_y, ntpc = p.n_tpc;
  // three floor divisions instead of five: floor(floor(a / b) / c) == floor(a / (b c)) for positive b, c
  const int t1 = floordiv_pos(pid, nx);
  const int xp = pid - t1 * nx;
  const int t2 = floordiv_pos(t1, ny);      // == floordiv(pid, nx * ny)
  const int yp = t1 - t2 * ny;
  const int ev = floordiv_pos(t2, ntpc);    // == floordiv(pid, nx * ny * ntpc)
  const int plane = t2 - ev * ntpc;
  const float z_anode = p.tpc_borders[plane][2][0], z_high = p.tpc_borders[plane][2][1];
  const float dz = __fsub_rn(z_high, z_anode);
  const float sgn = dz > 0.0f ? 1.0f : (dz < 0.0f ? -1.0f : 0.0f);
  const float adc_scale_den = p.v_ref_minus_cm;
  unsigned hit_mask = 0, spos_mask = 0, slope_mask = 0;
  int n_valid = 0;
  // The crossing scan reads pairs (t, t+1) that touch the window, i.e. ticks lo-1 .. hi+1: the two constant regions are
  // mirrored into the zero samples next to the window (c[lo-1] = hv, c[hi+1] = tv, refreshed after every pass) so that the
  // scan loads shared memory without a bounds select per sample.
  if (lane == 0) {
    if (lo >= 1 && lo <= Nt) c[lo - 1] = hv;
    if (hi >= 0 && hi + 1 < Nt) c[hi + 1] = tv;
  }
  __syncwarp();
  for (int it = 0; it < nmax; ++it) {
    // first t with q_sum[t] <= thr <= q_sum[t+1] (fee_jax.py:200-203); fill value Nt-2
    int idx_t = Nt - 2;
    bool found = false;
    if (lo >= 2 && __fadd_rn(base, hv) == thr) { idx_t = 0; found = true; }  // both samples in the constant head
    if (!found) {
      const int t_end = min(hi, Nt - 2);  // pairs (t, t+1) that touch the window
      for (int t0 = max(lo - 1, 0); t0 <= t_end; t0 += 32) {
        const int t = t0 + lane;
        bool hit = false;
        if (t <= t_end) {
          const float a = __fadd_rn(base, c[t]), b = __fadd_rn(base, c[t + 1]);
          hit = (b >= thr) && (a <= thr);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) { idx_t = t0 + __ffs(m) - 1; found = true; break; }
      }
    }
    if (!found && hi + 1 <= Nt - 2 && __fadd_rn(base, tv) == thr) idx_t = max(hi + 1, 0);  // both samples in the constant tail
    int end = idx_t + 1 + interval;
    if (end >= Nt) end = Nt - 1;
    const float q_nn = cget(end);
    const float q_vals = __fadd_rn(base, q_nn);
    const float extra = nz ? __fmul_rn(nz[np * (1 + it) + row], p.uncorrelated_noise_charge) : 0.0f;
    float adc = (q_nn != 0.0f) ? __fadd_rn(q_vals, extra) : q_nn;
    const bool cond = (adc < thr) || (idx_t == Nt - 2);
    if (cond) adc = 0.0f;
    const float ic = cond ? (float)(Nt - 2) : (float)idx_t;
    base = nz ? (cond ? __fmul_rn(nz[np * (21 + it) + row], p.uncorrelated_noise_charge)
                      : __fmul_rn(nz[np * (11 + it) + row], p.reset_noise_charge))
              : 0.0f;
    int end2 = idx_t + 1 + interval + 1;
    if (end2 >= Nt) end2 = Nt - 1;
    const float sub = cget(end2);
    __syncwarp();
    hv = __fsub_rn(hv, sub); hv = (hv < 0.0f) ? 0.0f : hv;
    tv = __fsub_rn(tv, sub); tv = (tv < 0.0f) ? 0.0f : tv;
    float vmax = fmaxf(lo > 0 ? hv : 0.0f, hi < Nt - 1 ? tv : 0.0f);
    for (int t = lo + lane; t <= hi; t += 32) {
      float v = __fsub_rn(c[t], sub);
      v = (v < 0.0f) ? 0.0f : v;
      c[t] = v;
      vmax = fmaxf(vmax, v);
    }
    if (lane == 0) {
      if (lo >= 1 && lo <= Nt) c[lo - 1] = hv;
      if (hi >= 0 && hi + 1 < Nt) c[hi + 1] = tv;
    }
    __syncwarp();
    // digitize (fee_jax.py:68-69)
    float inner = __fsub_rn(__fadd_rn(__fmul_rn(adc, p.gain), p.v_pedestal), p.v_cm);
    float dig = __fdiv_rn(__fmul_rn(fmaxf(inner, 0.0f), p.adc_counts), adc_scale_den);
    const bool sloped = (inner > 0.0f) && (dig < p.adc_counts);
    dig = fminf(dig, p.adc_counts);
    const float hp = (ic < (float)(Nt - 3)) ? 1.0f : 0.0f;
    // z_anode + ticks * (t_sampling * v) * sign: XLA folds the two scalars (pinned by the goldens' pix_z)
    const float pz = __fadd_rn(z_anode, __fmul_rn(__fmul_rn(ic, p.ts_vdrift), sgn));
    if (!cond) hit_mask |= 1u << it;
    if (sub > 0.0f) spos_mask |= 1u << it;
    if (sloped && !cond) slope_mask |= 1u << it;
    if (hp > p.hit_prob_threshold && ev >= 0 && pid >= 0) ++n_valid;
    if (lane == 0) {
      F.adc[(int64_t)row * nmax + it] = dig;
      F.ticks[(int64_t)row * nmax + it] = ic;
      F.pixel_z[(int64_t)row * nmax + it] = pz;
      if (F.saved) {
        F.saved[(int64_t)row * 32 + it] = (float)idx_t;
        F.saved[(int64_t)row * 32 + 13 + it] = adc;  // integrated charge before the digitiser = get_adc_values' own output
      }
    }
    // Early exit (noise-free only): after the clamp the row is >= 0, so every later subtraction is >= 0 and the row can only
    // shrink; once its maximum is below the threshold no later pass can find a crossing.  The remaining passes of the
    // reference all return (adc 0, tick Nt-2): write them directly.
    if (!nz && it + 1 < nmax) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (vmax < thr) {
        const float inner0 = __fsub_rn(__fadd_rn(__fmul_rn(0.0f, p.gain), p.v_pedestal), p.v_cm);
        const float dig0 = fminf(__fdiv_rn(__fmul_rn(fmaxf(inner0, 0.0f), p.adc_counts), adc_scale_den), p.adc_counts);
        const float ic0 = (float)(Nt - 2);
        const float pz0 = __fadd_rn(z_anode, __fmul_rn(__fmul_rn(ic0, p.ts_vdrift), sgn));
        const bool counts_valid = (ic0 < (float)(Nt - 3) ? 1.0f : 0.0f) > p.hit_prob_threshold && ev >= 0 && pid >= 0;
        for (int k = it + 1 + lane; k < nmax; k += 32) {
          F.adc[(int64_t)row * nmax + k] = dig0;
          F.ticks[(int64_t)row * nmax + k] = ic0;
          F.pixel_z[(int64_t)row * nmax + k] = pz0;
          if (F.saved) { F.saved[(int64_t)row * 32 + k] = ic0; F.saved[(int64_t)row * 32 + 13 + k] = 0.0f; }
        }
        if (counts_valid) n_valid += nmax - it - 1;
        break;
      }
    }
  }
  if (lane == 0) {
    // fma(pitch_index, pitch, border) + pitch/2: XLA:CPU contracts the multiply-add (pinned by the goldens' pix_x/pix_y)
    F.pixel_x[row] = __fadd_rn(__fmaf_rn((float)xp, p.pixel_pitch, p.tpc_borders[plane][0][0]), p.half_pitch);
    F.pixel_y[row] = __fadd_rn(__fmaf_rn((float)yp, p.pixel_pitch, p.tpc_borders[plane][1][0]), p.half_pitch);
    F.event[row] = ev;
    F.row_counts[row] = n_valid;
    if (F.saved) {
      F.saved[(int64_t)row * 32 + 10] = __int_as_float((int)hit_mask);
      F.saved[(int64_t)row * 32 + 11] = __int_as_float((int)spos_mask);
      F.saved[(int64_t)row * 32 + 12] = __int_as_float((int)slope_mask);
    }
  }
}

// exclusive scan of row_counts -> offsets, total.  Two launches over blocks of 4096 rows (1024 threads x 4 rows): per-block
// totals, then every block adds the totals in front of it (a few dozen values) to its own in-block scan — the single looped
// CTA this replaces took 0.07-0.14 ms for the 268 k rows of a spill.
constexpr int SCAN_ROWS = 4096;

__device__ __forceinline__ int block_sum_1024(int v, int* wsum) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) wsum[wid] = v;
  __syncthreads();
  int t = wsum[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(1024) k_count_block_sums(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ bsum) {
  __shared__ int wsum[32];
  const int i = blockIdx.x * SCAN_ROWS + 4 * threadIdx.x;
  int v = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) v += i + k < n ? counts[i + k] : 0;
  const int tot = block_sum_1024(v, wsum);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t* __restrict__ counts, int n, const int32_t* __restrict__ bsum,
                                                      int32_t* __restrict__ offsets, int32_t* __restrict__ total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int before = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += 1024) before += bsum[b];
  const int carry = block_sum_1024(before, wsum);
  const int i = blockIdx.x * SCAN_ROWS + 4 * threadIdx.x;
  int v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = i + k < n ? counts[i + k] : 0;
  const int tsum = v[0] + v[1] + v[2] + v[3];
  int inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int x = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += u;
    }
    wsum[lane] = x;
  }
  __syncthreads();
  int off = carry + (wid > 0 ? wsum[wid - 1] : 0) + inc - tsum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (i + k < n) offsets[i + k] = off;
    off += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = carry + wsum[31];
}

struct CompactArgs {
  const float* adc; const float* ticks; const float* pixel_z; const float* pixel_x; const float* pixel_y;
  const int32_t* event; const int32_t* unique_pixels; const int32_t* offsets;
  float* hit_adc; float* hit_x; float* hit_y; float* hit_z; float* hit_ticks; float* hit_prob;
  int32_t* hit_event; int32_t* hit_pixel;
  int npix, nmax, ntw;
  float hit_prob_threshold;
};

// parse_output (sim_jax.py:620-647): row-major stable compaction of valid (pixel, hit) slots
__global__ void k_compact_hits(const __grid_constant__ CompactArgs C) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= C.npix) return;
  const int pid = C.unique_pixels[row], ev = C.event[row];
  if (ev < 0 || pid < 0) return;
  int o = C.offsets[row];
  for (int k = 0; k < C.nmax; ++k) {
    float tk = C.ticks[(int64_t)row * C.nmax + k];
    float hp = (tk < (float)(C.ntw - 3)) ? 1.0f : 0.0f;
    if (hp > C.hit_prob_threshold) {
      C.hit_adc[o] = C.adc[(int64_t)row * C.nmax + k];
      C.hit_x[o] = C.pixel_x[row];
      C.hit_y[o] = C.pixel_y[row];
      C.hit_z[o] = C.pixel_z[(int64_t)row * C.nmax + k];
      C.hit_ticks[o] = tk;
      C.hit_prob[o] = hp;
      C.hit_event[o] = ev;
      C.hit_pixel[o] = pid;
      ++o;
    }
  }
}

// VJP of get_adc_values + digitize.  With c^k the running sum after k subtractions, hit k samples
// v_k = c^k[e_k] and subtracts S_k = c^k[e2_k]; c^{k+1} = relu(c^k - S_k).  A value that is still positive
// at step k was never clamped, so the adjoint of c^0 is supported on the <= 20 positions {e_k, e2_k}:
//   T_k = g_k*hit_k + Sbar_k*[S_k > 0],   Sbar_k = -sum_{k' > k} T_k',
//   cbar[e_k] += g_k*hit_k,  cbar[e2_k] += Sbar_k*[S_k > 0],   g_wfs[t] = t_sampling * sum_{t' >= t} cbar[t'].
// The <= 20 positions are ascending in k (hit k+1 is found after the subtraction tick of hit k), so g_wfs is a step
// function of t: on (pos[m-1], pos[m]] it equals t_sampling * (val[m] + val[m+1] + ...).  One warp per row: every lane
// replays the ten-step recursion in registers, lane l keeps entry l, the suffix sums are formed in the same order as
// the plain sum over k (bit-identical to it), and the warp writes the row range by range -- no per-tick loop over the
// list.  Rows whose list is not ascending (never seen) take the plain sum.
__global__ void __launch_bounds__(128)
k_fee_backward(const float* __restrict__ g_adc, const float* __restrict__ saved, int npix, int ntw,
               const __grid_constant__ larnd_params_t p, float* __restrict__ g_wfs, int64_t g_stride, const int raw_charge) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= npix) return;
  const int nmax = p.max_adc_values, nent = 2 * nmax;
  const float* sv = saved + (int64_t)row * 32;
  const unsigned hit_mask = (unsigned)__float_as_int(sv[10]);
  const unsigned spos_mask = (unsigned)__float_as_int(sv[11]);
  const unsigned slope_mask = (unsigned)__float_as_int(sv[12]);
  const float slope = p.gain * p.adc_counts / p.v_ref_minus_cm;
  float* out = g_wfs + (int64_t)row * g_stride;
  if (hit_mask == 0u) {  // no hit: zero gradient
    for (int t = lane; t < ntw; t += 32) out[t] = 0.0f;
    return;
  }
  int mypos = -1;       // entry 2k: sampling tick e_k, entry 2k+1: subtraction tick e2_k
  float myval = 0.0f;
  float tail = 0.0f;    // sum_{k' > k} T_k'
#pragma unroll
  for (int k = LARND_MAX_ADC - 1; k >= 0; --k) {
    if (k < nmax) {
      const int idx_t = (int)sv[k];
      int e = idx_t + 1 + p.hold_interval; if (e >= ntw) e = ntw - 1;
      int e2 = idx_t + 2 + p.hold_interval; if (e2 >= ntw) e2 = ntw - 1;
      // raw_charge: the upstream gradient is w.r.t. the integrated charge (get_adc_values), not the digitised ADC
      const bool live = ((hit_mask >> k) & 1u) && (raw_charge || ((slope_mask >> k) & 1u));
      const float gk = live ? g_adc[(int64_t)row * nmax + k] * (raw_charge ? 1.0f : slope) : 0.0f;
      const float sbar = ((spos_mask >> k) & 1u) ? -tail : 0.0f;
      if (lane == 2 * k) { mypos = e; myval = gk; }
      if (lane == 2 * k + 1) { mypos = e2; myval = sbar; }
      tail += gk + sbar;
    }
  }
  const int nextpos = __shfl_down_sync(0xffffffffu, mypos, 1);
  const bool ascending = __all_sync(0xffffffffu, lane >= nent - 1 || mypos <= nextpos);
  if (ascending) {
    float S = 0.0f;  // lane l: val[l] + val[l+1] + ... in ascending k, like the plain sum with its leading zeros
#pragma unroll
    for (int k = 0; k < 2 * LARND_MAX_ADC; ++k) {
      const float v = __shfl_sync(0xffffffffu, myval, k);
      if (k >= lane && k < nent) S += v;
    }
    int prev = -1;
    for (int m = 0; m < nent; ++m) {
      const int pm = __shfl_sync(0xffffffffu, mypos, m);
      const float sm = __shfl_sync(0xffffffffu, S, m) * p.t_sampling;
      for (int t = prev + 1 + lane; t <= pm; t += 32) out[t] = sm;
      prev = max(prev, pm);
    }
    for (int t = prev + 1 + lane; t < ntw; t += 32) out[t] = 0.0f;
  } else {
    for (int t0 = 0; t0 < ntw; t0 += 32) {  // warp-uniform trip count: the shuffles need every lane
      const int t = t0 + lane;
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < 2 * LARND_MAX_ADC; ++k) {
        const int pk = __shfl_sync(0xffffffffu, mypos, k);
        const float vk = __shfl_sync(0xffffffffu, myval, k);
        if (k < nent) s += (pk >= t) ? vk : 0.0f;
      }
      if (t < ntw) out[t] = s * p.t_sampling;
    }
  }
}

// The same VJP as k_fee_backward in its COMPACT form: the (position, coefficient) list itself, consumed directly by
// larnd_lut_backward_steps — the dense (npix, n_ticks) gradient (2 GB at spill size) is never written or read.
// thread <-> row; events with a zero coefficient are dropped; rows of pixel ids < 0 get no event (parse_output drops
// their hits, sim_jax.py:621, so no loss built on hits reaches them).
__global__ void __launch_bounds__(128)
k_fee_backward_steps(const float* __restrict__ g_adc, const float* __restrict__ saved, const int32_t* __restrict__ unique_pixels,
                     int npix, int ntw, const __grid_constant__ larnd_params_t p, int32_t* __restrict__ rec, const int raw_charge) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= npix) return;
  const int nmax = p.max_adc_values;
  const float* sv = saved + (int64_t)row * 32;
  const unsigned hit_mask = (unsigned)__float_as_int(sv[10]);
  int count = 0;
  {
    // unused entries read as "no event" (position -1, coefficient 0): consumers may walk a record without masking by its count
    int32_t* r = rec + (int64_t)row * LARND_STEPS_WORDS;
#pragma unroll
    for (int k = 0; k < LARND_STEPS_MAX; ++k) { r[k] = -1; r[LARND_STEPS_MAX + k] = 0; }
  }
  if (hit_mask != 0u && unique_pixels[row] >= 0) {
    const unsigned spos_mask = (unsigned)__float_as_int(sv[11]);
    const unsigned slope_mask = (unsigned)__float_as_int(sv[12]);
    const float slope = p.gain * p.adc_counts / p.v_ref_minus_cm;
    int32_t* pos = rec + (int64_t)row * LARND_STEPS_WORDS;
    float* val = reinterpret_cast<float*>(pos + LARND_STEPS_MAX);
    float tail = 0.0f;   // sum_{k' > k} T_k' (same recursion and order as k_fee_backward)
    for (int k = nmax - 1; k >= 0; --k) {
      const int idx_t = (int)sv[k];
      int e = idx_t + 1 + p.hold_interval; if (e >= ntw) e = ntw - 1;
      int e2 = idx_t + 2 + p.hold_interval; if (e2 >= ntw) e2 = ntw - 1;
      const bool live = ((hit_mask >> k) & 1u) && (raw_charge || ((slope_mask >> k) & 1u));
      const float gk = live ? g_adc[(int64_t)row * nmax + k] * (raw_charge ? 1.0f : slope) : 0.0f;
      const float sbar = ((spos_mask >> k) & 1u) ? -tail : 0.0f;
      // full-row column of FEE tick t is t + 1 (simulate_wfs returns wfs[:, 1:])
      if (sbar != 0.0f) { pos[count] = e2 + 1; val[count] = sbar * p.t_sampling; ++count; }
      if (gk != 0.0f) { pos[count] = e + 1; val[count] = gk * p.t_sampling; ++count; }
      tail += gk + sbar;
    }
  }
  rec[(int64_t)row * LARND_STEPS_WORDS + 2 * LARND_STEPS_MAX] = count;
}

}  // namespace

extern "C" size_t larnd_fee_steps_bytes(int32_t npix) { return larnd_steps_layout(npix > 0 ? npix : 0, nullptr, nullptr); }

extern "C" int larnd_fee_backward_steps(const float* g_adc_d, const float* saved_d, const int32_t* unique_pixels_d, int32_t npix,
                                        const larnd_params_t* params, void* steps_d, size_t steps_bytes, int32_t raw_charge,
                                        void* stream) {
  if (!g_adc_d || !saved_d || !unique_pixels_d || !params || !steps_d || npix < 0 || steps_bytes < larnd_fee_steps_bytes(npix)) {
    larnd_set_error("larnd_fee_backward_steps: null argument or steps buffer smaller than larnd_fee_steps_bytes()");
    return LARND_E_ARG;
  }
  if (params->max_adc_values > LARND_MAX_ADC) { larnd_set_error("MAX_ADC_VALUES > %d unsupported", LARND_MAX_ADC); return LARND_E_ARG; }
  if (npix == 0) return LARND_OK;
  StepsView v;
  larnd_steps_layout(npix, steps_d, &v);
  k_fee_backward_steps<<<(npix + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      g_adc_d, saved_d, unique_pixels_d, npix, params->n_ticks - 1, *params, const_cast<int32_t*>(v.rec), raw_charge != 0);
  LARND_LAUNCH_CHECK("k_fee_backward_steps");
  return LARND_OK;
}

// row counts | offsets | per-block totals of the hit-offset scan
extern "C" size_t larnd_fee_scratch_bytes(int32_t npix) {
  return align_up(((size_t)npix * 2 + (size_t)npix / SCAN_ROWS + 1) * sizeof(int32_t) + 64, 256);
}

extern "C" int larnd_fee_forward(const float* wfs_d, int64_t wfs_row_stride, const int32_t* unique_pixels_d, int32_t npix,
                                 const larnd_params_t* params, const float* noise_d, float* adc_d, float* ticks_d,
                                 float* pixel_z_d, float* pixel_x_d, float* pixel_y_d, int32_t* event_d, float* saved_d,
                                 float* hit_adc_d, float* hit_x_d, float* hit_y_d, float* hit_z_d, float* hit_ticks_d,
                                 float* hit_prob_d, int32_t* hit_event_d, int32_t* hit_pixel_d, int32_t* n_valid_d,
                                 void* scratch_d, size_t scratch_bytes, void* stream) {
  return larnd_fee_forward_ex(wfs_d, wfs_row_stride, unique_pixels_d, npix, params, noise_d, adc_d, ticks_d, pixel_z_d, pixel_x_d,
                              pixel_y_d, event_d, saved_d, hit_adc_d, hit_x_d, hit_y_d, hit_z_d, hit_ticks_d, hit_prob_d, hit_event_d,
                              hit_pixel_d, n_valid_d, scratch_d, scratch_bytes, 0, stream);
}

extern "C" int larnd_fee_forward_ex(const float* wfs_d, int64_t wfs_row_stride, const int32_t* unique_pixels_d, int32_t npix,
                                    const larnd_params_t* params, const float* noise_d, float* adc_d, float* ticks_d,
                                    float* pixel_z_d, float* pixel_x_d, float* pixel_y_d, int32_t* event_d, float* saved_d,
                                    float* hit_adc_d, float* hit_x_d, float* hit_y_d, float* hit_z_d, float* hit_ticks_d,
                                    float* hit_prob_d, int32_t* hit_event_d, int32_t* hit_pixel_d, int32_t* n_valid_d,
                                    void* scratch_d, size_t scratch_bytes, int32_t fee_flags, void* stream) {
  if (!wfs_d || !unique_pixels_d || !params || !adc_d || !ticks_d || !pixel_z_d || !pixel_x_d || !pixel_y_d || !event_d ||
      !n_valid_d || !scratch_d) {
    larnd_set_error("larnd_fee_forward: null argument");
    return LARND_E_ARG;
  }
  if (params->max_adc_values < 1 || params->max_adc_values > LARND_MAX_ADC) {
    larnd_set_error("larnd_fee_forward: MAX_ADC_VALUES=%d unsupported (1..%d)", params->max_adc_values, LARND_MAX_ADC);
    return LARND_E_ARG;
  }
  if (scratch_bytes < larnd_fee_scratch_bytes(npix)) {
    larnd_set_error("larnd_fee_forward: scratch too small");
    return LARND_E_CAPACITY;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int ntw = params->n_ticks - 1;
  if (npix == 0) { LARND_CUDA(cudaMemsetAsync(n_valid_d, 0, sizeof(int32_t), st)); return LARND_OK; }
  FeeArgs F;
  F.wfs = wfs_d; F.stride = wfs_row_stride; F.unique_pixels = unique_pixels_d; F.npix = npix; F.ntw = ntw;
  F.noise = noise_d; F.adc = adc_d; F.ticks = ticks_d; F.pixel_z = pixel_z_d; F.pixel_x = pixel_x_d; F.pixel_y = pixel_y_d;
  F.event = event_d; F.saved = saved_d;
  F.row_counts = reinterpret_cast<int32_t*>(scratch_d);
  int32_t* offsets = F.row_counts + npix;
  // rows per CTA (= warps): measured at the 10 M-segment workload 1 / 2 / 4 / 8 rows -> 0.574 / 0.576 / 0.543 / 0.622 ms
  int nw = LARND_FEE_WARPS;
  // rows + one mbarrier per warp + window bounds and total per row
  size_t smem = (size_t)nw * FEE_ROW_STRIDE(ntw) * sizeof(float) + nw * (sizeof(uint64_t) + 2 * sizeof(int) + sizeof(float));
  {
    // TMA row loads need 16-byte aligned sources: rows whose start is `sh` floats past a 16-byte boundary are copied from
    // the aligned address below (see the header: those floats and the round-up at the end must be readable)
    const int sh = (int)(((uintptr_t)wfs_d & 15u) >> 2);
    F.bulk = (((uintptr_t)wfs_d & 3u) == 0 && wfs_row_stride % 4 == 0 && ((sh + ntw + 3) & ~3) <= wfs_row_stride) ? 1 : 0;
  }
  F.clear = (fee_flags & LARND_FEE_CLEAR_WFS) ? 1 : 0;
  if (F.clear && !F.bulk) {
    larnd_set_error("larnd_fee_forward_ex: LARND_FEE_CLEAR_WFS needs the padded waveform layout (16-byte aligned buffer, row stride a "
                    "multiple of four floats)");
    return LARND_E_ARG;
  }
  static bool attr_set = false;
  if (!attr_set) {
    LARND_CUDA(cudaFuncSetAttribute(k_fee_forward<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LARND_CUDA(cudaFuncSetAttribute(k_fee_forward<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LARND_CUDA(cudaFuncSetAttribute(k_fee_forward<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LARND_CUDA(cudaFuncSetAttribute(k_fee_forward<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  while (smem > 200 * 1024 && nw > 1) {  // very long readouts: fewer rows per CTA
    nw >>= 1;
    smem = (size_t)nw * FEE_ROW_STRIDE(ntw) * sizeof(float) + nw * (sizeof(uint64_t) + 2 * sizeof(int) + sizeof(float));
  }
  if (smem > 200 * 1024) { larnd_set_error("larnd_fee_forward: n_ticks too large for shared memory"); return LARND_E_ARG; }
  prof_begin(3, st);
  const unsigned grid = (unsigned)((npix + nw - 1) / nw);
  if (nw == 8) k_fee_forward<8><<<grid, 256, smem, st>>>(F, *params);
  else if (nw == 4) k_fee_forward<4><<<grid, 128, smem, st>>>(F, *params);
  else if (nw == 2) k_fee_forward<2><<<grid, 64, smem, st>>>(F, *params);
  else k_fee_forward<1><<<grid, 32, smem, st>>>(F, *params);
  prof_end(3, st);
  LARND_LAUNCH_CHECK("k_fee_forward");
  {
    int32_t* bsum = offsets + npix;
    const int nblk = (npix + SCAN_ROWS - 1) / SCAN_ROWS;
    k_count_block_sums<<<nblk, 1024, 0, st>>>(F.row_counts, npix, bsum);
    LARND_LAUNCH_CHECK("k_count_block_sums");
    k_scan_counts<<<nblk, 1024, 0, st>>>(F.row_counts, npix, bsum, offsets, n_valid_d);
    LARND_LAUNCH_CHECK("k_scan_counts");
  }
  if (hit_adc_d) {
    CompactArgs C;
    C.adc = adc_d; C.ticks = ticks_d; C.pixel_z = pixel_z_d; C.pixel_x = pixel_x_d; C.pixel_y = pixel_y_d; C.event = event_d;
    C.unique_pixels = unique_pixels_d; C.offsets = offsets;
    C.hit_adc = hit_adc_d; C.hit_x = hit_x_d; C.hit_y = hit_y_d; C.hit_z = hit_z_d; C.hit_ticks = hit_ticks_d; C.hit_prob = hit_prob_d;
    C.hit_event = hit_event_d; C.hit_pixel = hit_pixel_d;
    C.npix = npix; C.nmax = params->max_adc_values; C.ntw = ntw; C.hit_prob_threshold = params->hit_prob_threshold;
    k_compact_hits<<<(npix + 127) / 128, 128, 0, st>>>(C);
    LARND_LAUNCH_CHECK("k_compact_hits");
  }
  return LARND_OK;
}

extern "C" int larnd_fee_backward(const float* g_adc_d, const float* ticks_d, const float* saved_d, int32_t npix,
                                  const larnd_params_t* params, float* g_wfs_d, int64_t g_row_stride, int32_t raw_charge,
                                  void* stream) {
  (void)ticks_d;
  if (!g_adc_d || !saved_d || !params || !g_wfs_d) { larnd_set_error("larnd_fee_backward: null argument"); return LARND_E_ARG; }
  if (npix == 0) return LARND_OK;
  k_fee_backward<<<(npix + 3) / 4, 128, 0, (cudaStream_t)stream>>>(g_adc_d, saved_d, npix, params->n_ticks - 1, *params,
                                                                 g_wfs_d, g_row_stride, raw_charge != 0);
  LARND_LAUNCH_CHECK("k_fee_backward");
  return LARND_OK;
}
