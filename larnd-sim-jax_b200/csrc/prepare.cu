// K1 segment_prepare + K2 unique/renumber (bitmap + popcount scan), sm_100a.
//
// Replaces the XLA lowering of simulate_drift_new (reference sim_jax.py:375-453: shift_tracks :109-119,
// quench quenching_jax.py:38-75, drift drifting_jax.py:19-58, get_bin_shifts detsim_jax.py:494-512,
// density_2d detsim_jax.py:332-354, pixel2id detsim_jax.py:232-244) and of the
// jnp.unique / sort / searchsorted block of simulate_wfs (sim_jax.py:717-725).
//
// Everything that feeds an integer result (bins, pixel ids, ticks, template index) is written with
// explicit round-to-nearest intrinsics so that nvcc cannot contract mul+add into FMA: the reference
// evaluates these op by op in float32 and the ids must match bit for bit.
#include "larnd_common.cuh"
#include "segment_physics.cuh"
#include "chop_math.cuh"

namespace {

constexpr int PREP_THREADS = 128;

// One segment: physics, bins, diffusion weights, tick / template records (tr: its row, cols: where the ten columns sit in it).
// Returns the main-pixel id.
__device__ __forceinline__ int prepare_segment(const float* tr, const larnd_columns_t& cols, const larnd_params_t& p, int nt, int bank_ntpl,
                                               float* __restrict__ rec, int64_t n, int64_t s, int32_t* __restrict__ counts) {
  const SegPhys ph = segment_physics(tr, cols, p);
  const float x = ph.x, y = ph.y, z = ph.z, q = ph.q, td = ph.td, sl_cm = ph.sl_cm, sT = ph.sT;
  const float recomb = ph.recomb, xi = ph.xi, cos2 = ph.cos2, z_anode = ph.z_anode, z_cath = ph.z_cath;
  const int plane = ph.plane;
  const bool inside = ph.inside;
  // sub-pixel bins and in-bin position
  float xr = fsub(x, p.tpc_borders[plane][0][0]);
  float yr = fsub(y, p.tpc_borders[plane][1][0]);
  int bx = (int)floor_divide_f(xr, p.bin_width);
  int by = (int)floor_divide_f(yr, p.bin_width);
  float x0 = remainder_f(xr, p.bin_width);
  float y0 = remainder_f(yr, p.bin_width);
  // transverse diffusion weights: 5 bin integrals per axis, outer edges forced to -1/+1
  const float s2sig = fmul(1.41421354f, sT);
  float ex[LARND_NB_TRAN_BINS + 1], ey[LARND_NB_TRAN_BINS + 1];
  ex[0] = ey[0] = -1.0f;
  ex[LARND_NB_TRAN_BINS] = ey[LARND_NB_TRAN_BINS] = 1.0f;
#pragma unroll
  for (int k = 1; k < LARND_NB_TRAN_BINS; ++k) {
    ex[k] = erff(fdiv(fsub(p.tran_bin_edges[k], x0), s2sig));
    ey[k] = erff(fdiv(fsub(p.tran_bin_edges[k], y0), s2sig));
  }
#pragma unroll
  for (int k = 0; k < LARND_NB_TRAN_BINS; ++k) {
    rec[(int64_t)(LARND_F_WX0 + k) * n + s] = fmul(0.5f, fsub(ex[k + 1], ex[k]));
    rec[(int64_t)(LARND_F_WY0 + k) * n + s] = fmul(0.5f, fsub(ey[k + 1], ey[k]));
  }
  // time to the cathode -> tick + fraction (sim_jax.py:418-420,157-159)
  float t0 = fdiv(fabsf(fsub(z, z_cath)), p.vdrift);
  float ft = fdiv(t0, p.t_sampling);
  int ct = (int)floorf(ft);
  ct = max(0, min(ct, nt - 1));
  float frac = fsub(ft, (float)ct);
  // longitudinal diffusion in ticks -> template index + Lagrange weights (sim_jax.py:423,162-168)
  float sl = fdiv(fdiv(sl_cm, p.vdrift), p.t_sampling);
  int lo = 0, hi = p.n_templates;  // searchsorted side='left': number of template values < sl
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (p.long_diff_template[mid] < sl) lo = mid + 1; else hi = mid;
  }
  int idx = max(1, min(lo, p.n_templates - 2));
  // a bank truncated to fewer rows than long_diff_template (tests bound memory that way) is valid as long as no segment
  // needs a missing row; one that does is flagged (bit 2 of counts[2]) and clamped so that nothing reads out of bounds
  if (idx + 1 >= bank_ntpl) { if (q != 0.0f) atomicOr(counts + 2, 4); idx = max(1, bank_ntpl - 2); }
  float t0v = p.long_diff_template[idx - 1], t1v = p.long_diff_template[idx], t2v = p.long_diff_template[idx + 1];
  float a = fdiv(fmul(fsub(sl, t1v), fsub(sl, t2v)), fmul(fsub(t0v, t1v), fsub(t0v, t2v)));
  float b = fdiv(fmul(fsub(sl, t0v), fsub(sl, t2v)), fmul(fsub(t1v, t0v), fsub(t1v, t2v)));
  float c = fdiv(fmul(fsub(sl, t0v), fsub(sl, t1v)), fmul(fsub(t2v, t0v), fsub(t2v, t1v)));
  int ev = (int)tr[cols.eventID];
  int ep = ev * p.n_tpc + plane;
  const int nb = p.nb_sampling_bins_per_pixel;
  const int pid = pixel2id_dev(floordiv_i(bx, nb), floordiv_i(by, nb), ep, p.n_pixels_x, p.n_pixels_y);
  int flags = (inside ? 1 : 0) | (fsub(z, z_anode) > 0.0f ? 2 : 0) | (fsub(z, z_cath) > 0.0f ? 4 : 0);
  rec[(int64_t)LARND_F_Q * n + s] = q;
  rec[(int64_t)LARND_F_FRAC * n + s] = frac;
  rec[(int64_t)LARND_F_SL * n + s] = sl;
  rec[(int64_t)LARND_F_A * n + s] = a;
  rec[(int64_t)LARND_F_B * n + s] = b;
  rec[(int64_t)LARND_F_C * n + s] = c;
  rec[(int64_t)LARND_F_TD * n + s] = td;
  rec[(int64_t)LARND_F_X0 * n + s] = x0;
  rec[(int64_t)LARND_F_Y0 * n + s] = y0;
  rec[(int64_t)LARND_F_ST * n + s] = sT;
  rec[(int64_t)LARND_F_REC * n + s] = recomb;
  rec[(int64_t)LARND_F_FT * n + s] = ft;
  rec[(int64_t)LARND_F_XI * n + s] = xi;
  rec[(int64_t)LARND_F_COS2 * n + s] = cos2;
  int* irec = reinterpret_cast<int*>(rec);
  irec[(int64_t)LARND_I_T0 * n + s] = nt - p.signal_length - ct;
  irec[(int64_t)LARND_I_IDX * n + s] = idx;
  irec[(int64_t)LARND_I_BX * n + s] = bx;
  irec[(int64_t)LARND_I_BY * n + s] = by;
  irec[(int64_t)LARND_I_EP * n + s] = ep;
  irec[(int64_t)LARND_I_FLAGS * n + s] = flags;
  irec[(int64_t)LARND_I_MAINPIX * n + s] = pid;
  return pid;
}

// mark the main pixels of a warp's segments in the bitmap: one atomic per distinct id per warp
__device__ __forceinline__ void mark_main_pixel(bool pid_ok, int pid, uint32_t* __restrict__ bitmap, int64_t n_words, int pid_offset,
                                                int32_t* __restrict__ counts) {
  unsigned live = __ballot_sync(0xffffffffu, pid_ok);
  if (pid_ok) {
    unsigned peers = __match_any_sync(live, pid);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
      long long bidx = (long long)pid + pid_offset;
      if (bidx < 0 || (bidx >> 5) >= n_words) {
        atomicOr(counts + 2, 2);  // event id outside the declared [-1, n_events) range
      } else {
        uint32_t bit = 1u << (bidx & 31);
        if (!(__ldg(bitmap + (bidx >> 5)) & bit)) atomicOr(bitmap + (bidx >> 5), bit);
      }
    }
  }
}

__global__ void __launch_bounds__(PREP_THREADS)
k_prepare(const float* __restrict__ tracks, int64_t n, const __grid_constant__ larnd_columns_t cols,
          const __grid_constant__ larnd_params_t p, int nt, int bank_ntpl, float* __restrict__ rec, uint32_t* __restrict__ bitmap,
          int64_t n_words, int pid_offset, int32_t* __restrict__ counts) {
  extern __shared__ float srow[];
  const int ncols = cols.ncols;
  const int stride = ncols | 1;  // odd stride: conflict-free row reads
  const int64_t base = (int64_t)blockIdx.x * PREP_THREADS;
  const int rows_here = (int)min((int64_t)PREP_THREADS, n - base);
  const int total = rows_here * ncols;
  const float* src = tracks + base * ncols;
  {
    // (row, column) of element i advance by a fixed step: no integer division per element (it was 25 % of the kernel's
    // instructions, ncu line profile profiles/r2y_k_prepare_lines.txt)
    const int dr = PREP_THREADS / ncols, dc = PREP_THREADS - dr * ncols;
    int r = threadIdx.x / ncols, c = threadIdx.x - r * ncols;
    int i = threadIdx.x;
    // eight coalesced loads in flight per thread before the first store (the kernel's top stall was the load -> store
    // dependency of a one-element loop body: long-scoreboard 5 warps per issue, profiles/r2_final_kernels_ncu_summary.txt)
    for (; i + 7 * PREP_THREADS < total; i += 8 * PREP_THREADS) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(src + i + k * PREP_THREADS);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        srow[r * stride + c] = v[k];
        r += dr;
        c += dc;
        if (c >= ncols) { c -= ncols; ++r; }
      }
    }
    for (; i < total; i += PREP_THREADS) {
      srow[r * stride + c] = __ldg(src + i);  // coalesced stream of the 104-byte records
      r += dr;
      c += dc;
      if (c >= ncols) { c -= ncols; ++r; }
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  const bool active = t < rows_here;
  int pid = 0;
  if (active) pid = prepare_segment(srow + t * stride, cols, p, nt, bank_ntpl, rec, n, base + t, counts);
  mark_main_pixel(active, pid, bitmap, n_words, pid_offset, counts);
}

// The same records straight from the RAW (un-chopped) rows: thread <-> chopped segment s of the batch.  The piece (raw row i,
// index k) is found in the prefix table of larnd_chop_count, its ten simulation columns are formed in registers with
// chop_tracks' arithmetic (chop_math.cuh: bit-identical to k_chop_expand) and handed to prepare_segment — the chopped
// (n, 26) batch (104 B per segment written by the chop kernel and read back here) never exists.  Segments s >= total
// (the batch is sized for `n` slots) are the invalid rows pad_batch appends: eventID -1, everything else 0.
__global__ void __launch_bounds__(PREP_THREADS)
k_prepare_raw(const float* __restrict__ raw, int64_t m, const __grid_constant__ larnd_chop_columns_t cc,
              const __grid_constant__ larnd_columns_t cols, double precision, float prec32, const int64_t* __restrict__ offsets,
              int64_t n, const __grid_constant__ larnd_params_t p, int nt, int bank_ntpl, float* __restrict__ rec,
              uint32_t* __restrict__ bitmap, int64_t n_words, int pid_offset, int32_t* __restrict__ counts) {
  __shared__ int64_t s_i0;
  const int64_t base = (int64_t)blockIdx.x * PREP_THREADS;
  const int64_t total = offsets[m];
  if (total > n) {  // more pieces than segment slots: flagged, the accumulate kernels bail out on counts[2] != 0
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(counts + 2, 8);
    return;
  }
  if (threadIdx.x == 0) {
    int64_t lo = 0, hi = m > 0 ? m - 1 : 0;  // largest i with offsets[i] <= base (every raw row has at least one piece)
    while (lo < hi) {
      const int64_t mid = (lo + hi + 1) >> 1;
      if (offsets[mid] <= base) lo = mid; else hi = mid - 1;
    }
    s_i0 = lo;
  }
  __syncthreads();
  const int64_t s = base + threadIdx.x;
  const bool active = s < n;
  int pid = 0;
  if (active) {
    // ten columns in the ABI's order: eventID, x, y, z, z_start, z_end, dx, dEdx, dE, t0
    float loc[10] = {-1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (s < total) {
      int64_t i = s_i0;
      while (i + 1 < m && offsets[i + 1] <= s) ++i;
      const float* tr = raw + i * cc.ncols;
      const ChopGeom g = chop_geom(tr, cc);
      const long long np = chop_nsteps(g.len, prec32), k = s - offsets[i];
      float mid[3], zs = 0.0f, ze = 0.0f;
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        const int cs = ax == 0 ? cc.x_start : (ax == 1 ? cc.y_start : cc.z_start);
        const int ce = ax == 0 ? cc.x_end : (ax == 1 ? cc.y_end : cc.z_end);
        const double s0 = (double)tr[cs], d = (double)g.dir[ax];
        const float vs = chop_start(s0, d, k, precision), ve = chop_end(s0, d, tr[ce], k, np, precision);
        mid[ax] = __fmul_rn(0.5f, __fadd_rn(vs, ve));
        if (ax == 2) { zs = vs; ze = ve; }
      }
      loc[0] = tr[cols.eventID];
      loc[1] = mid[0]; loc[2] = mid[1]; loc[3] = mid[2];
      loc[4] = zs; loc[5] = ze;
      loc[6] = chop_dx(g, k, np, precision, prec32);
      loc[7] = tr[cols.dEdx];
      loc[8] = chop_dE(tr[cc.dE], g, k, np, precision, prec32);
      loc[9] = tr[cols.t0];
    }
    larnd_columns_t lc;
    lc.ncols = 10; lc.eventID = 0; lc.x = 1; lc.y = 2; lc.z = 3; lc.z_start = 4; lc.z_end = 5; lc.dx = 6; lc.dEdx = 7; lc.dE = 8; lc.t0 = 9;
    pid = prepare_segment(loc, lc, p, nt, bank_ntpl, rec, n, s, counts);
  }
  mark_main_pixel(active, pid, bitmap, n_words, pid_offset, counts);
}

// ---- popcount scan over the bitmap -------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int WORDS_PER_THREAD = LARND_SCAN_WORDS_PER_BLOCK / SCAN_THREADS;  // 8

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  int offset = wid > 0 ? warp_sums[wid - 1] : 0;
  if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return offset + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_block_sums(const uint32_t* __restrict__ bitmap, int64_t n_words, uint32_t* __restrict__ bsums) {
  int64_t w0 = (int64_t)blockIdx.x * LARND_SCAN_WORDS_PER_BLOCK + threadIdx.x * WORDS_PER_THREAD;
  int c = 0;
#pragma unroll
  for (int k = 0; k < WORDS_PER_THREAD; ++k)
    if (w0 + k < n_words) c += __popc(bitmap[w0 + k]);
  int total;
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) bsums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_bsums(uint32_t* __restrict__ bsums, int64_t n_blocks, int32_t* __restrict__ counts) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < n_blocks; b0 += SCAN_THREADS) {
    int64_t i = b0 + threadIdx.x;
    int v = i < n_blocks ? (int)bsums[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    int c = carry;
    if (i < n_blocks) bsums[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[0] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_final(const uint32_t* __restrict__ bitmap, int64_t n_words, const uint32_t* __restrict__ bsums,
             uint32_t* __restrict__ wprefix, int pid_offset, int32_t* __restrict__ counts) {
  int64_t w0 = (int64_t)blockIdx.x * LARND_SCAN_WORDS_PER_BLOCK + threadIdx.x * WORDS_PER_THREAD;
  uint32_t words[WORDS_PER_THREAD];
  int c = 0;
#pragma unroll
  for (int k = 0; k < WORDS_PER_THREAD; ++k) {
    words[k] = (w0 + k < n_words) ? bitmap[w0 + k] : 0u;
    c += __popc(words[k]);
  }
  int ex = block_exclusive_scan(c, nullptr) + (int)bsums[blockIdx.x];
  const int64_t m1_bit = (int64_t)pid_offset - 1;  // bit of pixel id -1
#pragma unroll
  for (int k = 0; k < WORDS_PER_THREAD; ++k) {
    if (w0 + k < n_words) {
      wprefix[w0 + k] = (uint32_t)ex;
      if (w0 + k == (m1_bit >> 5)) counts[1] = ex + __popc(words[k] & ((1u << (m1_bit & 31)) - 1u));  // ids < -1
    }
    ex += __popc(words[k]);
  }
}

// unique_pixels = sort(pad(append(unique(main_pixels), -1), -1)) (sim_jax.py:717-721)
__global__ void k_emit_unique(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ wprefix, int64_t n_words,
                              int pid_offset, int32_t npix, int32_t extra, int32_t* __restrict__ unique_pixels,
                              int32_t* __restrict__ counts) {
  const int n_unique = counts[0];
  if (n_unique + extra > npix) {  // extra = 1: the -1 appended by simulate_wfs (sim_jax.py:718); 0 for the MC path (:363-366)
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(counts + 2, 1);
    return;
  }
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = gtid; w < n_words; w += gsz) {
    uint32_t word = bitmap[w];
    int rank = (int)wprefix[w];
    while (word) {
      int b = __ffs(word) - 1;
      word &= word - 1;
      int pid = (int)((w << 5) + b - pid_offset);
      int row = pid < 0 ? rank : npix - n_unique + rank;  // ids <= -1 keep their rank; ids >= 0 sit at the end
      unique_pixels[row] = pid;
      ++rank;
    }
  }
  // the block of -1 entries: one appended + padding (+ the natural -1 written above, same value)
  const int n_neg = counts[1];
  const int64_t m1_bit = (int64_t)pid_offset - 1;
  const int has_m1 = (bitmap[m1_bit >> 5] >> (m1_bit & 31)) & 1;
  const int n_pos = n_unique - n_neg - has_m1;
  for (int64_t i = n_neg + gtid; i < npix - n_pos; i += gsz) unique_pixels[i] = -1;
}

}  // namespace

int larnd_launch_prepare(const float* tracks, int64_t n, const larnd_columns_t& cols, const larnd_params_t& p,
                         const larnd_lut* lut, const Workspace& ws, int32_t* counts, cudaStream_t st) {
  LARND_CUDA(cudaMemsetAsync(ws.bitmap, 0, ws.n_words * sizeof(uint32_t), st));
  LARND_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st));
  if (n == 0) return LARND_OK;
  const int stride = cols.ncols | 1;
  size_t smem = (size_t)PREP_THREADS * stride * sizeof(float);
  int64_t blocks = (n + PREP_THREADS - 1) / PREP_THREADS;
  prof_begin(0, st);
  larnd_runs_cache_drop(ws.rec);  // the records change: sorted run tables built from the old ones are stale
  k_prepare<<<(unsigned)blocks, PREP_THREADS, smem, st>>>(tracks, n, cols, p, lut ? lut->nt : 0, lut ? lut->ntpl : p.n_templates, ws.rec, ws.bitmap,
                                                         ws.n_words, ws.pid_offset, counts);
  prof_end(0, st);
  LARND_LAUNCH_CHECK("k_prepare");
  return LARND_OK;
}

int larnd_launch_prepare_raw(const float* raw, int64_t m, const larnd_chop_columns_t& cc, const larnd_columns_t& cols, double precision,
                             const int64_t* offsets, int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                             int32_t* counts, cudaStream_t st) {
  LARND_CUDA(cudaMemsetAsync(ws.bitmap, 0, ws.n_words * sizeof(uint32_t), st));
  LARND_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st));
  if (n == 0) return LARND_OK;
  const int64_t blocks = (n + PREP_THREADS - 1) / PREP_THREADS;
  prof_begin(0, st);
  larnd_runs_cache_drop(ws.rec);
  k_prepare_raw<<<(unsigned)blocks, PREP_THREADS, 0, st>>>(raw, m, cc, cols, precision, (float)precision, offsets, n, p, lut ? lut->nt : 0,
                                                          lut ? lut->ntpl : p.n_templates, ws.rec, ws.bitmap, ws.n_words, ws.pid_offset, counts);
  prof_end(0, st);
  LARND_LAUNCH_CHECK("k_prepare_raw");
  return LARND_OK;
}

int larnd_launch_scan(const Workspace& ws, const larnd_params_t& p, int32_t* counts, cudaStream_t st) {
  (void)p;
  k_scan_block_sums<<<(unsigned)ws.n_scan_blocks, SCAN_THREADS, 0, st>>>(ws.bitmap, ws.n_words, ws.bsums);
  LARND_LAUNCH_CHECK("k_scan_block_sums");
  k_scan_bsums<<<1, SCAN_THREADS, 0, st>>>(ws.bsums, ws.n_scan_blocks, counts);
  LARND_LAUNCH_CHECK("k_scan_bsums");
  k_scan_final<<<(unsigned)ws.n_scan_blocks, SCAN_THREADS, 0, st>>>(ws.bitmap, ws.n_words, ws.bsums, ws.wprefix,
                                                                    ws.pid_offset, counts);
  LARND_LAUNCH_CHECK("k_scan_final");
  return LARND_OK;
}

int larnd_launch_unique(const Workspace& ws, const larnd_params_t& p, int32_t npix_capacity, int32_t extra, int32_t* unique_pixels,
                        int32_t* counts, cudaStream_t st) {
  (void)p;
  int64_t blocks = (ws.n_words + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_emit_unique<<<(unsigned)blocks, 256, 0, st>>>(ws.bitmap, ws.wprefix, ws.n_words, ws.pid_offset, npix_capacity, extra,
                                                  unique_pixels, counts);
  LARND_LAUNCH_CHECK("k_emit_unique");
  return LARND_OK;
}
