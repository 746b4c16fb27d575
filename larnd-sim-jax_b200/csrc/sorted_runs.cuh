// Shared by the class-sorted forward (accumulate_sorted.cu) and backward (accumulate_bwd_sorted.cu) kernels:
// run records, the counting sort of the runs by response class, and the tile table.
//
// A *run* = consecutive segments with the same (event, plane, sub-pixel bin, template index) whose start ticks lie
// within SPAN_MAX_S of each other (at most MAXLEN segments).  Its *class* = (template index, bin inside the pixel,
// tick span): all runs of a class read the same response rows for the same unit and have the same number of impulse
// positions, so a tile (<= TR runs of one class) is uniform work.  Inside a class the runs are ordered by the alignment
// shift s = (tmin - 1) mod 4 of their tick frame (four sub-buckets of the counting sort), so the runs of a tile that share
// a shift form a contiguous lane range and can share the register copy of the response rows (accumulate_sorted.cu).
#pragma once
#include "larnd_common.cuh"

namespace {

constexpr int TR = 32;              // runs per tile (lane <-> run in the build phases)
constexpr int KPT = 6;              // impulse positions per run
constexpr int SPAN_MAX_S = KPT - 2;  // max (T0max - T0min) inside a run
// Tiles are ordered by tick span (span-major class key), so runs with few impulse positions form contiguous ranges of the
// tile table and can be served by kernel variants that hold fewer response samples in registers and fit more CTAs per SM:
// gcnt[GC_SPAN + s] = first tile of span s (s = 0 .. SPAN_MAX_S), gcnt[GC_SPAN + SPAN_MAX_S + 1] = number of tiles.
constexpr int GC_SLOW = 3;          // gcnt slot: number of segments the tile kernels leave to the chunk kernels (window ends beyond the readout)
constexpr int GC_SPAN = 8;          // gcnt slots of the per-span tile offsets
constexpr int GC_FWD = 16;          // gcnt slots of the forward kernels' tile counters (one per launch)
constexpr int GC_BWD = 24;          // ... of the backward kernels'
constexpr int MAXLEN = 8;           // segments per run (bounds the divergence of the lane <-> run build loops)
// CTA shape of the forward tile kernel, measured at 10 M segments (accumulate slot, ms): 256 x 4 CTAs/SM 12.54 | 256 x 3 12.23 |
// 288 x 3 11.95 | 320 x 3 11.99 | 224 x 4 12.68 | 192 x 4 13.30 | 384 x 2 13.01 | 512 x 2 13.81.  Nine warps at 72 registers
// and three 58 KB CTAs per SM (88 KB left to the L1) beat eight warps at 64 registers and four CTAs (32 KB of L1).
#ifndef LARND_TILE_THREADS
#define LARND_TILE_THREADS 288
#endif
constexpr int TILE_THREADS = LARND_TILE_THREADS;
constexpr int NW = TILE_THREADS / 32;
constexpr int HS = 3 * KPT + 1;     // per-lane stride of the train buffer (odd: conflict-free)
constexpr int ES = KPT + 1;         // per-lane stride of the correction buffer (odd)
constexpr int MS = 5 * KPT + 1;     // per-run stride of the neighbour moments (odd)

struct SortArgs {
  const float* rec;
  int64_t n;
  const float* r0;
  const float* rm;
  const float* c0;
  const float* cm;
  const float* sr;
  const float* sc;
  int nt, L, Lp, ny_lut;
  int nticks;
  int nb, half2;
  int n_neigh, P;
  int nxp, nyp;
  int ntpl;
  RowLookup lk;
  const int32_t* counts;
  float* wfs;
  int64_t wstride;  // row stride of wfs in floats
  int v4ok;         // wfs is 16-byte aligned, wstride % 4 == 0 and >= nticks + 3: frames can be flushed with red.global.add.v4.f32
  int r0stride;     // row stride of the private row-0 copies
  const float* t0s; const float* tms;  // shifted response tables (larnd_lut)
  int lps, n0 /* nx * ny */, n0rows /* rows per shift copy of t0s: n0 + nb * nb */, nmrows;
  int skip_garbage;
  int4* runs_tmp;
  int4* runs;
  int* class_count;
  int* class_start;
  int* cursor;
  int4* tile_info;
  int* gcnt;  // [0] number of runs, [1] number of tiles, [GC_SPAN ..] per-span tile offsets, [GC_FWD ..] / [GC_BWD ..] tile counters
  int ncls;
  float* row0;  // [gridDim][nticks] per-CTA private copies of waveform row 0 (the garbage row every CTA adds to)
};

// ------------------------------------------------------------------------------------------------ run building
__global__ void __launch_bounds__(LARND_CHUNK)
k_build_runs(const __grid_constant__ SortArgs A) {
  __shared__ int s_ep[LARND_CHUNK], s_bx[LARND_CHUNK], s_by[LARND_CHUNK], s_idx[LARND_CHUNK], s_T0[LARND_CHUNK];
  __shared__ unsigned char s_fast[LARND_CHUNK], s_kh[LARND_CHUNK], s_head[LARND_CHUNK];
  __shared__ int s_wcnt[LARND_CHUNK / 32], s_base;
  if (A.counts[2] != 0) return;
  const int t = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * LARND_CHUNK;
  const int ns = (int)min((int64_t)LARND_CHUNK, A.n - base);
  const int* irec = reinterpret_cast<const int*>(A.rec);
  if (t < ns) {
    const int64_t s = base + t;
    s_ep[t] = irec[(int64_t)LARND_I_EP * A.n + s];
    s_bx[t] = irec[(int64_t)LARND_I_BX * A.n + s];
    s_by[t] = irec[(int64_t)LARND_I_BY * A.n + s];
    s_idx[t] = irec[(int64_t)LARND_I_IDX * A.n + s];
    const int T0 = irec[(int64_t)LARND_I_T0 * A.n + s];
    s_T0[t] = T0;
    s_fast[t] = seg_is_fast(T0, A.L, A.nticks);
  }
  s_head[t] = 0;
  // segments left to the chunk kernels' boundary pass: counted here so that pass can return at once when there are none
  const int nslow = __syncthreads_count(t < ns && !s_fast[t]);
  if (t == 0 && nslow) atomicAdd(A.gcnt + GC_SLOW, nslow);
  bool kh = false;
  if (t < ns) {
    kh = t == 0 || s_ep[t] != s_ep[t - 1] || s_bx[t] != s_bx[t - 1] || s_by[t] != s_by[t - 1] || s_idx[t] != s_idx[t - 1] ||
         s_fast[t] != s_fast[t - 1];
    s_kh[t] = kh;
  }
  __syncthreads();
  // the head of every key-run walks it once and cuts it greedily on the tick span / length limits
  if (kh && s_fast[t]) {
    int tmin = s_T0[t], tmax = tmin, start = t;
    s_head[t] = 1;
    for (int u = t + 1; u < ns && !s_kh[u]; ++u) {
      const int T0 = s_T0[u];
      if (max(tmax, T0) - min(tmin, T0) > SPAN_MAX_S || u - start >= MAXLEN) {
        s_head[u] = 1;
        start = u;
        tmin = tmax = T0;
      } else {
        tmin = min(tmin, T0);
        tmax = max(tmax, T0);
      }
    }
  }
  __syncthreads();
  const bool head = t < ns && s_head[t];
  const unsigned bal = __ballot_sync(0xffffffffu, head);
  const int lane = t & 31, wid = t >> 5;
  if (lane == 0) s_wcnt[wid] = __popc(bal);
  __syncthreads();
  int before = __popc(bal & ((1u << lane) - 1u));
  int total = 0;
#pragma unroll
  for (int w = 0; w < LARND_CHUNK / 32; ++w) {
    if (w < wid) before += s_wcnt[w];
    total += s_wcnt[w];
  }
  if (t == 0) s_base = total > 0 ? atomicAdd(A.gcnt, total) : 0;
  __syncthreads();
  if (head) {
    int tmin = s_T0[t], tmax = tmin, len = 1;
    for (int u = t + 1; u < ns && !s_kh[u] && !s_head[u]; ++u) {
      tmin = min(tmin, s_T0[u]);
      tmax = max(tmax, s_T0[u]);
      ++len;
    }
    const int nb = A.nb;
    const int bxm = s_bx[t] - floordiv_i(s_bx[t], nb) * nb, bym = s_by[t] - floordiv_i(s_by[t], nb) * nb;
    // span-major key: the tiles of every tick span form one contiguous range of the tile table
    const int cls = (tmax - tmin) * (A.ncls / (SPAN_MAX_S + 1)) + (s_idx[t] * nb + bxm) * nb + bym;
    A.runs_tmp[s_base + before] = make_int4((int)(base + t), len | ((tmax - tmin) << 16), tmin, cls);
    atomicAdd(A.class_count + 4 * cls + ((tmin - 1) & 3), 1);
  }
}

// exclusive scans over the class histogram: run offsets and tile offsets; then the tile table.  Two launches over blocks of
// 1024 classes: per-block totals (runs, tiles), then every block adds the totals in front of it to its in-block scans (the
// single looped CTA this replaces took ~0.12 ms for the 50 k classes of a 100-template bank).
constexpr int GC_BSUM = 64;         // gcnt[GC_BSUM + 2 b], [.. + 1]: runs / tiles of class block b (the counter area holds 64 + 2 * 128 ints)
constexpr int CLASS_BLOCKS_MAX = 128;

__global__ void __launch_bounds__(1024)
k_class_sums(const __grid_constant__ SortArgs A) {
  __shared__ int s_w[2][32];
  if (A.counts[2] != 0) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 1024 + threadIdx.x;
  const int4 sub = c < A.ncls ? reinterpret_cast<const int4*>(A.class_count)[c] : make_int4(0, 0, 0, 0);
  const int cnt = sub.x + sub.y + sub.z + sub.w;
  int v[2] = {cnt, (cnt + TR - 1) / TR};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) s_w[k][wid] = v[k];
  }
  __syncthreads();
  if (wid < 2) {
    int t = s_w[wid][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) A.gcnt[GC_BSUM + 2 * blockIdx.x + wid] = t;
  }
}

__global__ void __launch_bounds__(1024)
k_class_scan(const __grid_constant__ SortArgs A) {
  __shared__ int s_w[2][32];
  __shared__ int carry[2];
  if (A.counts[2] != 0) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (wid < 2) {  // totals of the class blocks in front of this one (<= CLASS_BLOCKS_MAX values per quantity)
    int t = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) t += A.gcnt[GC_BSUM + 2 * b + wid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) carry[wid] = t;
  }
  __syncthreads();
  {
    const int c = blockIdx.x * 1024 + threadIdx.x;
    const int4 sub = c < A.ncls ? reinterpret_cast<const int4*>(A.class_count)[c] : make_int4(0, 0, 0, 0);
    const int cnt = sub.x + sub.y + sub.z + sub.w;
    int v[2] = {cnt, (cnt + TR - 1) / TR};
    int inc[2] = {v[0], v[1]};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc[k], o);
        if (lane >= o) inc[k] += u;
      }
      if (lane == 31) s_w[k][wid] = inc[k];
    }
    __syncthreads();
    if (wid < 2) {
      int w = s_w[wid][lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      s_w[wid][lane] = w;
    }
    __syncthreads();
    int ex[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) ex[k] = carry[k] + (wid > 0 ? s_w[k][wid - 1] : 0) + inc[k] - v[k];
    if (c < A.ncls) {
      if (c % (A.ncls / (SPAN_MAX_S + 1)) == 0) A.gcnt[GC_SPAN + c / (A.ncls / (SPAN_MAX_S + 1))] = ex[1];
      A.class_start[c] = ex[0];
      reinterpret_cast<int4*>(A.cursor)[c] = make_int4(ex[0], ex[0] + sub.x, ex[0] + sub.x + sub.y, ex[0] + sub.x + sub.y + sub.z);
      for (int i = 0; i < v[1]; ++i)  // tiles of this class
        A.tile_info[ex[1] + i] = make_int4(c, ex[0] + i * TR, min(TR, cnt - i * TR), 0);
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
      const int ntiles = carry[1] + s_w[1][31];
      A.gcnt[1] = ntiles;
      A.gcnt[GC_SPAN + SPAN_MAX_S + 1] = ntiles;
    }
  }
}

__global__ void k_scatter_runs(const __grid_constant__ SortArgs A) {
  if (A.counts[2] != 0) return;
  const int nruns = A.gcnt[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nruns; i += gridDim.x * blockDim.x) {
    const int4 e = A.runs_tmp[i];
    const int pos = atomicAdd(A.cursor + 4 * e.w + ((e.z - 1) & 3), 1);
    A.runs[pos] = e;
  }
}


// fills the fields every sorted kernel needs and (re)builds the sorted run table + tile table on `st`
// reuse: LARND_FLAG_REUSE_RUNS was given — when the library's record says this workspace still holds the tables of these
// records, only the tile counters are reset (the three kernels below cost 0.36 ms per 10 M segments).
inline int sorted_fill_and_build(SortArgs& A, int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                 int32_t npix_capacity, const int32_t* counts, cudaStream_t st, bool reuse = false) {
  if (n >= (int64_t)1 << 31) { larnd_set_error("sorted accumulate: n must be < 2^31"); return LARND_E_ARG; }
  A.rec = ws.rec; A.n = n;
  A.r0 = lut->r0; A.rm = lut->rm; A.c0 = lut->c0; A.cm = lut->cm; A.sr = lut->sr; A.sc = lut->sc;
  A.nt = lut->nt; A.L = lut->L; A.Lp = lut->Lp; A.ny_lut = lut->ny;
  A.t0s = lut->t0s; A.tms = lut->tms; A.lps = lut->lps;
  A.n0 = lut->nx * lut->ny; A.n0rows = A.n0 + p.nb_sampling_bins_per_pixel * p.nb_sampling_bins_per_pixel; A.nmrows = lut->ntpl * 25;
  A.nticks = p.n_ticks;
  A.nb = p.nb_sampling_bins_per_pixel;
  A.half2 = 2 * (A.nb / 2) - 1;
  A.n_neigh = p.number_pix_neighbors;
  A.P = 2 * A.n_neigh + 1;
  A.nxp = p.n_pixels_x; A.nyp = p.n_pixels_y;
  A.ntpl = lut->ntpl;
  A.lk.bitmap = ws.bitmap; A.lk.wprefix = ws.wprefix; A.lk.n_words = ws.n_words; A.lk.pid_offset = ws.pid_offset;
  A.lk.n_unique = 0; A.lk.n_neg = 0; A.lk.npix = npix_capacity;
  A.counts = counts;
  A.runs_tmp = reinterpret_cast<int4*>(ws.runs_tmp);
  A.runs = reinterpret_cast<int4*>(ws.runs);
  A.class_count = ws.class_count; A.class_start = ws.class_start; A.cursor = ws.cursor;
  A.tile_info = reinterpret_cast<int4*>(ws.tile_info);
  A.gcnt = ws.gcnt;
  A.ncls = lut->ntpl * A.nb * A.nb * (SPAN_MAX_S + 1);
  A.row0 = ws.row0;
  if (reuse && larnd_runs_cache_valid(ws.rec, n, lut, p.n_ticks)) {
    LARND_CUDA(cudaMemsetAsync(ws.gcnt + GC_FWD, 0, 16 * sizeof(int), st));  // tile counters of the forward and backward launches
    return LARND_OK;
  }
  LARND_CUDA(cudaMemsetAsync(ws.class_count, 0, (size_t)A.ncls * 4 * sizeof(int), st));
  LARND_CUDA(cudaMemsetAsync(ws.gcnt, 0, 256, st));
  const int64_t chunks = (n + LARND_CHUNK - 1) / LARND_CHUNK;
  k_build_runs<<<(unsigned)chunks, LARND_CHUNK, 0, st>>>(A);
  LARND_LAUNCH_CHECK("k_build_runs");
  const int class_blocks = (A.ncls + 1023) / 1024;
  if (class_blocks > CLASS_BLOCKS_MAX) { larnd_set_error("sorted accumulate: too many response classes"); return LARND_E_ARG; }
  k_class_sums<<<class_blocks, 1024, 0, st>>>(A);
  LARND_LAUNCH_CHECK("k_class_sums");
  k_class_scan<<<class_blocks, 1024, 0, st>>>(A);
  LARND_LAUNCH_CHECK("k_class_scan");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  k_scatter_runs<<<(unsigned)blocks, 256, 0, st>>>(A);
  LARND_LAUNCH_CHECK("k_scatter_runs");
  larnd_runs_cache_set(ws.rec, n, lut, p.n_ticks);
  return LARND_OK;
}

// How the tile table is split over kernel variants: by default the tiles of spans 0-2 go to the variant that holds 4 impulse
// positions in registers (more CTAs per SM) and the rest to the KPT-position variant; LARND_FLAG_NO_SPLIT sends everything to
// the latter.
inline int sorted_split_mode(int flags) { return (flags & LARND_FLAG_NO_SPLIT) ? 0 : 1; }

inline int sorted_grid(int per_sm, int cap) {
  int nsm = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  int g = nsm * per_sm;
  return g < cap ? g : cap;
}

}  // namespace
