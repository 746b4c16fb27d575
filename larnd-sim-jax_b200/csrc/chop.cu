// Device-side chop_tracks (reference: optimize/dataio.py:63-106) — SURVEY.md §8f.2.
//
// The reference subdivides every raw segment into ceil(length / precision) pieces with a Python loop per row
// (np.vstack([split_track(...) for i in range(N)])) on the host and ships 104 B per *chopped* segment to the device.
// Here the raw rows (100-700x fewer) are uploaded and expanded on the GPU:
//   k_chop_count    one thread per raw row: float32 length, number of pieces (numpy's float32 arithmetic)
//   k_chop_scan     exclusive scan of the piece counts (single CTA; a few 10^5 raw rows at most)
//   k_chop_expand   one warp per raw row, lanes <-> columns: every output row is one coalesced 104-byte store
// Arithmetic follows numpy's promotion rules of the reference expressions: `steps*precision*direction` is
// float64 (int64 array * Python float * float32 scalar), rounded to float32 on assignment; dE of the inner pieces is
// float32 (float32 array * weak Python float / float32 scalar); the last piece is computed in float64 and rounded.
#include "larnd_common.cuh"
#include "chop_math.cuh"

namespace {

__global__ void k_chop_count(const float* __restrict__ raw, int64_t m, const __grid_constant__ larnd_chop_columns_t c,
                             float prec32, int64_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const ChopGeom g = chop_geom(raw + i * c.ncols, c);
  counts[i] = chop_nsteps(g.len, prec32);
}

// in-place exclusive scan of counts[0..m) (int64), total written to counts[m]
__global__ void __launch_bounds__(1024) k_chop_scan(int64_t* __restrict__ counts, int64_t m) {
  __shared__ long long wsum[32];
  __shared__ long long carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < m; b0 += 1024) {
    const int64_t i = b0 + threadIdx.x;
    const long long v = i < m ? counts[i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      long long w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const long long ex = carry + (wid > 0 ? wsum[wid - 1] : 0) + inc - v;
    if (i < m) counts[i] = ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[m] = carry;
}

constexpr int CHOP_ROWS = 32;  // output rows per warp task

// One warp per chunk of CHOP_ROWS consecutive OUTPUT rows (lanes <-> columns, every output row one coalesced store), the raw
// row of the chunk's first piece found by binary search in the prefix table: the work is balanced over the pieces, not over
// the raw rows (a raw row expands into ~2 000 pieces at 0.01 cm, and a batch has only a few thousand raw rows).
__global__ void __launch_bounds__(256)
k_chop_expand(const float* __restrict__ raw, int64_t m, const __grid_constant__ larnd_chop_columns_t c, double precision,
              float prec32, const int64_t* __restrict__ offsets, float* __restrict__ out, int64_t capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int ncols = c.ncols;
  const int64_t total = offsets[m];
  if (total > capacity) return;  // caller checks counts[m] against its capacity
  const int64_t nchunks = (total + CHOP_ROWS - 1) / CHOP_ROWS;
  for (int64_t chunk = warp; chunk < nchunks; chunk += nwarps) {
    int64_t o = chunk * CHOP_ROWS;
    const int64_t o_end = min(total, o + CHOP_ROWS);
    int64_t lo = 0, hi = m - 1;  // largest i with offsets[i] <= o (every raw row has at least one piece)
    while (lo < hi) {
      const int64_t mid = (lo + hi + 1) >> 1;
      if (offsets[mid] <= o) lo = mid; else hi = mid - 1;
    }
    for (int64_t i = lo; o < o_end; ++i) {
      const float* tr = raw + i * ncols;
      const ChopGeom g = chop_geom(tr, c);
      const long long n = chop_nsteps(g.len, prec32);
      const int64_t o0 = offsets[i];
      const long long k0 = o - o0, k1 = min((long long)n, k0 + (long long)(o_end - o));
      // which role this lane's column plays (lanes >= ncols idle; rows wider than 32 columns loop)
      for (int col0 = 0; col0 < ncols; col0 += 32) {
        const int col = col0 + lane;
        if (col >= ncols) continue;
        const float base = tr[col];
        int axis = -1, kind = 0;  // kind: 1 start, 2 end, 3 mid, 4 dx, 5 dE
        if (col == c.x_start) { axis = 0; kind = 1; } else if (col == c.y_start) { axis = 1; kind = 1; } else if (col == c.z_start) { axis = 2; kind = 1; }
        else if (col == c.x_end) { axis = 0; kind = 2; } else if (col == c.y_end) { axis = 1; kind = 2; } else if (col == c.z_end) { axis = 2; kind = 2; }
        else if (col == c.x) { axis = 0; kind = 3; } else if (col == c.y) { axis = 1; kind = 3; } else if (col == c.z) { axis = 2; kind = 3; }
        else if (col == c.dx) kind = 4;
        else if (col == c.dE) kind = 5;
        const int cs = axis == 0 ? c.x_start : (axis == 1 ? c.y_start : c.z_start);
        const int ce = axis == 0 ? c.x_end : (axis == 1 ? c.y_end : c.z_end);
        const double s0 = axis >= 0 ? (double)tr[cs] : 0.0, d = axis >= 0 ? (double)g.dir[axis] : 0.0;
        const float e_last = axis >= 0 ? tr[ce] : 0.0f;
        const float len_eps = __fadd_rn(g.len, 1e-10f);            // np.float32 + weak Python float
        const float dE_in = __fdiv_rn(__fmul_rn(base, prec32), len_eps);
        for (long long k = k0; k < k1; ++k) {
          const bool last = k == n - 1;
          float v = base;
          if (kind == 1 || kind == 2 || kind == 3) {
            const float vs = (float)(s0 + __dmul_rn(__dmul_rn((double)k, precision), d));
            const float ve = last ? e_last : (float)(s0 + __dmul_rn(__dmul_rn(precision, (double)(k + 1)), d));
            v = kind == 1 ? vs : (kind == 2 ? ve : __fmul_rn(0.5f, __fadd_rn(vs, ve)));
          } else if (kind == 4) {
            v = last ? (float)((double)g.len - __dmul_rn(precision, (double)(n - 1))) : prec32;
          } else if (kind == 5) {
            v = last ? (float)__dmul_rn((double)base, 1.0 - __ddiv_rn(__dmul_rn(precision, (double)(n - 1)), (double)len_eps)) : dE_in;
          }
          out[(o0 + k) * ncols + col] = v;
        }
      }
      o += k1 - k0;
    }
  }
}

}  // namespace

extern "C" int larnd_chop_count(const float* raw_d, int64_t m, const larnd_chop_columns_t* cols, double precision,
                                int64_t* offsets_d, void* stream) {
  if ((!raw_d && m > 0) || !cols || !offsets_d || m < 0 || !(precision > 0)) { larnd_set_error("larnd_chop_count: bad argument"); return LARND_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (m > 0) {
    k_chop_count<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(raw_d, m, *cols, (float)precision, offsets_d);
    LARND_LAUNCH_CHECK("k_chop_count");
  }
  k_chop_scan<<<1, 1024, 0, st>>>(offsets_d, m);
  LARND_LAUNCH_CHECK("k_chop_scan");
  return LARND_OK;
}

extern "C" int larnd_chop_tracks(const float* raw_d, int64_t m, const larnd_chop_columns_t* cols, double precision,
                                 const int64_t* offsets_d, float* out_d, int64_t capacity, void* stream) {
  if ((!raw_d && m > 0) || !cols || !offsets_d || (!out_d && capacity > 0) || m < 0 || !(precision > 0)) {
    larnd_set_error("larnd_chop_tracks: bad argument");
    return LARND_E_ARG;
  }
  if (m == 0) return LARND_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t blocks = 148 * 16;  // persistent warps striding over the chunks of output rows
  k_chop_expand<<<(unsigned)blocks, 256, 0, st>>>(raw_d, m, *cols, precision, (float)precision, offsets_d, out_d, capacity);
  LARND_LAUNCH_CHECK("k_chop_expand");
  return LARND_OK;
}

// ---- batch assembly in front of the chop (TracksDataset.__getitem__ / pad_batch, optimize/dataio.py:340-406) -----------
// The reference builds every batch on the host: gather the rows of the batch's trajectories from the structured file
// array, convert to float32, remap global event ids to batch-local ones (remap_event_ids_to_local :47-61), chop, pad
// with invalid rows.  Here the file's rows live on the device once; a batch is a list of row indices + local ids.
namespace {
__global__ void k_batch_gather(const float* __restrict__ raw, int ncols, const int64_t* __restrict__ rows, const int32_t* __restrict__ local_event,
                               int64_t m, int event_col, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * ncols) return;
  const int64_t r = i / ncols;
  const int c = (int)(i - r * ncols);
  out[i] = (c == event_col) ? (float)local_event[r] : raw[rows[r] * ncols + c];
}

__global__ void k_pad_rows(float* __restrict__ batch, int ncols, const int64_t* __restrict__ n_valid, int64_t capacity,
                           const __grid_constant__ larnd_pad_columns_t pc) {
  const int64_t first = min(max(*n_valid, (int64_t)0), capacity);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + first * ncols;
  if (i >= capacity * ncols) return;
  const int c = (int)(i % ncols);
  batch[i] = (c == pc.eventID || c == pc.trackID || c == pc.pixel_plane) ? -1.0f : 0.0f;
}
}  // namespace

extern "C" int larnd_batch_gather(const float* raw_d, int32_t ncols, const int64_t* rows_d, const int32_t* local_event_d, int64_t m,
                                  int32_t event_col, float* out_d, void* stream) {
  if (m < 0 || ncols < 1 || event_col < 0 || event_col >= ncols || (m > 0 && (!raw_d || !rows_d || !local_event_d || !out_d))) {
    larnd_set_error("larnd_batch_gather: bad argument");
    return LARND_E_ARG;
  }
  if (m == 0) return LARND_OK;
  const int64_t n = m * ncols;
  k_batch_gather<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw_d, ncols, rows_d, local_event_d, m, event_col, out_d);
  LARND_LAUNCH_CHECK("k_batch_gather");
  return LARND_OK;
}

extern "C" int larnd_pad_rows(float* batch_d, int32_t ncols, const int64_t* n_valid_d, int64_t capacity, const larnd_pad_columns_t* cols,
                              void* stream) {
  if (capacity < 0 || ncols < 1 || !cols || !n_valid_d || (capacity > 0 && !batch_d)) { larnd_set_error("larnd_pad_rows: bad argument"); return LARND_E_ARG; }
  if (capacity == 0) return LARND_OK;
  const int64_t n = capacity * ncols;   // the grid covers the worst case (n_valid = 0); threads beyond the tail exit
  k_pad_rows<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(batch_d, ncols, n_valid_d, capacity, *cols);
  LARND_LAUNCH_CHECK("k_pad_rows");
  return LARND_OK;
}
