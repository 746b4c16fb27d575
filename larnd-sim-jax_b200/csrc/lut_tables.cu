// Device tables derived once from the response template bank (reference: consts_jax.py:387-449 builds the
// bank; sim_jax.py:228 recomputes jnp.cumsum(response_template) over the whole 1.58 GB bank on EVERY call —
// here the running sums are hoisted to LUT-load time and restricted to the rows the path can touch:
// template 0 for every (ci,cj) bin (neighbours, sim_jax.py:221-222,250-255) and all templates for the 5x5
// collecting bins (main pixels: currents_idx in [0,4], sim_jax.py:435,184-190,236-241)).
#include "larnd_common.cuh"

namespace {

// compact rows: [0, 0, R[nt-L .. nt-1], 0, 0]
__global__ void k_compact_rows(const float* __restrict__ bank, int nx, int ny, int nt, int L, int Lp, int ntpl,
                               float* __restrict__ r0, float* __restrict__ rm) {
  const int n0 = nx * ny;
  const int nrows = n0 + ntpl * 25;
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const float* src;
    float* dst;
    if (r < n0) {
      src = bank + (int64_t)r * nt;
      dst = r0 + (int64_t)r * Lp;
    } else {
      int m = r - n0, t = m / 25, b = m % 25, ci = b / 5, cj = b % 5;
      src = bank + (((int64_t)t * nx + ci) * ny + cj) * nt;
      dst = rm + (int64_t)m * Lp;
    }
    for (int k = threadIdx.x; k < Lp; k += blockDim.x) {
      int kk = k - 2;
      dst[k] = (kk >= 0 && kk < L) ? src[nt - L + kk] : 0.0f;
    }
  }
}

// Shifted copies for the lane <-> 4-ticks tile kernels (accumulate_sorted.cu): dst[s][r][i] = R_r[i - s - 8] (0 outside
// [0, L)), s = 0..3, so that a lane whose four output ticks start at frame position 4x reads its response samples with
// aligned 128-bit loads whatever the run's start tick modulo 4 is.  src rows are the compact rows (sample k at k + 2).
// dst holds dst_rows rows per shift copy; the nrows source rows land at rows dst_row0 .. dst_row0 + nrows - 1 of every copy.
__global__ void k_shift_rows(const float* __restrict__ src, int nrows, int L, int Lp, int lps, float* __restrict__ dst, int dst_rows,
                             int dst_row0) {
  const int64_t total = (int64_t)4 * nrows * lps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % lps);
    const int64_t rr = e / lps;
    const int r = (int)(rr % nrows), s = (int)(rr / nrows);
    const int k = i - s - 8;
    dst[((int64_t)s * dst_rows + dst_row0 + r) * lps + i] = (k >= 0 && k < L) ? src[(int64_t)r * Lp + k + 2] : 0.0f;
  }
}

// float32 running sum along time, strictly left to right (bit-identical to a sequential cumsum)
__global__ void k_cumsum_rows(const float* __restrict__ bank, int nx, int ny, int nt, int ntpl,
                              float* __restrict__ c0, float* __restrict__ cm) {
  const int n0 = nx * ny;
  const int nrows = n0 + ntpl * 25;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const float* src;
  float* dst;
  if (r < n0) {
    src = bank + (int64_t)r * nt;
    dst = c0 + (int64_t)r * nt;
  } else {
    int m = r - n0, t = m / 25, b = m % 25, ci = b / 5, cj = b % 5;
    src = bank + (((int64_t)t * nx + ci) * ny + cj) * nt;
    dst = cm + (int64_t)m * nt;
  }
  float acc = 0.0f;
  for (int k = 0; k < nt; ++k) {
    acc = __fadd_rn(acc, src[k]);
    dst[k] = acc;
  }
}

// neighbourhood sums (see larnd_common.cuh): one block per in-pixel bin
__global__ void k_neighbour_sums(const float* __restrict__ r0, const float* __restrict__ c0, int ny, int nt, int Lp, int nb, int n,
                                 float* __restrict__ sr, float* __restrict__ sc) {
  const int bxm = blockIdx.x / nb, bym = blockIdx.x % nb;
  const int half2 = 2 * (nb / 2) - 1;
  for (int k = threadIdx.x; k < Lp + nt; k += blockDim.x) {
    float acc = 0.0f;
    for (int dx = -n; dx <= n; ++dx) {
      const int ci = abs(2 * bxm - half2 - 2 * nb * dx) >> 1;
      for (int dy = -n; dy <= n; ++dy) {
        const int cj = abs(2 * bym - half2 - 2 * nb * dy) >> 1;
        const int bin = ci * ny + cj;
        acc += (k < Lp) ? r0[(int64_t)bin * Lp + k] : c0[(int64_t)bin * nt + (k - Lp)];
      }
    }
    if (k < Lp) sr[(int64_t)blockIdx.x * Lp + k] = acc;
    else sc[(int64_t)blockIdx.x * nt + (k - Lp)] = acc;
  }
}

}  // namespace

// Neighbourhood-sum rows for (nb, n): explicit, mutating call made once after larnd_lut_create (never from a per-batch
// entry point: those take a const handle and only CHECK that the tables match their parameters).
extern "C" int larnd_lut_prepare_neighbours(larnd_lut_t* lut, int nb, int n, void* stream) {
  if (!lut || nb < 2 || n < 0 || n > 7) { larnd_set_error("larnd_lut_prepare_neighbours: invalid argument"); return LARND_E_ARG; }
  if (lut->sr && lut->sum_nb == nb && lut->sum_n == n) return LARND_OK;
  if (nb * n + nb / 2 > lut->nx || nb * n + nb / 2 > lut->ny) {
    larnd_set_error("number_pix_neighbors=%d needs %d response bins per axis, LUT has %dx%d", n, nb * n + nb / 2, lut->nx, lut->ny);
    return LARND_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (lut->sr && lut->sum_nb != nb) {
    LARND_CUDA(cudaStreamSynchronize(st));
    cudaFree(lut->sr); cudaFree(lut->sc); cudaFree(lut->t0s);
    lut->sr = lut->sc = lut->t0s = nullptr;
  }
  if (!lut->sr) {
    LARND_CUDA(cudaMalloc(&lut->sr, (size_t)nb * nb * lut->Lp * sizeof(float)));
    LARND_CUDA(cudaMalloc(&lut->sc, (size_t)nb * nb * lut->nt * sizeof(float)));
    if (lut->nsv) LARND_CUDA(cudaMalloc(&lut->t0s, (size_t)4 * (lut->nx * lut->ny + nb * nb) * lut->lps * sizeof(float)));
  }
  k_neighbour_sums<<<nb * nb, 256, 0, st>>>(lut->r0, lut->c0, lut->ny, lut->nt, lut->Lp, nb, n, lut->sr, lut->sc);
  LARND_LAUNCH_CHECK("k_neighbour_sums");
  if (lut->nsv) {  // shifted template-0 rows followed by the shifted neighbourhood-sum rows, one table per shift
    const int n0 = lut->nx * lut->ny;
    k_shift_rows<<<148 * 4, 256, 0, st>>>(lut->r0, n0, lut->L, lut->Lp, lut->lps, lut->t0s, n0 + nb * nb, 0);
    k_shift_rows<<<64, 256, 0, st>>>(lut->sr, nb * nb, lut->L, lut->Lp, lut->lps, lut->t0s, n0 + nb * nb, n0);
    LARND_LAUNCH_CHECK("k_shift_rows(t0s)");
  }
  lut->sum_nb = nb;
  lut->sum_n = n;
  return LARND_OK;
}

int larnd_lut_check_neighbours(const larnd_lut* lut, int nb, int n) {
  if (lut->sr && lut->sum_nb == nb && lut->sum_n == n) return LARND_OK;
  larnd_set_error("LUT neighbour tables were not prepared for nb_sampling_bins_per_pixel=%d, number_pix_neighbors=%d: call "
                  "larnd_lut_prepare_neighbours() after larnd_lut_create()", nb, n);
  return LARND_E_ARG;
}

extern "C" int larnd_lut_create(const float* bank_d, int n_templates, int nx, int ny, int nt, int signal_length,
                                void* stream, larnd_lut_t** out) {
  if (!bank_d || !out || n_templates < 3 || n_templates > LARND_MAX_TEMPLATES || nx < 5 || ny < 5 || nt < 2 ||
      signal_length < 1 || signal_length > nt) {
    larnd_set_error("larnd_lut_create: invalid shape (ntpl=%d nx=%d ny=%d nt=%d L=%d)", n_templates, nx, ny, nt,
                    signal_length);
    return LARND_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  larnd_lut* lut = new larnd_lut();
  lut->ntpl = n_templates; lut->nx = nx; lut->ny = ny; lut->nt = nt; lut->L = signal_length;
  lut->Lp = signal_length + LARND_OFF_PAD;
  size_t n0 = (size_t)nx * ny, nm = (size_t)n_templates * 25;
  lut->r0 = lut->rm = lut->c0 = lut->cm = nullptr;
  lut->sr = lut->sc = nullptr;
  lut->t0s = lut->tms = nullptr;
  lut->sum_nb = lut->sum_n = -1;
  // frames of the tile kernels: run window (L + up to 6 impulse positions) + alignment shift (<= 3) in 128-tick slots
  const int frame = signal_length + 6 + 3;
  lut->nsv = frame <= 128 ? 1 : (frame <= 256 ? 2 : 0);
  lut->lps = 128 * lut->nsv + 8;
  int rc = LARND_OK;
  auto fail = [&](int code) { larnd_lut_destroy(lut); return code; };
  if ((rc = larnd_check_cuda(cudaMalloc(&lut->r0, n0 * lut->Lp * sizeof(float)), "cudaMalloc r0"))) return fail(rc);
  if ((rc = larnd_check_cuda(cudaMalloc(&lut->rm, nm * lut->Lp * sizeof(float)), "cudaMalloc rm"))) return fail(rc);
  if ((rc = larnd_check_cuda(cudaMalloc(&lut->c0, n0 * nt * sizeof(float)), "cudaMalloc c0"))) return fail(rc);
  if ((rc = larnd_check_cuda(cudaMalloc(&lut->cm, nm * nt * sizeof(float)), "cudaMalloc cm"))) return fail(rc);
  int nrows = (int)(n0 + nm);
  k_compact_rows<<<min(nrows, 148 * 8), 128, 0, st>>>(bank_d, nx, ny, nt, signal_length, lut->Lp, n_templates, lut->r0, lut->rm);
  if ((rc = larnd_check_cuda(cudaGetLastError(), "k_compact_rows"))) return fail(rc);
  k_cumsum_rows<<<(nrows + 63) / 64, 64, 0, st>>>(bank_d, nx, ny, nt, n_templates, lut->c0, lut->cm);
  if ((rc = larnd_check_cuda(cudaGetLastError(), "k_cumsum_rows"))) return fail(rc);
  if (lut->nsv) {
    if ((rc = larnd_check_cuda(cudaMalloc(&lut->tms, 4 * nm * lut->lps * sizeof(float)), "cudaMalloc tms"))) return fail(rc);
    k_shift_rows<<<148 * 4, 256, 0, st>>>(lut->rm, (int)nm, signal_length, lut->Lp, lut->lps, lut->tms, (int)nm, 0);
    if ((rc = larnd_check_cuda(cudaGetLastError(), "k_shift_rows"))) return fail(rc);
  }
  *out = lut;
  return LARND_OK;
}

extern "C" void larnd_lut_destroy(larnd_lut_t* lut) {
  if (!lut) return;
  cudaFree(lut->r0);
  cudaFree(lut->rm);
  cudaFree(lut->c0);
  cudaFree(lut->cm);
  cudaFree(lut->sr);
  cudaFree(lut->sc);
  cudaFree(lut->t0s);
  cudaFree(lut->tms);
  delete lut;
}

// ---- template bank builder (reference: load_lut, consts_jax.py:427-447) ---------------------------------------------
// bank[t][row][k] = sum_m gauss[t][m] * response[row][k + half - m]   ('same' convolution, zero padded), t >= 1;
// bank[0] = response.  One CTA per (row, block of templates): the response row sits in shared memory, every thread
// produces one output sample per template; 100 x 2025 x 1950 x 41 multiply-adds and one 1.58 GB write in total.
namespace {
constexpr int BANK_TPL_PER_CTA = 10;
__global__ void __launch_bounds__(256)
k_build_bank(const float* __restrict__ response, int nrows, int nt, const float* __restrict__ gauss, int ntpl, int taps,
             float* __restrict__ bank) {
  extern __shared__ float s_row[];  // [half zeros | response row | half zeros], then the Gaussians of this CTA's templates
  const int half = taps / 2;
  const int row = blockIdx.x;
  const int t_lo = blockIdx.y * BANK_TPL_PER_CTA, t_hi = min(ntpl, t_lo + BANK_TPL_PER_CTA);
  float* s_g = s_row + nt + 2 * half;
  for (int k = threadIdx.x; k < nt + 2 * half; k += blockDim.x) {
    const int kk = k - half;
    s_row[k] = (kk >= 0 && kk < nt) ? response[(int64_t)row * nt + kk] : 0.0f;
  }
  for (int k = threadIdx.x; k < (t_hi - t_lo) * taps; k += blockDim.x) s_g[k] = gauss[(int64_t)t_lo * taps + k];
  __syncthreads();
  for (int t = t_lo; t < t_hi; ++t) {
    float* dst = bank + ((int64_t)t * nrows + row) * nt;
    const float* g = s_g + (t - t_lo) * taps;
    for (int k = threadIdx.x; k < nt; k += blockDim.x) {
      float acc;
      if (t == 0) {
        acc = s_row[k + half];
      } else {
        acc = 0.0f;
        for (int m = 0; m < taps; ++m) acc = fmaf(g[m], s_row[k + 2 * half - m], acc);  // response[k + half - m]
      }
      dst[k] = acc;
    }
  }
}
}  // namespace

extern "C" int larnd_build_bank(const float* response_d, int nx, int ny, int nt, const float* gauss_d, int n_templates, int taps,
                                float* bank_d, void* stream) {
  if (!response_d || !gauss_d || !bank_d || nx < 1 || ny < 1 || nt < 1 || n_templates < 1 || taps < 1 || (taps & 1) == 0) {
    larnd_set_error("larnd_build_bank: invalid argument (taps must be odd)");
    return LARND_E_ARG;
  }
  const size_t smem = (size_t)(nt + 2 * (taps / 2) + BANK_TPL_PER_CTA * taps) * sizeof(float);
  if (smem > 200 * 1024) { larnd_set_error("larnd_build_bank: response row too long for shared memory"); return LARND_E_ARG; }
  LARND_CUDA(cudaFuncSetAttribute(k_build_bank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(nx * ny), (unsigned)((n_templates + BANK_TPL_PER_CTA - 1) / BANK_TPL_PER_CTA));
  k_build_bank<<<grid, 256, smem, (cudaStream_t)stream>>>(response_d, nx * ny, nt, gauss_d, n_templates, taps, bank_d);
  LARND_LAUNCH_CHECK("k_build_bank");
  return LARND_OK;
}
