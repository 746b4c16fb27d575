// C-ABI entry points (include/larnd_b200.h): argument checking, workspace carving, kernel sequencing.
#include <stdarg.h>
#include <string.h>

#include "larnd_common.cuh"
#include <mutex>

static thread_local char g_err[512] = "";

void larnd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int larnd_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return LARND_OK;
  larnd_set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return LARND_E_CUDA;
}

unsigned long long g_launch_count = 0;
extern "C" uint64_t larnd_launch_count(void) { return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED); }

bool g_prof_on = false;
ProfSlot g_prof[LARND_PROF_SLOTS];

extern "C" int larnd_profile_enable(int on) {
  static bool created = false;
  if (on && !created) {
    for (int i = 0; i < LARND_PROF_SLOTS; ++i) {
      LARND_CUDA(cudaEventCreate(&g_prof[i].start));
      LARND_CUDA(cudaEventCreate(&g_prof[i].stop));
    }
    created = true;
  }
  for (int i = 0; i < LARND_PROF_SLOTS; ++i) g_prof[i].used = false;
  g_prof_on = on != 0;
  return LARND_OK;
}

extern "C" int larnd_profile_read(float* ms_out) {
  if (!ms_out) { larnd_set_error("larnd_profile_read: null argument"); return LARND_E_ARG; }
  for (int i = 0; i < LARND_PROF_SLOTS; ++i) {
    ms_out[i] = -1.0f;
    if (g_prof_on && g_prof[i].used) {
      LARND_CUDA(cudaEventSynchronize(g_prof[i].stop));
      LARND_CUDA(cudaEventElapsedTime(&ms_out[i], g_prof[i].start, g_prof[i].stop));
    }
  }
  return LARND_OK;
}

extern "C" const char* larnd_last_error(void) { return g_err; }
extern "C" int larnd_abi_version(void) { return LARND_ABI_VERSION; }

static int64_t bitmap_words(int32_t n_events, int32_t ntpc, int32_t nx, int32_t ny) {
  int64_t bits = ((int64_t)n_events + 1) * ntpc * nx * ny;  // events -1 .. n_events-1
  return (bits + 31) / 32;
}

extern "C" size_t larnd_workspace_bytes(int64_t n, int32_t n_events, int32_t ntpc, int32_t nx, int32_t ny) {
  if (n < 0 || n_events < 0 || ntpc < 1 || nx < 1 || ny < 1) return 0;
  int64_t nw = bitmap_words(n_events, ntpc, nx, ny);
  int64_t nsb = (nw + LARND_SCAN_WORDS_PER_BLOCK - 1) / LARND_SCAN_WORDS_PER_BLOCK;
  int64_t nchunks = (n + LARND_CHUNK - 1) / LARND_CHUNK + 1 + LARND_SMALL_CHUNK_SLOTS;
  size_t b = 0;
  b += align_up((size_t)LARND_NFIELDS * (size_t)(n > 0 ? n : 1) * sizeof(float), 256);
  b += align_up((size_t)nw * 4, 256) * 2;
  b += align_up((size_t)nsb * 4, 256);
  b += align_up((size_t)(nchunks + LARND_BWD_SORTED_SLOTS) * 16 * sizeof(float), 256);
  b += larnd_sorted_workspace_bytes(n);
  return b;
}

namespace {
struct RunsCacheEntry { const void* rec; int64_t n; const void* lut; int n_ticks; };
constexpr int RUNS_CACHE_N = 16;
RunsCacheEntry g_runs_cache[RUNS_CACHE_N];
int g_runs_cache_next = 0;
std::mutex g_runs_cache_mu;
}  // namespace

void larnd_runs_cache_drop(const void* rec) {
  std::lock_guard<std::mutex> lock(g_runs_cache_mu);
  for (auto& e : g_runs_cache)
    if (e.rec == rec) e.rec = nullptr;
}

void larnd_runs_cache_set(const void* rec, int64_t n, const void* lut, int n_ticks) {
  std::lock_guard<std::mutex> lock(g_runs_cache_mu);
  RunsCacheEntry* slot = nullptr;
  for (auto& e : g_runs_cache)
    if (e.rec == rec) slot = &e;
  if (!slot) { slot = &g_runs_cache[g_runs_cache_next]; g_runs_cache_next = (g_runs_cache_next + 1) % RUNS_CACHE_N; }
  *slot = RunsCacheEntry{rec, n, lut, n_ticks};
}

bool larnd_runs_cache_valid(const void* rec, int64_t n, const void* lut, int n_ticks) {
  std::lock_guard<std::mutex> lock(g_runs_cache_mu);
  for (auto& e : g_runs_cache)
    if (e.rec == rec && rec) return e.n == n && e.lut == lut && e.n_ticks == n_ticks;
  return false;
}

bool larnd_carve_workspace(void* base, size_t bytes, int64_t n, int32_t n_events, int32_t ntpc, int32_t nx, int32_t ny,
                           Workspace* ws) {
  if (!base || bytes < larnd_workspace_bytes(n, n_events, ntpc, nx, ny)) return false;
  char* p = reinterpret_cast<char*>(base);
  int64_t nw = bitmap_words(n_events, ntpc, nx, ny);
  ws->n = n;
  ws->n_words = nw;
  ws->n_scan_blocks = (nw + LARND_SCAN_WORDS_PER_BLOCK - 1) / LARND_SCAN_WORDS_PER_BLOCK;
  ws->n_chunks_max = (n + LARND_CHUNK - 1) / LARND_CHUNK + 1 + LARND_SMALL_CHUNK_SLOTS;
  ws->pid_offset = ntpc * nx * ny;
  ws->rec = reinterpret_cast<float*>(p);
  p += align_up((size_t)LARND_NFIELDS * (size_t)(n > 0 ? n : 1) * sizeof(float), 256);
  ws->bitmap = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)nw * 4, 256);
  ws->wprefix = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)nw * 4, 256);
  ws->bsums = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)ws->n_scan_blocks * 4, 256);
  ws->partials = reinterpret_cast<float*>(p);
  p += align_up((size_t)(ws->n_chunks_max + LARND_BWD_SORTED_SLOTS) * 16 * sizeof(float), 256);
  larnd_carve_sorted(p, n, ws);
  return true;
}

static int check_common(const larnd_params_t* p, const larnd_lut_t* lut, bool need_lut) {
  if (!p) { larnd_set_error("params is null"); return LARND_E_ARG; }
  if (p->n_tpc < 1 || p->n_tpc > LARND_MAX_TPC) { larnd_set_error("n_tpc=%d unsupported", p->n_tpc); return LARND_E_ARG; }
  if (p->n_templates < 3 || p->n_templates > LARND_MAX_TEMPLATES) { larnd_set_error("n_templates=%d unsupported", p->n_templates); return LARND_E_ARG; }
  if (p->nb_sampling_bins_per_pixel < 2 || p->nb_sampling_bins_per_pixel / 2 > 5 || p->nb_sampling_bins_per_pixel > 255) {
    larnd_set_error("nb_sampling_bins_per_pixel=%d unsupported", p->nb_sampling_bins_per_pixel);
    return LARND_E_ARG;
  }
  if (p->n_pixels_x > 32767 || p->n_pixels_y > 32767) { larnd_set_error("pixel plane too large"); return LARND_E_ARG; }
  if (need_lut) {
    if (!lut) { larnd_set_error("lut is null"); return LARND_E_ARG; }
    if (lut->L != p->signal_length) { larnd_set_error("LUT was built for signal_length %d, params say %d", lut->L, p->signal_length); return LARND_E_ARG; }
    // fewer bank rows than long_diff_template entries = a truncated bank: allowed, overflow is flagged on the device
    if (lut->ntpl > p->n_templates) { larnd_set_error("LUT has %d templates, long_diff_template only %d", lut->ntpl, p->n_templates); return LARND_E_ARG; }
    int need = p->nb_sampling_bins_per_pixel * p->number_pix_neighbors + p->nb_sampling_bins_per_pixel / 2;
    if (need > lut->nx || need > lut->ny) {
      // the reference would index past the LUT (take(mode='fill') -> NaN), sim_jax.py:221-222,443
      larnd_set_error("number_pix_neighbors=%d needs %d response bins per axis, LUT has %dx%d", p->number_pix_neighbors, need, lut->nx, lut->ny);
      return LARND_E_ARG;
    }
    if (p->number_pix_neighbors < 0 || p->number_pix_neighbors > 7) { larnd_set_error("number_pix_neighbors unsupported"); return LARND_E_ARG; }
  }
  return LARND_OK;
}

extern "C" int larnd_lut_prepare(const float* tracks_d, int64_t n, const larnd_columns_t* cols, const larnd_params_t* p,
                                 const larnd_lut_t* lut, int32_t n_events, void* workspace_d, size_t workspace_bytes,
                                 int32_t* counts_d, void* stream) {
  int rc = check_common(p, lut, true);
  if (rc) return rc;
  if ((!tracks_d && n > 0) || !cols || !counts_d || n < 0) { larnd_set_error("larnd_lut_prepare: bad argument"); return LARND_E_ARG; }
  Workspace ws;
  if (!larnd_carve_workspace(workspace_d, workspace_bytes, n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y, &ws)) {
    larnd_set_error("workspace too small: need %zu bytes", larnd_workspace_bytes(n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y));
    return LARND_E_CAPACITY;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = larnd_launch_prepare(tracks_d, n, *cols, *p, lut, ws, counts_d, st))) return rc;
  return larnd_launch_scan(ws, *p, counts_d, st);
}

extern "C" int larnd_lut_prepare_raw(const float* raw_d, int64_t m, const larnd_chop_columns_t* chop_cols, const larnd_columns_t* cols,
                                     double precision, const int64_t* offsets_d, int64_t n, const larnd_params_t* p,
                                     const larnd_lut_t* lut, int32_t n_events, void* workspace_d, size_t workspace_bytes,
                                     int32_t* counts_d, void* stream) {
  int rc = check_common(p, lut, true);
  if (rc) return rc;
  if ((!raw_d && m > 0) || !chop_cols || !cols || !offsets_d || !counts_d || n < 0 || m < 0 || !(precision > 0)) {
    larnd_set_error("larnd_lut_prepare_raw: bad argument");
    return LARND_E_ARG;
  }
  // both column maps describe the same raw row: the chopped columns the simulation reads must be the ones the chop produces
  if (chop_cols->ncols != cols->ncols || chop_cols->x != cols->x || chop_cols->y != cols->y || chop_cols->z != cols->z ||
      chop_cols->z_start != cols->z_start || chop_cols->z_end != cols->z_end || chop_cols->dx != cols->dx || chop_cols->dE != cols->dE) {
    larnd_set_error("larnd_lut_prepare_raw: chop_cols and cols disagree about the row layout");
    return LARND_E_ARG;
  }
  Workspace ws;
  if (!larnd_carve_workspace(workspace_d, workspace_bytes, n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y, &ws)) {
    larnd_set_error("workspace too small: need %zu bytes", larnd_workspace_bytes(n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y));
    return LARND_E_CAPACITY;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = larnd_launch_prepare_raw(raw_d, m, *chop_cols, *cols, precision, offsets_d, n, *p, lut, ws, counts_d, st))) return rc;
  return larnd_launch_scan(ws, *p, counts_d, st);
}

static int lut_accumulate_impl(int64_t n, const larnd_params_t* p, const larnd_lut_t* lut, int32_t n_events,
                               int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                               int32_t* unique_pixels_d, float* wfs_d, int64_t wfs_row_stride, int32_t* counts_d, void* stream,
                               unsigned long long* det_acc) {
  int rc = check_common(p, lut, true);
  if (rc) return rc;
  if (!unique_pixels_d || !wfs_d || !counts_d || npix_capacity < 1 || wfs_row_stride < p->n_ticks) {
    larnd_set_error("larnd_lut_accumulate: bad argument (null pointer, capacity < 1 or row stride < n_ticks)");
    return LARND_E_ARG;
  }
  Workspace ws;
  if (!larnd_carve_workspace(workspace_d, workspace_bytes, n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y, &ws)) {
    larnd_set_error("workspace too small");
    return LARND_E_CAPACITY;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & LARND_FLAG_WFS_ZERO)) LARND_CUDA(cudaMemsetAsync(wfs_d, 0, (size_t)npix_capacity * wfs_row_stride * sizeof(float), st));
  if ((rc = larnd_launch_unique(ws, *p, npix_capacity, /*extra=*/1, unique_pixels_d, counts_d, st))) return rc;
  return larnd_launch_accumulate(n, *p, lut, ws, npix_capacity, flags & LARND_FLAG_PUBLIC_MASK, wfs_d, wfs_row_stride, counts_d, st, det_acc);
}

extern "C" int larnd_lut_accumulate(int64_t n, const larnd_params_t* p, const larnd_lut_t* lut, int32_t n_events,
                                    int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                                    int32_t* unique_pixels_d, float* wfs_d, int64_t wfs_row_stride, int32_t* counts_d, void* stream) {
  return lut_accumulate_impl(n, p, lut, n_events, npix_capacity, flags, workspace_d, workspace_bytes, unique_pixels_d, wfs_d, wfs_row_stride,
                             counts_d, stream, nullptr);
}

extern "C" size_t larnd_deterministic_scratch_bytes(int32_t npix_capacity, int32_t n_ticks) {
  return (size_t)(npix_capacity > 0 ? npix_capacity : 0) * (size_t)(n_ticks > 0 ? n_ticks : 0) * sizeof(unsigned long long);
}

extern "C" int larnd_lut_accumulate_deterministic(int64_t n, const larnd_params_t* p, const larnd_lut_t* lut, int32_t n_events,
                                                  int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                                                  int32_t* unique_pixels_d, float* wfs_d, int64_t wfs_row_stride, int32_t* counts_d,
                                                  void* det_scratch_d, size_t det_scratch_bytes, void* stream) {
  if (!p || !det_scratch_d || det_scratch_bytes < larnd_deterministic_scratch_bytes(npix_capacity, p->n_ticks)) {
    larnd_set_error("larnd_lut_accumulate_deterministic: scratch missing or smaller than larnd_deterministic_scratch_bytes()");
    return LARND_E_ARG;
  }
  return lut_accumulate_impl(n, p, lut, n_events, npix_capacity, flags, workspace_d, workspace_bytes, unique_pixels_d, wfs_d, wfs_row_stride,
                             counts_d, stream, reinterpret_cast<unsigned long long*>(det_scratch_d));
}

extern "C" int larnd_lut_forward_deterministic(const float* tracks_d, int64_t n, const larnd_columns_t* cols, const larnd_params_t* p,
                                               const larnd_lut_t* lut, int32_t n_events, int32_t npix_capacity, int32_t flags,
                                               void* workspace_d, size_t workspace_bytes, int32_t* unique_pixels_d, float* wfs_d,
                                               int64_t wfs_row_stride, int32_t* counts_d, void* det_scratch_d, size_t det_scratch_bytes,
                                               void* stream) {
  int rc = larnd_lut_prepare(tracks_d, n, cols, p, lut, n_events, workspace_d, workspace_bytes, counts_d, stream);
  if (rc) return rc;
  return larnd_lut_accumulate_deterministic(n, p, lut, n_events, npix_capacity, flags, workspace_d, workspace_bytes, unique_pixels_d, wfs_d,
                                            wfs_row_stride, counts_d, det_scratch_d, det_scratch_bytes, stream);
}

extern "C" int larnd_lut_forward(const float* tracks_d, int64_t n, const larnd_columns_t* cols, const larnd_params_t* p,
                                 const larnd_lut_t* lut, int32_t n_events, int32_t npix_capacity, int32_t flags,
                                 void* workspace_d, size_t workspace_bytes, int32_t* unique_pixels_d, float* wfs_d,
                                 int64_t wfs_row_stride, int32_t* counts_d, void* stream) {
  int rc = larnd_lut_prepare(tracks_d, n, cols, p, lut, n_events, workspace_d, workspace_bytes, counts_d, stream);
  if (rc) return rc;
  return larnd_lut_accumulate(n, p, lut, n_events, npix_capacity, flags, workspace_d, workspace_bytes, unique_pixels_d,
                              wfs_d, wfs_row_stride, counts_d, stream);
}

extern "C" int larnd_lut_backward(int64_t n, const larnd_params_t* p, const larnd_lut_t* lut, int32_t n_events,
                                  int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                                  const int32_t* counts_d, const float* g_wfs_d, int64_t g_row_stride,
                                  float* grad_params_d, void* stream) {
  int rc = check_common(p, lut, true);
  if (rc) return rc;
  if (!g_wfs_d || !grad_params_d || !counts_d) { larnd_set_error("larnd_lut_backward: null argument"); return LARND_E_ARG; }
  Workspace ws;
  if (!larnd_carve_workspace(workspace_d, workspace_bytes, n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y, &ws)) {
    larnd_set_error("workspace too small");
    return LARND_E_CAPACITY;
  }
  return larnd_launch_accumulate_bwd(n, *p, lut, ws, npix_capacity, flags & LARND_FLAG_PUBLIC_MASK, g_wfs_d, g_row_stride, grad_params_d,
                                     counts_d, (cudaStream_t)stream);
}

extern "C" int larnd_lut_backward_steps(int64_t n, const larnd_params_t* p, const larnd_lut_t* lut, int32_t n_events,
                                        int32_t npix_capacity, int32_t flags, void* workspace_d, size_t workspace_bytes,
                                        const int32_t* counts_d, const void* steps_d, size_t steps_bytes, float* grad_params_d,
                                        void* stream) {
  int rc = check_common(p, lut, true);
  if (rc) return rc;
  if (!steps_d || !grad_params_d || !counts_d || steps_bytes < larnd_steps_layout(npix_capacity, nullptr, nullptr)) {
    larnd_set_error("larnd_lut_backward_steps: null argument or steps buffer smaller than larnd_fee_steps_bytes(npix_capacity)");
    return LARND_E_ARG;
  }
  Workspace ws;
  if (!larnd_carve_workspace(workspace_d, workspace_bytes, n, n_events, p->n_tpc, p->n_pixels_x, p->n_pixels_y, &ws)) {
    larnd_set_error("workspace too small");
    return LARND_E_CAPACITY;
  }
  StepsView v;
  larnd_steps_layout(npix_capacity, const_cast<void*>(steps_d), &v);
  return larnd_launch_accumulate_bwd(n, *p, lut, ws, npix_capacity, (flags & LARND_FLAG_PUBLIC_MASK) | 1, nullptr, 0, grad_params_d, counts_d,
                                     (cudaStream_t)stream, &v);
}
