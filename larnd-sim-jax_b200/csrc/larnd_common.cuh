// Internal declarations shared by the larnd_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "larnd_b200.h"

#define LARND_WARP 32
#define LARND_OFF_PAD 4  // response rows are stored as [0,0,R[0..L-1],0,0]: Lp = L + 4, sample k at index k+2

// Device tables built by larnd_lut_create (see lut_tables.cu).
struct larnd_lut {
  int ntpl, nx, ny, nt, L, Lp;
  float* r0;  // [nx*ny][Lp]          template 0, last L samples of every (ci,cj) bin
  float* rm;  // [ntpl][25][Lp]       all templates, collecting bins ci,cj < 5
  float* c0;  // [nx*ny][nt]          running sum of template 0
  float* cm;  // [ntpl][25][nt]       running sum of all templates, collecting bins
  // Sum of the neighbour rows over the whole (2n+1)^2 neighbourhood for every in-pixel bin (bxm, bym):
  //   sr[bxm*nb+bym][Lp] = sum_{dx,dy} r0[ci(bxm,dx)][cj(bym,dy)],   sc likewise for c0.
  // Everything a segment deposits on neighbour pixels that are NOT main pixels lands in waveform row 0
  // (sim_jax.py:724-725); by linearity that is (sum over all neighbours) - (the few that are main pixels).
  int sum_nb, sum_n;
  float* sr;  // [nb*nb][Lp]
  float* sc;  // [nb*nb][nt]
  // Shifted copies for the lane <-> 4-ticks tile kernels (lut_tables.cu::k_shift_rows): T[s][row][i] = R_row[i - s - 8],
  // s = 0..3, rows of lps = 128 * nsv + 8 floats (nsv = 128-tick slots of a run frame; 0 = L too long for those kernels).
  int nsv, lps;
  float* t0s;  // [4][nx*ny + nb*nb][lps]: template 0, then the neighbourhood-sum rows (built by larnd_lut_prepare_neighbours)
  float* tms;  // [4][ntpl*25][lps]
};
int larnd_lut_check_neighbours(const larnd_lut* lut, int nb, int n);

// Views into the caller-provided workspace.
struct Workspace {
  float* rec;         // LARND_NFIELDS x N (SoA)
  uint32_t* bitmap;   // n_words
  uint32_t* wprefix;  // n_words (exclusive popcount prefix)
  uint32_t* bsums;    // block sums for the scan
  float* partials;    // per-chunk gradient partials (n_chunks_max x 16)
  int64_t n;
  int64_t n_words;
  int64_t n_scan_blocks;
  int32_t pid_offset;  // nx*ny*ntpc: bit index of pixel id p is p + pid_offset (event -1 ids are negative)
  int64_t n_chunks_max;
  // class-sorted accumulate (accumulate_sorted.cu): run records before / after the counting sort, class histogram,
  // offsets and scatter cursors, tile table, counters {n_runs, n_tiles, tile counter}
  char* runs_tmp;
  char* runs;
  int* class_count;
  int* class_start;
  int* cursor;
  char* tile_info;
  int* gcnt;
  float* row0;
};

#define LARND_NCLS_MAX (LARND_MAX_TEMPLATES * 11 * 11 * 5)  // response classes: template index x in-pixel bin x tick span (x 4 alignment-shift sub-buckets in the sort)
#define LARND_ROW0_COPIES 512                            // private garbage-row copies (>= CTAs of k_acc_tiles)
#define LARND_ROW0_TICKS_MAX 8192
#define LARND_BWD_SORTED_SLOTS 1280                      // per-warp gradient partials of the class-sorted backward kernel
#define LARND_ACC_SLOW_ONLY 0x100                        // internal flag of larnd_launch_accumulate
#define LARND_SORTED_MIN_SEGMENTS 200000                 // below this the chunk kernel wins (few runs per class)
size_t larnd_sorted_workspace_bytes(int64_t n);
void larnd_carve_sorted(char* p, int64_t n, Workspace* ws);
int larnd_sorted_supported(const larnd_params_t& p, const larnd_lut* lut);

#define LARND_SCAN_WORDS_PER_BLOCK 2048
#define LARND_CHUNK 128  // segments per accumulate CTA (capacity of the chunk kernels' staging arrays)
#define LARND_SMALL_CHUNK_SLOTS 2560  // extra per-chunk partial slots for batches cut into chunks smaller than LARND_CHUNK
// Segments per CTA of the chunk kernels: a fit-sized batch (~20 k segments) cut into 128-segment chunks is 155 CTAs — one per
// SM, each walking ~30 runs serially (0.40 / 0.73 ms forward / backward); smaller chunks trade a few more window flushes for
// four times the CTAs.  A function of n only, so the deterministic mode stays reproducible.
#ifndef LARND_CHUNK_MIN
#define LARND_CHUNK_MIN 32
#endif
#ifndef LARND_CHUNK_CTAS
#define LARND_CHUNK_CTAS (4 * 148)
#endif
static inline int larnd_chunk_size(int64_t n) {
  int c = LARND_CHUNK;
  while (c > LARND_CHUNK_MIN && (n + c - 1) / c < LARND_CHUNK_CTAS) c >>= 1;
  return c;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Compact upstream gradient of the waveforms when it comes from the front end (larnd_fee_backward_steps): for every waveform
// row the VJP of get_adc_values is a step function of the tick, dL/dwfs_full[row, col] = sum over the row's events e with
// pos[e] >= col of val[e] (col >= 1; zero for col 0), with at most 2 * MAX_ADC_VALUES events (fee.cu).  168 bytes per row
// instead of n_ticks floats.
constexpr int LARND_STEPS_MAX = 2 * LARND_MAX_ADC;
constexpr int LARND_STEPS_WORDS = 48;   // one 192-byte record per row: pos[0..19] | val[20..39] | n[40] | padding
struct StepsView {
  const int32_t* rec;   // [npix][LARND_STEPS_WORDS]
};
inline size_t larnd_steps_layout(int32_t npix, void* base, StepsView* v) {
  if (v) v->rec = reinterpret_cast<const int32_t*>(base);
  return align_up((size_t)npix * LARND_STEPS_WORDS * sizeof(int32_t), 256);
}
#ifdef __CUDACC__
// The events of one row spread over the lanes of a warp (lane e <-> event e); value(col) is a warp-collective call.
struct RowSteps {
  int n, pos;
  float val;
  // three independent loads (one record, two cache lines): issue them for several rows before using any (fix())
  __device__ __forceinline__ void load(const StepsView& s, int row, int lane) {
    const int32_t* r = s.rec + (int64_t)row * LARND_STEPS_WORDS;
    const int l = min(lane, LARND_STEPS_MAX - 1);
    n = __ldg(r + 2 * LARND_STEPS_MAX);
    pos = __ldg(r + l);
    val = __int_as_float(__ldg(r + LARND_STEPS_MAX + l));
  }
  __device__ __forceinline__ void fix(int lane) {
    if (lane >= n) { pos = -1; val = 0.0f; }
  }
  // dL/dwfs_full[row, col] for this lane's col (every lane of the warp must call)
  __device__ __forceinline__ float value(int col) const {
    float g = 0.0f;
    for (int e = 0; e < n; ++e) {
      const int pe = __shfl_sync(0xffffffffu, pos, e);
      const float ve = __shfl_sync(0xffffffffu, val, e);
      g += (pe >= col && col >= 1) ? ve : 0.0f;
    }
    return g;
  }
};
#endif


bool larnd_carve_workspace(void* base, size_t bytes, int64_t n, int32_t n_events, int32_t ntpc, int32_t nx,
                           int32_t ny, Workspace* ws);
void larnd_set_error(const char* fmt, ...);
int larnd_check_cuda(cudaError_t e, const char* what);
#define LARND_CUDA(call)                                   \
  do {                                                     \
    int _rc = larnd_check_cuda((call), #call);             \
    if (_rc != 0) return _rc;                              \
  } while (0)
// every kernel launch of the library goes through this macro: it also feeds larnd_launch_count()
extern unsigned long long g_launch_count;
#define LARND_LAUNCH_CHECK(name)                                      \
  do {                                                               \
    __atomic_add_fetch(&g_launch_count, 1ull, __ATOMIC_RELAXED);     \
    LARND_CUDA(cudaGetLastError());                                  \
  } while (0)

// optional event timing around the dominant kernels (larnd_profile_enable / larnd_profile_read)
struct ProfSlot { cudaEvent_t start, stop; bool used; };
extern bool g_prof_on;
extern ProfSlot g_prof[LARND_PROF_SLOTS];
inline void prof_begin(int slot, cudaStream_t st) { if (g_prof_on) cudaEventRecord(g_prof[slot].start, st); }
inline void prof_end(int slot, cudaStream_t st) { if (g_prof_on) { cudaEventRecord(g_prof[slot].stop, st); g_prof[slot].used = true; } }

// ---- device helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ int floordiv_i(int a, int b) {  // python-style floor division, b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

// Segments whose response window ends inside the readout are handled by the class-sorted kernel (windows sticking out
// at the LOW end are handled there too: garbage column 0); the rest goes through accumulate.cu's per-segment path.
__device__ __forceinline__ bool seg_is_fast(int T0, int L, int nticks) { return T0 + L <= nticks - 2; }

// The 5-wide transverse-diffusion stencil around in-pixel bin bq touches bins bq-2 .. bq+2; bins that fall on the same
// pixel (offset -1 / 0 / +1) with the same response row (|distance to the pixel centre|) are merged into *groups*.
// Fills, for one bq: number of groups, and per group pixel offset + 1, response index, member mask (bit i = stencil bin i).
__device__ __forceinline__ void build_bin_groups(int bq, int nb, int half2, unsigned char& g_n, unsigned char (&g_ox)[5],
                                                 unsigned char (&g_ci)[5], unsigned char (&g_mask)[5]) {
  int ng = 0;
  for (int i = 0; i < LARND_NB_TRAN_BINS; ++i) {
    int qb = bq + i - (LARND_NB_TRAN_BINS - 1) / 2, ox = 0;
    if (qb < 0) { qb += nb; ox = -1; } else if (qb >= nb) { qb -= nb; ox = 1; }
    const int ci = abs(2 * qb - half2) >> 1;
    int g = -1;
    for (int k = 0; k < ng; ++k)
      if (g_ox[k] == ox + 1 && g_ci[k] == ci) g = k;
    if (g < 0) { g = ng++; g_ox[g] = ox + 1; g_ci[g] = ci; g_mask[g] = 0; }
    g_mask[g] |= 1 << i;
  }
  g_n = ng;
}

// pixel2id with int32 wrap-around (detsim_jax.py:232-244; x64 is never enabled in the reference)
__device__ __forceinline__ int pixel2id_dev(int px, int py, int ep, int nx, int ny) {
  if (px >= nx || py >= ny || px < 0 || py < 0) return -1;
  unsigned u = (unsigned)ep * (unsigned)ny + (unsigned)py;
  u = u * (unsigned)nx + (unsigned)px;
  return (int)u;
}

struct RowLookup {
  const uint32_t* bitmap;
  const uint32_t* wprefix;
  int64_t n_words;
  int pid_offset;
  int n_unique;  // number of distinct main pixel ids
  int n_neg;     // ids < -1
  int npix;      // padded length of unique_pixels
};

// Row of pixel id `pid` in the reference's sorted, -1 padded unique_pixels (sim_jax.py:717-725,152-154).
// Returns -1 when pid is not in the list.  pid == -1 always matches (the appended -1 entry).
__device__ __forceinline__ int lookup_row(const RowLookup& lk, int pid) {
  if (pid == -1) return lk.n_neg;
  long long b = (long long)pid + lk.pid_offset;
  if (b < 0 || (b >> 5) >= lk.n_words) return -1;
  uint32_t w = __ldg(lk.bitmap + (b >> 5));
  uint32_t bit = 1u << (b & 31);
  if (!(w & bit)) return -1;
  int rank = (int)__ldg(lk.wprefix + (b >> 5)) + __popc(w & (bit - 1));
  return pid < 0 ? rank : lk.npix - lk.n_unique + rank;
}

// Host-side record of which workspaces hold valid sorted run / tile tables (sorted_runs.cuh), keyed on the record base
// pointer: set by a build, dropped by whatever rewrites the records (prepare kernels).  LARND_FLAG_REUSE_RUNS consults it.
void larnd_runs_cache_set(const void* rec, int64_t n, const void* lut, int n_ticks);
void larnd_runs_cache_drop(const void* rec);
bool larnd_runs_cache_valid(const void* rec, int64_t n, const void* lut, int n_ticks);

// kernels / launchers implemented in the .cu files
int larnd_launch_prepare(const float* tracks, int64_t n, const larnd_columns_t& cols, const larnd_params_t& p,
                         const larnd_lut* lut, const Workspace& ws, int32_t* counts, cudaStream_t st);
int larnd_launch_prepare_raw(const float* raw, int64_t m, const larnd_chop_columns_t& cc, const larnd_columns_t& cols, double precision,
                             const int64_t* offsets, int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                             int32_t* counts, cudaStream_t st);
int larnd_launch_unique(const Workspace& ws, const larnd_params_t& p, int32_t npix_capacity, int32_t extra, int32_t* unique_pixels,
                        int32_t* counts, cudaStream_t st);
int larnd_launch_scan(const Workspace& ws, const larnd_params_t& p, int32_t* counts, cudaStream_t st);
int larnd_launch_accumulate(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                            int32_t npix_capacity, int32_t flags, float* wfs, int64_t wfs_stride, const int32_t* counts, cudaStream_t st,
                            unsigned long long* det_acc = nullptr);
int larnd_launch_accumulate_sorted(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                   int32_t npix_capacity, int32_t flags, float* wfs, int64_t wfs_stride, const int32_t* counts,
                                   cudaStream_t st);
int larnd_launch_accumulate_bwd_sorted(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                       int32_t npix_capacity, int32_t flags, const float* g_wfs, int64_t g_stride,
                                       float* sorted_partials, int* n_slots_out, const int* gflag, const int32_t* counts,
                                       cudaStream_t st, const StepsView* steps = nullptr);
int larnd_launch_accumulate_bwd(int64_t n, const larnd_params_t& p, const larnd_lut* lut, const Workspace& ws,
                                int32_t npix_capacity, int32_t flags, const float* g_wfs, int64_t g_stride, float* grad_params,
                                const int32_t* counts, cudaStream_t st, const StepsView* steps = nullptr);
