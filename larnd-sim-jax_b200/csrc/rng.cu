// jax.random-compatible normals on the device: Threefry-2x32-20 driven the way jax/_src/prng.py drives
// it, so that noise-on FEE runs (fee_jax.py:186,237-255,271) and mc_diff smearing (detsim_jax.py:393) use the same
// random BITS as the reference for a given jax.random.key(seed).  normal = sqrt(2) * erfinv(uniform(-1,1)) with XLA's
// float32 ErfInv polynomial (Giles); transcendental rounding may differ from XLA's by an ulp.
#include "larnd_common.cuh"

namespace {

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int d) { return (x << d) | (x >> (32 - d)); }

__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl32(x1, rot[r & 1][i]) ^ x0;
    }
    x0 += ks[(r + 1) % 3];
    x1 += ks[(r + 2) % 3] + (uint32_t)(r + 1);
  }
}

__device__ __forceinline__ float erfinv_xla(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return fabsf(x) == 1.0f ? copysignf(INFINITY, x) : p * x;
}

__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float fl = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(fl, 1.0f - lo), lo));
  return 1.41421354f * erfinv_xla(u);
}

// partitionable layout: element i uses the 64-bit counter i (hi, lo) and the XOR of the two output words
__global__ void k_normal_partitionable(uint32_t k0, uint32_t k1, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t x0 = (uint32_t)((uint64_t)i >> 32), x1 = (uint32_t)i;
    threefry2x32(k0, k1, x0, x1);
    out[i] = bits_to_normal(x0 ^ x1);
  }
}

// original layout: counts = iota(n) (zero padded to even length) cut in two halves; block b = (counts[b], counts[half+b])
// yields elements b and half + b
__global__ void k_normal_original(uint32_t k0, uint32_t k1, int64_t n, float* __restrict__ out) {
  const int64_t half = (n + 1) / 2;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < half; b += (int64_t)gridDim.x * blockDim.x) {
    uint32_t x0 = (uint32_t)b, x1 = (half + b < n) ? (uint32_t)(half + b) : 0u;
    threefry2x32(k0, k1, x0, x1);
    out[b] = bits_to_normal(x0);
    if (half + b < n) out[half + b] = bits_to_normal(x1);
  }
}

void split_one(uint32_t& k0, uint32_t& k1, int partitionable) {
  // random.split(key, 1)[0]: partitionable -> block of counter (0, 0); original -> counts iota(2) = block (0, 1)
  uint32_t x0 = 0, x1 = partitionable ? 0u : 1u;
  threefry2x32(k0, k1, x0, x1);
  k0 = x0;
  k1 = x1;
}

int launch_normal(uint32_t k0, uint32_t k1, int64_t n, int partitionable, float* out, cudaStream_t st) {
  if (n <= 0) return LARND_OK;
  if (!partitionable && n >= ((int64_t)1 << 32)) { larnd_set_error("original threefry layout supports < 2^32 elements"); return LARND_E_ARG; }
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (partitionable) k_normal_partitionable<<<(unsigned)blocks, 256, 0, st>>>(k0, k1, n, out);
  else k_normal_original<<<(unsigned)blocks, 256, 0, st>>>(k0, k1, n, out);
  LARND_LAUNCH_CHECK("k_normal");
  return LARND_OK;
}

}  // namespace

extern "C" int larnd_rng_split(const uint32_t key[2], int num, int partitionable, uint32_t* keys_out /* [num][2], host */) {
  if (!key || !keys_out || num < 1) { larnd_set_error("larnd_rng_split: bad argument"); return LARND_E_ARG; }
  if (partitionable) {
    for (int i = 0; i < num; ++i) {
      uint32_t x0 = 0, x1 = (uint32_t)i;
      threefry2x32(key[0], key[1], x0, x1);
      keys_out[2 * i] = x0;
      keys_out[2 * i + 1] = x1;
    }
  } else {  // flat = threefry_2x32(key, iota(2 num)) with the two-halves pairing; key i = (flat[2i], flat[2i+1])
    for (int b = 0; b < num; ++b) {
      uint32_t x0 = (uint32_t)b, x1 = (uint32_t)(num + b);
      threefry2x32(key[0], key[1], x0, x1);
      // flat[b] = x0, flat[num + b] = x1
      keys_out[(b / 2) * 2 + (b % 2)] = x0;
      const int f = num + b;
      keys_out[(f / 2) * 2 + (f % 2)] = x1;
    }
  }
  return LARND_OK;
}

extern "C" int larnd_rng_normal(const uint32_t key[2], int64_t n, int partitionable, float* out_d, void* stream) {
  if (!key || (!out_d && n > 0) || n < 0) { larnd_set_error("larnd_rng_normal: bad argument"); return LARND_E_ARG; }
  return launch_normal(key[0], key[1], n, partitionable, out_d, (cudaStream_t)stream);
}

extern "C" int larnd_rng_fee_noise(const uint32_t key[2], int32_t npix, int32_t n_adc, int partitionable, float* noise_d, void* stream) {
  if (!key || !noise_d || npix < 0 || n_adc < 1) { larnd_set_error("larnd_rng_fee_noise: bad argument"); return LARND_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_normal(key[0], key[1], npix, partitionable, noise_d, st);  // q_sum_base, fee_jax.py:186
  if (rc) return rc;
  uint32_t k0 = key[0], k1 = key[1];
  split_one(k0, k1, partitionable);  // init_loop key, fee_jax.py:271
  for (int i = 0; i < n_adc; ++i) {
    split_one(k0, k1, partitionable);  // extra_noise, :237-238
    if ((rc = launch_normal(k0, k1, npix, partitionable, noise_d + (size_t)(1 + i) * npix, st))) return rc;
    split_one(k0, k1, partitionable);  // q_adc_pass, :252-253
    if ((rc = launch_normal(k0, k1, npix, partitionable, noise_d + (size_t)(1 + n_adc + i) * npix, st))) return rc;
    split_one(k0, k1, partitionable);  // q_adc_fail, :254-255
    if ((rc = launch_normal(k0, k1, npix, partitionable, noise_d + (size_t)(1 + 2 * n_adc + i) * npix, st))) return rc;
  }
  return LARND_OK;
}
