// Micro-benchmark: throughput of red.global.add.f32 flush patterns into a large (Npix, 2048) float32 waveform
// buffer, to decide the output layout of the accumulate kernels.  Each warp flushes 32-run x 128-tick "unit tiles"
// (16 KB) to pseudo-random rows, with no other work:
//   A  scalar RED, lane <-> tick          (4 instr per run window, 128 B contiguous per instr)   [k_acc_tiles today]
//   B  RED.v2 in the mma.sync C layout    (8 rows x 32 B per instr)
//   C  RED.v4, thread <-> row             (32 rows x 16 B per instr)                             [TMEM 32x32b layout]
//   D  RED.v4, lane <-> 4 consecutive ticks (1 instr per run window, 512 B contiguous)
//   E  scalar RED in the mma.sync C layout (8 rows x 4 lanes x 4 B, stride 2)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/ubench_red.cu -o gpurun_out/ubench_red
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) k_red(float* wfs, int npix, int stride, int iters) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int it = 0; it < iters; ++it) {
    // 32 runs of a tile: row and first tick of every run (lane <-> run)
    const uint32_t h = hash32((uint32_t)(gw * 7919 + it) * 32u + lane);
    const int myrow = 1 + (int)(h % (uint32_t)(npix - 1));
    const int mytick = 4 * (int)((h >> 20) % 450u);  // 16-byte aligned window start
    const float v = 1.0f + (float)lane;
    if (MODE == 0) {
      for (int p = 0; p < 32; ++p) {
        const int row = __shfl_sync(0xffffffffu, myrow, p), tk = __shfl_sync(0xffffffffu, mytick, p);
        float* dst = wfs + (int64_t)row * stride + tk + lane;
#pragma unroll
        for (int s = 0; s < 4; ++s) atomicAdd(dst + 32 * s, v);
      }
    } else if (MODE == 1 || MODE == 4) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int row = __shfl_sync(0xffffffffu, myrow, 8 * mt + g), tk = __shfl_sync(0xffffffffu, mytick, 8 * mt + g);
        float* dst = wfs + (int64_t)row * stride + tk + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
          if (MODE == 1) atomicAdd(reinterpret_cast<float2*>(dst + 8 * nt), make_float2(v, v));
          else { atomicAdd(dst + 8 * nt, v); atomicAdd(dst + 8 * nt + 1, v); }
        }
      }
    } else if (MODE == 2) {
      float* dst = wfs + (int64_t)myrow * stride + mytick;
#pragma unroll 8
      for (int c = 0; c < 32; ++c) atomicAdd(reinterpret_cast<float4*>(dst + 4 * c), make_float4(v, v, v, v));
    } else if (MODE == 3) {
      for (int p = 0; p < 32; ++p) {
        const int row = __shfl_sync(0xffffffffu, myrow, p), tk = __shfl_sync(0xffffffffu, mytick, p);
        float* dst = wfs + (int64_t)row * stride + tk + 4 * lane;
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v, v, v, v));
      }
    }
  }
}

template <int MODE>
static void run(const char* name, float* wfs, int npix, int stride, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 2;
  k_red<MODE><<<grid, 256>>>(wfs, npix, stride, 4);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_red<MODE><<<grid, 256>>>(wfs, npix, stride, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)grid * 8 * iters * 32.0 * 128 * 4;
  const double windows = (double)grid * 8 * iters * 32.0;
  printf("%-46s %8.3f ms  %8.1f GB/s  %8.2f G run-windows/s  err=%s\n", name, ms, bytes / ms * 1e-6, windows / ms * 1e-6,
         cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const int npix = argc > 1 ? atoi(argv[1]) : 262144, stride = 2048;
  const int iters = argc > 2 ? atoi(argv[2]) : 400;
  float* wfs = nullptr;
  if (cudaMalloc(&wfs, (size_t)npix * stride * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(wfs, 0, (size_t)npix * stride * 4);
  printf("npix %d (%.2f GB), %d tiles per warp\n", npix, npix * (double)stride * 4e-9, iters);
  for (int rep = 0; rep < 2; ++rep) {
    run<0>("A scalar, lane<->tick (128 B/instr)", wfs, npix, stride, iters);
    run<1>("B v2, mma C layout (8 rows x 32 B)", wfs, npix, stride, iters);
    run<4>("E scalar, mma C layout (8 rows x 4 x 4 B)", wfs, npix, stride, iters);
    run<2>("C v4, thread<->row (32 rows x 16 B)", wfs, npix, stride, iters);
    run<3>("D v4, lane<->4 ticks (512 B/instr)", wfs, npix, stride, iters);
  }
  cudaFree(wfs);
  return 0;
}
