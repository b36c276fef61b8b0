// Micro-benchmark: per-SM global store rate of the conv epilogue's access patterns (one CTA per SM, 8 warps).
//   A: lane = pixel, 32 B per lane (st.global.v8), pixel stride `cs` floats -> 32 different 128-byte lines per instruction
//   B: 8-lane groups write 128 contiguous bytes (st.global.v4), 4 lines per instruction
//   C: lanes write 16 B each, 512 contiguous bytes per instruction
// and the matching load patterns (scalar per-lane-line loads vs v8 vs group-coalesced v4).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__global__ void k_store(float* y, int cs, int tiles, int mode, long long* clk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ float dyn[];
  if (tiles < 0) dyn[threadIdx.x] = 1.f;
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    const int64_t tile = (int64_t)blockIdx.x * tiles + t;
    // tile = 128 pixels x 128 channels (fp32); warp w -> pixels (w&3)*32.., channel half (w>>2)*... two 32-ch chunks
    const int q = warp & 3, half = warp >> 2;
    float* base = y + (tile * 128 + q * 32) * (int64_t)cs;
    for (int k = 0; k < 2; ++k) {
      const int c0 = (half + 2 * k) * 32;
      if (mode == 0) {
        float* p = base + (int64_t)lane * cs + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + j), "r"(t) : "memory");
      } else if (mode == 1) {
        const int g = lane >> 3, i = lane & 7;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          float* p = base + (int64_t)(g * 8 + kk) * cs + c0 + 4 * i;
          asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(t) : "memory");
        }
      } else if (mode == 2) {   // lane = pixel, v4 stores (16 B per lane, 32 lines per instr)
        float* p = base + (int64_t)lane * cs + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p + j), "r"(t) : "memory");
      }
    }
  }
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

__global__ void k_load(const float* y, int cs, int tiles, int mode, long long* clk, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    const int64_t tile = (int64_t)blockIdx.x * tiles + t;
    const int q = warp & 3, half = warp >> 2;
    const float* base = y + (tile * 128 + q * 32) * (int64_t)cs;
    for (int k = 0; k < 2; ++k) {
      const int c0 = (half + 2 * k) * 32;
      if (mode == 0) {
        const float* p = base + (int64_t)lane * cs + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += __ldg(p + j);
      } else if (mode == 1) {
        const float* p = base + (int64_t)lane * cs + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t v[8];
          asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]),
                       "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p + j));
#pragma unroll
          for (int e = 0; e < 8; ++e) acc += __uint_as_float(v[e]);
        }
      } else {
        const int g = lane >> 3, i = lane & 7;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const float4 v = *reinterpret_cast<const float4*>(base + (int64_t)(g * 8 + kk) * cs + c0 + 4 * i);
          acc += v.x + v.y + v.z + v.w;
        }
      }
    }
  }
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles = 28, cs = 256;
  const size_t n = (size_t)sms * tiles * 128 * cs;
  float* y; long long* clk; float* sink;
  cudaMalloc(&y, n * 4); cudaMalloc(&clk, sms * 8); cudaMalloc(&sink, 4);
  cudaMemset(y, 0, n * 4);
  long long h[256];
  cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int smem : {0, 200 * 1024})
  for (int grid : {sms}) {
    printf("-- dynamic smem %d KB\n", smem / 1024);
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) k_store<<<grid, 256, smem>>>(y, cs, tiles, mode, clk);
      cudaDeviceSynchronize();
      cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < grid; ++i) m += (double)h[i] / grid;
      printf("store grid %3d mode %d: %.0f clk per 64 KB tile (%.1f B/clk/SM)\n", grid, mode, m / tiles, 65536.0 * tiles / m);
    }
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) k_load<<<grid, 256>>>(y, cs, tiles, mode, clk, sink);
      cudaDeviceSynchronize();
      cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < grid; ++i) m += (double)h[i] / grid;
      printf("load  grid %3d mode %d: %.0f clk per 64 KB tile (%.1f B/clk/SM)\n", grid, mode, m / tiles, 65536.0 * tiles / m);
    }
  }
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
