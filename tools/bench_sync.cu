// micro-benchmark of inter-CTA synchronisation flavours (cooperative launch):
//   0 counter: atomicAdd + one poller per CTA (+ fences), then read R partials from L2
//   1 flags:   16-byte {value, epoch} descriptors, warp 0 polls all ranks (as pf_device.cuh)
//   2 flags, fence only on lane 0 after __syncwarp
//   3 flags, no fences at all (lower bound; NOT safe for bulk data exchange)
//   4 flags, fence.release.gpu before publishing only (round 2 filter kernel)
//   5 flags as 4, two-level: groups of 16 CTAs, leaders republish group totals
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_sync tools/bench_sync.cu
#include <cstdio>
#include <cuda_runtime.h>
struct Desc { unsigned long long val, tag; };
__device__ __forceinline__ void dst(Desc* p, unsigned long long v, unsigned long long t) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v), "l"(t) : "memory");
}
__device__ __forceinline__ void dld(const Desc* p, unsigned long long& v, unsigned long long& t) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(p) : "memory");
}
__global__ void k(int flavour, int iters, unsigned* counter, double* part, Desc* slots, double* out) {
  const int R = gridDim.x, rank = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ double res;
  double acc = 0;
  unsigned target = 0;
  unsigned long long epoch = 0;
  for (int it = 0; it < iters; ++it) {
    double mine = rank + it;
    if (flavour == 0) {
      if (threadIdx.x == 0) part[rank] = mine;
      __syncthreads();
      if (threadIdx.x == 0) {
        target += R;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
        __threadfence();
      }
      __syncthreads();
      if (warp == 0) {
        double x = 0;
        for (int r = lane; r < R; r += 32) x += __ldcg(part + r);
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) res = x;
      }
      __syncthreads();
      acc += res;
      __syncthreads();
    } else {
      __syncthreads();
      if (warp == 0) {
        Desc* base = slots + (epoch & 1) * R;
        if (lane == 0) {
          if (flavour == 4) asm volatile("fence.release.gpu;" ::: "memory");
          else if (flavour != 3) __threadfence();
          dst(base + rank, (unsigned long long)__double_as_longlong(mine), epoch + 1);
        }
        double x = 0;
        for (int r = lane; r < R; r += 32) {
          unsigned long long v, t;
          do { dld(base + r, v, t); } while ((long long)(t - (epoch + 1)) < 0);
          x += __longlong_as_double((long long)v);
        }
        if (flavour == 1) __threadfence();
        if (flavour == 2) { __syncwarp(); if (lane == 0) __threadfence(); }
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) res = x;
      }
      epoch++;
      __syncthreads();
      acc += res;
    }
  }
  if (threadIdx.x == 0) out[rank] = acc;
}
int main() {
  unsigned* counter; double *part, *out; Desc* slots;
  cudaMalloc(&counter, 4); cudaMalloc(&part, 8 * 1024); cudaMalloc(&out, 8 * 1024); cudaMalloc(&slots, 16 * 2048);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  for (int R : {8, 16, 32, 64, 128, 148}) for (int threads : {256, 1024}) for (int fl = 0; fl < 5; ++fl) {
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemset(counter, 0, 4); cudaMemset(slots, 0, 16 * 2048);
      int it = iters;
      void* args[] = {&fl, &it, &counter, &part, &slots, &out};
      cudaEventRecord(e0);
      cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(R), dim3(threads), args, 0, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      if (e != cudaSuccess) { printf("launch failed %s\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("R=%3d threads=%4d flavour=%d : %.2f us per collective\n", R, threads, fl, best * 1e3 / iters);
  }
  return 0;
}
