// launch.cuh -- launch / occupancy of the persistent filter kernel variants.
//
// The filter kernel is instantiated per (model policy, particles per thread,
// launch bound); those instantiations dominate build time, so each model family
// lives in its own translation unit (k_*.cu) behind the plain functions of
// filters.h, and this header is the code they share.  Status convention of this
// layer: 0 = ok, > 0 = a cudaError_t, -1 = variant not instantiated.
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include "pf_device.cuh"
#include "pf_small.cuh"

#ifndef PF_SMALL_MAXT
#define PF_SMALL_MAXT 512
#define PF_SMALL_MINB 2
#endif

namespace pfl {

struct Geometry {
  int P, block;
  uint32_t ranks;
  int sync_mode;
};

#define PFL_CU(call)                       \
  do {                                     \
    cudaError_t e_ = (call);               \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

template <class Model, int P, int MAXT, int MINB>
int launch_V(const pf::FilterParams& fp, const typename Model::Params& mp, const Geometry& g,
             uint32_t n_filters, cudaStream_t stream) {
  auto kern = pf::filter_kernel<Model, P, MAXT, MINB>;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_filters * g.ranks, 1, 1);
  cfg.blockDim = dim3(g.block, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (g.sync_mode == pf::SYNC_CLUSTER) {
    if (g.ranks > 8) PFL_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g.ranks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
  } else if (g.sync_mode == pf::SYNC_GROUP) {
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.numAttrs = 1;
  }
  PFL_CU(cudaLaunchKernelEx(&cfg, kern, fp, mp));
  return 0;
}

// co-resident filters (groups / clusters) or CTAs (single-CTA filters) of this geometry
template <class Model, int P, int MAXT, int MINB>
int occupancy_V(const Geometry& g, int* out) {
  auto kern = pf::filter_kernel<Model, P, MAXT, MINB>;
  int dev = 0, sms = 0;
  PFL_CU(cudaGetDevice(&dev));
  PFL_CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (g.sync_mode == pf::SYNC_CLUSTER) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(g.ranks * 64, 1, 1);
    cfg.blockDim = dim3(g.block, 1, 1);
    cudaLaunchAttribute attr[1];
    if (g.ranks > 8) PFL_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g.ranks; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    PFL_CU(cudaOccupancyMaxActiveClusters(&nc, kern, &cfg));
    *out = nc;
    return 0;
  }
  int per_sm = 0;
  PFL_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, g.block, 0));
  *out = g.sync_mode == pf::SYNC_GROUP ? (per_sm * sms) / (int)g.ranks : per_sm * sms;
  return 0;
}

// the latency path (pf_small.cuh): one CTA per filter, N <= 1024, half the CTA produces the random numbers
template <class Model>
int launch_small(const pf::FilterParams& fp, const typename Model::Params& mp, const Geometry& g,
                 uint32_t n_filters, cudaStream_t stream) {
  if constexpr (pf::ModelDyn<Model>::value) {
    return -1;
  } else {
    if (g.P != 2) return -1;
    auto kern = g.block <= 512 ? pf::small_filter_kernel<Model, 2, 512> : pf::small_filter_kernel<Model, 2, 1024>;
    const size_t smem = pf::small_smem_bytes<Model, 2>(fp.N);
    if (smem > 48 * 1024) PFL_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_filters, g.block, smem, stream>>>(fp, mp);
    PFL_CU(cudaGetLastError());
    return 0;
  }
}

// kernel variants: P particles per thread, 64 registers per thread in all of them.
// A 128-register variant (no spills, half the warps) was measured and is 15-40%
// slower on every workload: latency hiding by resident warps matters more than
// the ~50-300 bytes of spills.
template <class Model>
int launch(const pf::FilterParams& fp, const typename Model::Params& mp, const Geometry& g,
           uint32_t n_filters, cudaStream_t stream) {
  if (g.sync_mode == pf::SYNC_SMALL) return launch_small<Model>(fp, mp, g, n_filters, stream);
  const bool small = g.block <= 512;
  switch (g.P) {
    case 2: return small ? launch_V<Model, 2, PF_SMALL_MAXT, PF_SMALL_MINB>(fp, mp, g, n_filters, stream)
                         : launch_V<Model, 2, 1024, 1>(fp, mp, g, n_filters, stream);
    case 4: return small ? launch_V<Model, 4, PF_SMALL_MAXT, PF_SMALL_MINB>(fp, mp, g, n_filters, stream)
                         : launch_V<Model, 4, 1024, 1>(fp, mp, g, n_filters, stream);
    case 8: return small ? launch_V<Model, 8, PF_SMALL_MAXT, PF_SMALL_MINB>(fp, mp, g, n_filters, stream)
                         : launch_V<Model, 8, 1024, 1>(fp, mp, g, n_filters, stream);
  }
  return -1;
}

template <class Model>
int occupancy(const Geometry& g, int* out) {
  const bool small = g.block <= 512;
  switch (g.P) {
    case 2: return small ? occupancy_V<Model, 2, PF_SMALL_MAXT, PF_SMALL_MINB>(g, out) : occupancy_V<Model, 2, 1024, 1>(g, out);
    case 4: return small ? occupancy_V<Model, 4, PF_SMALL_MAXT, PF_SMALL_MINB>(g, out) : occupancy_V<Model, 4, 1024, 1>(g, out);
    case 8: return small ? occupancy_V<Model, 8, PF_SMALL_MAXT, PF_SMALL_MINB>(g, out) : occupancy_V<Model, 8, 1024, 1>(g, out);
  }
  return -1;
}

} // namespace pfl
