// kernels.hpp — host-callable launchers of the kernels in kernels.cu / wavefront.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_scene.cuh"

namespace rt {

cudaError_t launch_trace_rays(cudaStream_t st, const DScene &sc, int mode, uint64_t n, const double *o, const double *d,
                              void *hits, DCounters *ctr);
cudaError_t launch_trace_primary(cudaStream_t st, const DScene &sc, int mode, double jx, double jy, void *hits,
                                 double *rays, DCounters *ctr);
cudaError_t launch_light_pdf(cudaStream_t st, const DScene &sc, int mode, uint64_t n, const double *pos, const double *dir,
                             double *pdf, DCounters *ctr);
cudaError_t launch_render_mega(cudaStream_t st, const DScene &sc, int mode, uint32_t sample_begin, uint32_t sample_count,
                               uint64_t seed, double *accum, DCounters *ctr);
cudaError_t launch_postprocess(cudaStream_t st, uint64_t n, const double *accum, uint32_t samples, double *rgb);
cudaError_t launch_rng_probe(cudaStream_t st, uint64_t seed, uint32_t pixel, uint32_t sample, uint64_t n,
                             unsigned long long *u64_out, double *probe8);

}  // namespace rt
