// sfw_kernels.h — launch wrappers of the scorer kernels (sfw_kernels.cu), called by sfw_abi.cu.
#ifndef SFW_KERNELS_H
#define SFW_KERNELS_H
#include <cuda.h>
#include <cuda_runtime.h>

#include "sfw_dev.h"

// dynamic shared memory of sfw_score_small for a block of T threads
size_t sfw_small_smem_bytes(uint32_t win_wp, uint32_t win_h, uint32_t P, uint32_t M, uint32_t F, uint32_t T);
cudaError_t sfw_small_max_dynamic_smem(size_t *bytes);
const char *sfw_small_kernel_name(uint32_t T, bool share = false); // variant a block of T threads dispatches to
cudaError_t sfw_small_occupancy(uint32_t T, size_t smem_bytes, int *blocks_per_sm);
cudaError_t sfw_warp_paths_occupancy(size_t smem_bytes, int *blocks_per_sm);
// dependent: programmatic dependent launch behind the previous kernel of the stream (prefix sharing: the prologue of
// a launch overlaps the tail of the path launch it continues from)
cudaError_t sfw_launch_warp_paths(const SfwBatchDev &B, const CUtensorMap &tmap, size_t smem_bytes,
                                  cudaStream_t stream, bool dependent = false);
cudaError_t sfw_launch_small(const SfwBatchDev &B, const CUtensorMap &tmap, uint32_t T,
                             size_t smem_bytes, cudaStream_t stream, bool dependent = false);
// block-per-trajectory kernel for dense crowds (sfw_crowd.cu) + stand-alone arg-min
// (threads: SFW_CROWD_THREADS or SFW_CROWD_THREADS_SMALL — the two instantiations of the kernel)
size_t sfw_crowd_smem_bytes(uint32_t P, uint32_t M, uint32_t F, uint32_t S, uint32_t threads = SFW_CROWD_THREADS);
cudaError_t sfw_crowd_prepare(uint32_t threads, size_t smem_bytes, int *blocks_per_sm);
bool sfw_crowd_fuses_argmin(const SfwBatchDev &B); // the scorer's last block reduces the winners itself
cudaError_t sfw_launch_crowd(const SfwBatchDev &B, unsigned int *work_counter, uint32_t grid, uint32_t threads,
                             size_t smem_bytes, cudaStream_t stream, bool with_argmin);
cudaError_t sfw_launch_export_invalid(const SfwExchangeDev &X, uint32_t n_scenes, cudaStream_t stream); // sfw_exchange.cu
cudaError_t sfw_launch_points(const SfwBatchDev &B, uint32_t scene, uint32_t idx, uint32_t n_points,
                              double *out_xyz, cudaStream_t stream);
cudaError_t sfw_launch_marker_points(const SfwBatchDev &B, uint32_t scene, uint32_t first, uint32_t stride,
                                     uint32_t count, uint32_t max_points, double *out_xyz, uint16_t *out_n,
                                     cudaStream_t stream);
cudaError_t sfw_launch_may_i_stop(const SfwBatchDev &B, uint32_t scene, double vl_x, double vl_y, double va, double x,
                                  double y, double th, double dt, int *out, cudaStream_t stream);
#endif
