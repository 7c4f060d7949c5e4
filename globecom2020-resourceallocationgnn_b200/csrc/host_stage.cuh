// Host staging of strided caller memory into pinned buffers + pipelined H2D (see host_stage.cu).
#pragma once
#include "../../include/v2v_gnn.h"
#include "v2v_common.cuh"

namespace v2v {

constexpr int kHostStageMaxTensors = 8;
constexpr int kCheckBinary = 1, kCheckNonzero = 2;       // HostStageTensor::check
constexpr int kFlagNonbinary = 1, kFlagNonzero = 2;      // result bits

struct HostStageTensor {
  const v2v_host_view* views;
  int n_views;
  float* pinned;         // destination (pinned host memory)
  float* device;         // H2D target (nullptr: stage only)
  size_t bytes;          // bytes to copy to the device once all views have landed
  const float* h2d_from; // source of that copy (nullptr: `pinned`); lets the last of several adjacent tensors ship all of them at once
  int check;             // 0, kCheckBinary (values outside {0,1}?) or kCheckNonzero (any non-zero value?)
  int copy_if;           // 0: always copy to the device; else only if (flags & copy_if)
  // adjacency tensors [B][N][N] with N <= 32: the workers also build the bit masks (adj_pack_kernel's job) in pinned
  // memory, and those travel instead of the 32x larger dense matrix
  int pack_N;
  uint32_t* pin_in_mask; uint32_t* pin_out_mask;
  uint32_t* dev_in_mask; uint32_t* dev_out_mask;
  size_t mask_bytes;
};

int host_stage_threads();
int host_stage_validate(const v2v_host_view* views, int n, long dst_elems, const char* what);
// Gathers every tensor's views into its pinned buffer on the worker pool and enqueues the H2D copies on `st` in order.
// flags_out[t] (optional) receives the kFlag* bits found in tensor t.
int host_stage_run(const HostStageTensor* tensors, int n_tensors, cudaStream_t st, int* flags_out);

}  // namespace v2v
