// Device-side parameter block shared by tsb_kernels.cu (kernels) and tsb_capi.cu (C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "tsb_plan.h"

namespace tsb {

struct KParams {
  // plan (read-only, built once by tsb_create)
  const TileDesc *tiles;
  const uint4 *idx8;            // [n_tiles*tile_tets] 8 x u16 local vertex ids
  const float *Bsoa;            // [n_tiles][9][tile_tets]
  const int32_t *vlist;
  const float *Xloc;
  const int32_t *dest;
  const uint16_t *ell;
  const int32_t *ell_grp_ptr;
  const int32_t *cg_list;
  const int32_t *need;
  const int32_t *gsv_ptr;
  const int32_t *sv_vid;
  const int32_t *sv_slot_ptr;
  // per-handle scratch
  int32_t *done;                // [n_tiles] arrival counters, self-resetting
  float *scratch;               // [3*n_slots] shared-vertex partials
  float *tile_energy;           // [2*n_tiles] (smooth, barrier) per tile
  uint32_t *energy_counter;     // self-resetting
  // per launch
  const float *x;               // [3n]
  float *grad;                  // [3n] or nullptr (energy only)
  float *energy_out;            // [3]: total, smooth, barrier
  const float *gradH_dev;       // optional device scalar
  float c1, c2, gradH;
  int32_t order;                // 2 or 4
  int32_t laplacian_scale;
  int32_t n_tiles;
};

// Launch the fused kernel.  tile_tets selects the compiled variant.  Returns cudaError_t.
cudaError_t launch_energy_grad(const KParams &p, int tile_tets, int max_local_vertices, cudaStream_t stream);
// One-time per-process attribute setup for a variant (dynamic smem opt-in).  Returns cudaError_t.
cudaError_t prepare_energy_grad(int tile_tets, int max_local_vertices);
bool variant_supported(int tile_tets, int max_local_vertices);
int nvmax_for(int tile_tets);
void set_threads_512(int nt);

cudaError_t launch_scale(const float *g, int64_t count, float gradH, const float *gradH_dev, float *out, cudaStream_t s);
cudaError_t launch_grad_limit(float *g, int64_t count, float thr, float s, float *work2, cudaStream_t st);
cudaError_t launch_adam_uniform(float *p, const float *grad, float *g1, float *g2, int64_t count, float lr,
                                float b1, float b2, int step, float grad_limit, float *work2, cudaStream_t st);

}  // namespace tsb
