// Device-side parameter block shared by tsb_kernels.cu (kernels) and tsb_capi.cu (C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "tsb_plan.h"

namespace tsb {

struct KParams {
  // plan (read-only, built once by tsb_create)
  const unsigned char *vblob;   // n_tiles * vblob_bytes(NV)
  const unsigned char *tblob;   // n_tiles * tblob_bytes(TT)
  const uint16_t *ell;          // gather tables
  const int2 *tile_ell;         // (ell_off, nell) per tile
  // per-handle scratch
  float *scratch;               // [4*n_slots] per-(tile,vertex) partial gradients (float4)
  float *tile_energy;           // [2*n_energy] (smooth, barrier) partials: per tile (v4) or per CTA (pipelined)
  // per launch
  const float *x;               // [3n]
  float *grad;                  // [3n] or nullptr (energy only)
  float *energy_out;            // [3]: total, smooth, barrier
  const float *gradH_dev;       // optional device scalar
  float c1, c2, gradH;
  int32_t order;                // 2 or 4
  int32_t laplacian_scale;
  int32_t n_tiles;
  int32_t n_energy;             // number of (smooth, barrier) partials the tile kernel writes
  int32_t fill;                 // tets per tile upper bound (sizes the tet-blob TMA copy)
  int32_t exp_flags;            // developer A/B switches (0 in production)
};

// Launch the fused kernel.  tile_tets selects the compiled variant.  Returns cudaError_t.
cudaError_t launch_energy_grad(const KParams &p, int tile_tets, int n_vertices, const int32_t *slot_ptr, cudaStream_t stream);
// One-time per-process attribute setup for a variant (dynamic smem opt-in).  Returns cudaError_t.
cudaError_t prepare_energy_grad(int tile_tets);
int nvmax_for(int tile_tets);     // staged-vertex capacity of the compiled variant (0 = not compiled)

void set_threads_512(int nt);
void set_skip_combine(int v);
void set_pdl_tile(int v);
void set_exp_flags(int v);

cudaError_t launch_scale(const float *g, int64_t count, float gradH, const float *gradH_dev, float *out, cudaStream_t s);
cudaError_t launch_grad_limit(float *g, int64_t count, float thr, float s, float *work2, cudaStream_t st);
cudaError_t launch_adam_uniform(float *p, const float *grad, float *g1, float *g2, int64_t count, double lr,
                                double b1, double b2, int step, double grad_limit, float *work2, cudaStream_t st);

}  // namespace tsb
