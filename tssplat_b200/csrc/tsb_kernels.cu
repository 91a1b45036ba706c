// Fused geometry-energy + gradient kernel for sm_100a, plus the small level-1 helpers.
//
// One launch replaces the reference's forward+backward pipeline
// (tssplat_ext/tet_spheres/tet_spheres_cuda.cu:118-263: SpMV GTLTLG.x, Sdot, SpMV G.x,
// cuda_forward_det, Sasum, SpMV c1.GTLTLG.x, SpMV G.x again, cuda_backward_det, SpMV G^T, Sscal,
// with three host syncs).
//
// Math.  For tet t with own vertices v0..v3, rest inverse B = Dm^-1 (rows
// a1,a2,a3 are the rest gradients of the hat functions of v1..v3, a0 = -(a1+a2+a3)):
//     F_t = sum_k x_vk (x) a_k                                (geometry/mesh_utils.py:38-69)
// The reference's smoothness term 1/2 x^T G^T L^T L G x equals 1/2 sum_t ||H_t||^2 with
// H_t = (L F)_t = deg_t F_t - sum_{s face-nbr t} F_s.  Two tets sharing a face agree on that face,
// so F_s - F_t is rank one:  F_s - F_t = d_k (x) a_k / lambda_kk  where o_k is the vertex of s
// opposite the shared face k, lambda_k. are the barycentric coordinates of REST(o_k) in t and
//     d_k = x_ok - sum_j lambda_kj x_vj          (how far o_k is from t's affine map)
// Hence  H_t = sum_k rho_k d_k (x) a_k,  rho_k = -1/lambda_kk > 0:  an 8-vertex stencil per tet
// whose gradient scatters to those same 8 vertices -- no neighbour-tet intermediates, no 2-ring
// passes, no grid sync.  The barrier term is the reference's: max(-det F,0)^p, p in {2,4}
// (cu:48-66), gradient -p(-J)^(p-1) cof(F) (cu:68-102).
//
// Execution.  Persistent CTAs (2 per SM, 256 threads) loop over tiles of <= TT tets / <= NV staged
// vertices.  One thread issues TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of each tile's
// vertex blob, tet blob and gather table from global to shared memory one tile ahead; the only
// dependent global chain (vertex id -> x) is issued one phase ahead and lands in registers.
//   phase 0  x (registers) + rest X (staged) -> float4 array in shared memory
//   phase 1  one tet per thread, branch-free: 13 smem gathers, energy terms, 8 output 3-vectors
//            written to a [24][TT+4] smem table (conflict-free stores)
//   phase 2  one gather row per thread (<= 16 entries, bank-aware order, padding -> zero column):
//            sum the table entries and store the partial gradient to the row's float4 scratch slot
// A second, tiny kernel chained with programmatic dependent launch (griddepcontrol) sums each
// vertex's slots in fixed order into grad (scaled by gradH) and folds the per-CTA energies in
// fp64.  No atomics, no fences, no grid sync: bitwise deterministic.
// History (profiles/r01_summary.md): a single-launch last-arriver combine cost 47% of warp time in
// fences/atomics/barriers; a warp-specialised pipeline that overlapped the phases lost to shared-
// memory contention (row gather 1.5k -> 4.3k cycles when it overlaps the tet math); overlapping
// only the global-memory latency (this design) won.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "tsb_kernels.cuh"

namespace tsb {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + TMA bulk copy (1-D) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

constexpr int align_up(int v, int a) { return (v + a - 1) / a * a; }
int g_skip_combine = 0;   // developer switch (timing experiments only)
int g_pdl_tile = 1;       // tile kernel launched with programmatic stream serialisation
int g_exp_flags = 0;



template <int TT, int NV>
struct Smem {
  static constexpr int NR = NV + 8 * TT / kRowCap;           // == rows_cap(TT, NV)
  static constexpr int ELLCAP = 8 * TT + 32 * kRowCap + NR + 64;   // == ell_cap(TT, NV)
  static constexpr int kVBytes = 64 + 16 * NV + 4 * NR + 4 * (NR / 32 + 4);
  static constexpr int kVStage = align_up(kVBytes, 128);     // two vertex-blob stages
  static constexpr int kTOff = 2 * kVStage;
  static constexpr int kEllOff = align_up(kTOff + 52 * TT, 128);
  static constexpr int kXs4Off = align_up(kEllOff + 2 * ELLCAP, 128);
  static constexpr int kOutOff = align_up(kXs4Off + 16 * NV, 128);
  static constexpr int kTTP = TT + 4;                 // output-table row stride; column TT holds zeros
  static constexpr int kBytes = kOutOff + 96 * kTTP;
};

// Sum one vertex's table entries.  ep points at this lane's first (entry0, entry1) pair; pairs of
// successive k are 32 words apart.  Padding entries point at the table's zero column.
template <int TTP>
__device__ __forceinline__ void gather_vertex(const float *outb, const uint32_t *ep, int len2, float &g0, float &g1, float &g2) {
#pragma unroll 8
  for (int k = 0; k < len2; ++k) {
    const uint32_t pr = ep[k * 32];
    const float *o0 = outb + (pr & 0xffffu), *o1 = outb + (pr >> 16);
    g0 += o0[0]; g1 += o0[TTP]; g2 += o0[2 * TTP];
    g0 += o1[0]; g1 += o1[TTP]; g2 += o1[2 * TTP];
  }
}

// One tet: 13 shared-memory gathers, barrier + smoothness energy terms, and (WITH_GRAD) its 8 output
// 3-vectors written to column `lt` of the [24][TTP] table.  Shared by both tile kernels.
template <int TTP, bool WITH_GRAD>
__device__ __forceinline__ void tet_body(const int lt, const uint4 *__restrict__ idx_s, const float *__restrict__ B_s,
                                         const float4 *__restrict__ xs4, const float2 *__restrict__ xs2,
                                         float *__restrict__ outb, const float c1, const float c2, const int order,
                                         const bool lscale, float &es, float &eb) {
  const uint4 iv = idx_s[lt];
  float b[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) b[i] = B_s[lt * 9 + i];
  const unsigned iown[4] = {iv.x & 0xffffu, iv.x >> 16, iv.y & 0xffffu, iv.y >> 16};
  const unsigned ioppr[4] = {iv.z & 0xffffu, iv.z >> 16, iv.w & 0xffffu, iv.w >> 16};

  const float4 p0 = xs4[iown[0]];
  const float2 q0 = xs2[iown[0]];
  float e[3][3];  // e[j][r] = x_{v_{j+1}}[r] - x_{v0}[r]
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float4 pj = xs4[iown[j + 1]];
    e[j][0] = pj.x - p0.x; e[j][1] = pj.y - p0.y; e[j][2] = pj.z - p0.z;
  }
  float4 po[4];
  float2 qo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { po[k] = xs4[ioppr[k] & 0x7fffu]; qo[k] = xs2[ioppr[k] & 0x7fffu]; }

  // hat gradients: a[0] = -(a1+a2+a3), a[j] = row j-1 of B
  float a[4][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    a[1][c] = b[c]; a[2][c] = b[3 + c]; a[3][c] = b[6 + c];
    a[0][c] = -(b[c] + b[3 + c] + b[6 + c]);
  }

  float z[3][3];  // gradient contributions to own vertices 1..3 (vertex 0 follows from momentum)
#pragma unroll
  for (int j = 0; j < 3; ++j) { z[j][0] = 0.f; z[j][1] = 0.f; z[j][2] = 0.f; }
  {
    float F[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) F[r][c] = e[0][r] * a[1][c] + e[1][r] * a[2][c] + e[2][r] * a[3][c];
    // cofactors = d det / dF  (tet_spheres_cuda.cu:32-46)
    const float C00 = F[1][1] * F[2][2] - F[1][2] * F[2][1];
    const float C01 = F[1][2] * F[2][0] - F[1][0] * F[2][2];
    const float C02 = F[1][0] * F[2][1] - F[1][1] * F[2][0];
    const float J = F[0][0] * C00 + F[0][1] * C01 + F[0][2] * C02;
    if (J < 0.f) {   // rare: inverted tet
      const float m = -J;
      float coef;
      if (order == 2) { eb += m * m; coef = 2.f * m; }
      else { const float m2 = m * m; eb += m2 * m2; coef = 4.f * m2 * m; }
      if (WITH_GRAD) {
        float C[3][3];
        C[0][0] = C00; C[0][1] = C01; C[0][2] = C02;
        C[1][0] = F[0][2] * F[2][1] - F[0][1] * F[2][2];
        C[1][1] = F[0][0] * F[2][2] - F[0][2] * F[2][0];
        C[1][2] = F[0][1] * F[2][0] - F[0][0] * F[2][1];
        C[2][0] = F[0][1] * F[1][2] - F[0][2] * F[1][1];
        C[2][1] = F[0][2] * F[1][0] - F[0][0] * F[1][2];
        C[2][2] = F[0][0] * F[1][1] - F[0][1] * F[1][0];
        const float pc = -c2 * coef;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float P0 = pc * C[r][0], P1 = pc * C[r][1], P2 = pc * C[r][2];
#pragma unroll
          for (int j = 0; j < 3; ++j) z[j][r] = P0 * a[j + 1][0] + P1 * a[j + 1][1] + P2 * a[j + 1][2];
        }
      }
    }
  }

  // smoothness stencil: H = w * sum_k rho_k d_k (x) a_k      (branch-free; boundary faces have
  // rho = 0 and gather the tet's own vertex)
  float H[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) { H[r][0] = 0.f; H[r][1] = 0.f; H[r][2] = 0.f; }
  float lam[4][3], rho[4];
  int deg = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool valid = (ioppr[k] & 0x8000u) != 0u;
    deg += valid ? 1 : 0;
    const float rx = po[k].w - p0.w, ry = qo[k].x - q0.x, rz = qo[k].y - q0.y;
    const float l1 = b[0] * rx + b[1] * ry + b[2] * rz;
    const float l2 = b[3] * rx + b[4] * ry + b[5] * rz;
    const float l3 = b[6] * rx + b[7] * ry + b[8] * rz;
    const float lkk = (k == 0) ? (1.f - l1 - l2 - l3) : (k == 1 ? l1 : (k == 2 ? l2 : l3));
    const float rk = valid ? __fdividef(-1.f, lkk) : 0.f;
    lam[k][0] = l1; lam[k][1] = l2; lam[k][2] = l3; rho[k] = rk;
    const float dx = (po[k].x - p0.x) - l1 * e[0][0] - l2 * e[1][0] - l3 * e[2][0];
    const float dy = (po[k].y - p0.y) - l1 * e[0][1] - l2 * e[1][1] - l3 * e[2][1];
    const float dz = (po[k].z - p0.z) - l1 * e[0][2] - l2 * e[1][2] - l3 * e[2][2];
    const float sx = rk * dx, sy = rk * dy, sz = rk * dz;
    H[0][0] += sx * a[k][0]; H[0][1] += sx * a[k][1]; H[0][2] += sx * a[k][2];
    H[1][0] += sy * a[k][0]; H[1][1] += sy * a[k][1]; H[1][2] += sy * a[k][2];
    H[2][0] += sz * a[k][0]; H[2][1] += sz * a[k][1]; H[2][2] += sz * a[k][2];
  }
  const float w = (lscale && deg > 0) ? __fdividef(1.f, float(deg)) : 1.f;
  float hh = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { H[r][c] *= w; hh += H[r][c] * H[r][c]; }
  es += 0.5f * hh;

  if (WITH_GRAD) {
    const float cw = c1 * w;
    float ys[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sk = cw * rho[k];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float y = sk * (H[r][0] * a[k][0] + H[r][1] * a[k][1] + H[r][2] * a[k][2]);
        outb[((4 + k) * 3 + r) * TTP + lt] = y;
        ys[r] += y;
        z[0][r] -= lam[k][0] * y; z[1][r] -= lam[k][1] * y; z[2][r] -= lam[k][2] * y;
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      outb[(0 * 3 + r) * TTP + lt] = -(z[0][r] + z[1][r] + z[2][r] + ys[r]);   // translation invariance
      outb[(1 * 3 + r) * TTP + lt] = z[0][r];
      outb[(2 * 3 + r) * TTP + lt] = z[1][r];
      outb[(3 * 3 + r) * TTP + lt] = z[2][r];
    }
  }
}

// Tile kernel: persistent CTAs (MINB per SM), each looping over tiles b, b+G, b+2G, ...  The three
// phases of a tile run in lock-step inside the CTA (they are all shared-memory heavy, so overlapping
// them only makes them contend -- profiles/r01_summary.md), but every global access of tile j+1 is
// in flight while tile j computes:
//   * vertex blob j+1 (double buffered) is requested by TMA as soon as tile j-1 has been consumed,
//   * tet blob j+1 the moment the tet math of tile j is done, the gather table of tile j right after
//     the row gather of tile j-1,
//   * the only dependent chain, vertex id -> x, is issued straight from global memory one phase ahead
//     (ids at the top of tile j, x right after its tet math) and lands in registers.
template <int TT, int NV, int NT, int MINB, bool WITH_GRAD>
__global__ void __launch_bounds__(NT, MINB) energy_grad_kernel(const __grid_constant__ KParams p) {
  using L = Smem<TT, NV>;
  constexpr int NR = L::NR, TTP = L::kTTP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint4 *idx_s = reinterpret_cast<const uint4 *>(smem_raw + L::kTOff);
  const float *B_s = reinterpret_cast<const float *>(smem_raw + L::kTOff + 16 * TT);
  const uint16_t *ell_s = reinterpret_cast<const uint16_t *>(smem_raw + L::kEllOff);
  float4 *xs4 = reinterpret_cast<float4 *>(smem_raw + L::kXs4Off);                               // x, y, z, X
  float *outb = reinterpret_cast<float *>(smem_raw + L::kOutOff);
  __shared__ __align__(8) uint64_t bar_v[2], bar_t, bar_e;
  __shared__ float s_red[2 * (NT / 32)];

  const int tid = threadIdx.x;
  const int G = int(gridDim.x);
  const int n_my = (p.n_tiles - int(blockIdx.x) + G - 1) / G;
  const uint32_t nt_b = uint32_t(p.fill);
  constexpr int kVPer = (NV + NT - 1) / NT;

  auto issue_v = [&](int j) {
    const int tile = int(blockIdx.x) + j * G;
    mbar_expect_tx(&bar_v[j & 1], L::kVBytes);
    bulk_g2s(smem_raw + (j & 1) * L::kVStage, p.vblob + size_t(tile) * L::kVBytes, L::kVBytes, &bar_v[j & 1]);
  };
  auto issue_t = [&](int j) {
    const int tile = int(blockIdx.x) + j * G;
    mbar_expect_tx(&bar_t, 52u * nt_b);
    const unsigned char *tb = p.tblob + size_t(tile) * (52 * TT);
    bulk_g2s(smem_raw + L::kTOff, tb, 16u * nt_b, &bar_t);
    bulk_g2s(smem_raw + L::kTOff + 16 * TT, tb + 16 * TT, 36u * nt_b, &bar_t);
  };
  auto issue_e = [&](const int2 el) {
    mbar_expect_tx(&bar_e, 2u * uint32_t(el.y));
    if (el.y > 0) bulk_g2s(smem_raw + L::kEllOff, p.ell + el.x, 2u * uint32_t(el.y), &bar_e);
  };
  // TMA is issued by the LAST thread: its warp processes the fewest tets of a tile, so the issue
  // latency stays off the critical path of the tet math.
  const bool issuer = (tid == NT - 1);
  auto load_vids = [&](int j, int (&vid)[kVPer]) {     // entries past nvert are zero padding (-> x[0])
    const int32_t *vl_g = reinterpret_cast<const int32_t *>(p.vblob + size_t(int(blockIdx.x) + j * G) * L::kVBytes + 64);
#pragma unroll
    for (int q = 0; q < kVPer; ++q) vid[q] = (tid + q * NT < NV) ? __ldg(vl_g + tid + q * NT) : 0;
  };
  auto load_x = [&](const int (&vid)[kVPer], float (&px)[kVPer][3]) {
#pragma unroll
    for (int q = 0; q < kVPer; ++q) {
      const float *xp = p.x + 3 * size_t(vid[q]);
      px[q][0] = __ldg(xp); px[q][1] = __ldg(xp + 1); px[q][2] = __ldg(xp + 2);
    }
  };

  int vid[kVPer];
  float px[kVPer][3];
  if (n_my > 0) load_vids(0, vid);
  if (issuer) {
    mbar_init(&bar_v[0], 1); mbar_init(&bar_v[1], 1); mbar_init(&bar_t, 1); mbar_init(&bar_e, 1);
    mbar_fence_init();
    if (n_my > 0) { issue_v(0); issue_t(0); }
  }
  if (WITH_GRAD && tid < 3) outb[tid * TTP + TT] = 0.f;   // zero column for gather-table padding
  // Everything above touches only plan data.  x (may have been updated by the optimiser), the scratch
  // slots and the energy partials belong to the previous kernels in the stream: wait for them here.
  // (The kernel is launched with programmatic stream serialisation, so this prologue overlaps the
  // tail of the previous combine kernel.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (n_my > 0) load_x(vid, px);
  __syncthreads();

  float es = 0.f, eb = 0.f;
  const float c1 = p.c1, c2 = p.c2;
  const int order = p.order;
  const bool lscale = p.laplacian_scale != 0;
  for (int j = 0; j < n_my; ++j) {
    const unsigned char *vb = smem_raw + (j & 1) * L::kVStage;
    const TileHeader *hd = reinterpret_cast<const TileHeader *>(vb);
    const float *Xx_s = reinterpret_cast<const float *>(vb + 64 + 4 * NV);
    const float2 *xs2 = reinterpret_cast<const float2 *>(vb + 64 + 8 * NV);                      // (Y, Z) rest
    const int32_t *slot_s = reinterpret_cast<const int32_t *>(vb + 64 + 16 * NV);
    const int32_t *grp_s = reinterpret_cast<const int32_t *>(vb + 64 + 16 * NV + 4 * NR);
    if (j + 1 < n_my) load_vids(j + 1, vid);           // ids of the next tile: needed only after this tile's tet math
    int2 el = make_int2(0, 0);
    if (WITH_GRAD && issuer) el = __ldg(p.tile_ell + int(blockIdx.x) + j * G);   // consumed after (A): latency hidden

    // ---------------- phase 0: x (already in registers) + rest X -> shared -------------------
    mbar_wait(&bar_v[j & 1], (j >> 1) & 1);
    const int ntet = hd->ntet, nvert = hd->nvert, nrow = hd->nrow;
#pragma unroll
    for (int q = 0; q < kVPer; ++q) {
      const int i = tid + q * NT;
      if (i < nvert) xs4[i] = make_float4(px[q][0], px[q][1], px[q][2], Xx_s[i]);
    }
    __syncthreads();     // (A) xs4 complete; every thread has left the previous tile's row gather
    if (issuer) {
      if (j + 1 < n_my) issue_v(j + 1);               // its stage held tile j-1, now fully consumed
      if (WITH_GRAD) issue_e(el);                      // gather-table buffer is free since (A)
    }

    // ---------------- phase 1: tets -----------------------------------------------------------------
    mbar_wait(&bar_t, j & 1);
    for (int lt = tid; lt < ntet; lt += NT)
      tet_body<TTP, WITH_GRAD>(lt, idx_s, B_s, xs4, xs2, outb, c1, c2, order, lscale, es, eb);
    __syncthreads();     // (B) output table complete; tet blob and xs4 are free
    if (j + 1 < n_my) {
      if (issuer) issue_t(j + 1);
      load_x(vid, px);                                 // lands while the row gather below runs
    } else {
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // last tile: let the combine kernel launch
    }

    // ---------------- phase 2: per-row gather --------------------------------------------------------
    if (WITH_GRAD) {
      mbar_wait(&bar_e, j & 1);
      float4 *scratch4 = reinterpret_cast<float4 *>(p.scratch);
      for (int r = tid; r < nrow; r += NT) {
        const int g = r >> 5, lane = r & 31;
        const int beg = grp_s[g], end = grp_s[g + 1];
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        gather_vertex<TTP>(outb, reinterpret_cast<const uint32_t *>(ell_s + beg) + lane, (end - beg) >> 6, g0, g1, g2);
        scratch4[slot_s[r]] = make_float4(g0, g1, g2, 0.f);
      }
    }
  }

  // one (smooth, barrier) partial per CTA (tree reduction; the cross-CTA sum is done in fp64 by the combine kernel)
  {
    const float ws = warp_sum(es), wb = warp_sum(eb);
    if ((tid & 31) == 0) { s_red[tid >> 5] = ws; s_red[NT / 32 + (tid >> 5)] = wb; }
  }
  __syncthreads();
  if (tid < 32) {
    float vs = (tid < NT / 32) ? s_red[tid] : 0.f, vb2 = (tid < NT / 32) ? s_red[NT / 32 + tid] : 0.f;
    vs = warp_sum(vs); vb2 = warp_sum(vb2);
    if (tid == 0) { p.tile_energy[2 * blockIdx.x] = vs; p.tile_energy[2 * blockIdx.x + 1] = vb2; }
  }
  if (n_my == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Combine kernel: grad[v] = gradH * sum of v's scratch slots (fixed order); block 0 also folds the
// per-tile energies in fp64.  Launched with programmatic stream serialization right behind the
// tile kernel; griddepcontrol.wait blocks until that grid has completed and flushed.
template <int NT>
__global__ void __launch_bounds__(NT) combine_kernel(const __grid_constant__ KParams p, int n_vertices, const int32_t *__restrict__ slot_ptr) {
  const int tid = threadIdx.x;
  const int v = blockIdx.x * NT + tid;
  int s0 = 0, s1 = 0;
  if (v < n_vertices) { s0 = __ldg(slot_ptr + v); s1 = __ldg(slot_ptr + v + 1); }   // plan data: safe before the wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next tile kernel may start its prologue
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float gh = p.gradH;
  if (p.gradH_dev) gh *= __ldg(p.gradH_dev);   // produced by earlier kernels in the stream: read after the wait
  if (v < n_vertices) {
    const float4 *scratch4 = reinterpret_cast<const float4 *>(p.scratch);
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    // slots are summed in index order (deterministic); loads are issued 4 at a time so the L2 round
    // trips overlap instead of forming a dependent chain
    for (int s = s0; s < s1; s += 4) {
      float4 q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] = (s + j < s1) ? __ldcg(scratch4 + s + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) { g0 += q[j].x; g1 += q[j].y; g2 += q[j].z; }
    }
    float *gp = p.grad + 3 * size_t(v);
    gp[0] = gh * g0; gp[1] = gh * g1; gp[2] = gh * g2;
  }
  if (blockIdx.x == 0) {
    double as = 0.0, ab = 0.0;
    for (int t = tid; t < p.n_energy; t += NT) { as += double(__ldcg(p.tile_energy + 2 * t)); ab += double(__ldcg(p.tile_energy + 2 * t + 1)); }
    as = warp_sum(as); ab = warp_sum(ab);
    __shared__ double s_dred[2 * (NT / 32)];
    if ((tid & 31) == 0) { s_dred[tid >> 5] = as; s_dred[NT / 32 + (tid >> 5)] = ab; }
    __syncthreads();
    if (tid == 0) {
      double ts = 0.0, tb = 0.0;
      for (int wgt = 0; wgt < NT / 32; ++wgt) { ts += s_dred[wgt]; tb += s_dred[NT / 32 + wgt]; }
      p.energy_out[0] = float(double(p.c1) * ts + double(p.c2) * tb);
      p.energy_out[1] = float(ts);
      p.energy_out[2] = float(tb);
    }
  }
}

// ---- level-1 helpers -------------------------------------------------------------------------------
__global__ void scale_kernel(const float *__restrict__ g, int64_t count, float gradH, const float *gradH_dev,
                             float *__restrict__ out) {
  const float s = gradH * (gradH_dev ? __ldg(gradH_dev) : 1.f);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = s * g[i];
}

__device__ __forceinline__ void block_max_to(float v, float *dst) {
  // v >= 0.  Order-preserving uint compare for non-negative floats.
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __shared__ float s_m[32];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float m = (threadIdx.x < (blockDim.x + 31) / 32) ? s_m[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned int *>(dst), __float_as_uint(m));
  }
  __syncthreads();
}

__global__ void absmax_kernel(const float *__restrict__ g, int64_t count, float *work) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(g[i]));
  block_max_to(m, work);
}

// if max|g| > thr: g *= s / max|g|.  Consumes and re-zeroes work[0] through a ticket in work[1].
__global__ void grad_limit_apply_kernel(float *g, int64_t count, float thr, float s, float *work) {
  const float m = __ldcg(work);
  if (m > thr) {
    const float f = s / m;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) g[i] *= f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int *ticket = reinterpret_cast<unsigned int *>(work + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) { work[0] = 0.f; *ticket = 0u; }
  }
}

// AdamUniform (utils/optimizer.py:37-89), pass 1: moments + the two global maxima.
__global__ void adam_uniform_moments_kernel(const float *__restrict__ grad, float *__restrict__ g1, float *__restrict__ g2,
                                            int64_t count, float b1, float b2, float omb1, float omb2, float inv_bc1,
                                            float inv_bc2, float *work) {
  float mx2 = 0.f, mx1 = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const float g = grad[i];
    const float m1 = b1 * g1[i] + omb1 * g;                 // optimizer.py:61  (1 - beta formed in double on the host, like Python)
    const float m2 = b2 * g2[i] + omb2 * (g * g);           // optimizer.py:62
    g1[i] = m1; g2[i] = m2;
    mx2 = fmaxf(mx2, sqrtf(m2 * inv_bc2));                  // optimizer.py:68,74
    mx1 = fmaxf(mx1, fabsf(m1 * inv_bc1));                  // optimizer.py:67,83
  }
  block_max_to(mx2, work);
  block_max_to(mx1, work + 1);
}

// pass 2: p -= lr * clamp(m1_hat / (1e-8 + max sqrt(m2_hat)))   (optimizer.py:74-88)
__global__ void adam_uniform_apply_kernel(float *__restrict__ p, const float *__restrict__ g1, int64_t count, float lr,
                                          float inv_bc1, float grad_limit, float *work, unsigned int *ticket) {
  const float denom = 1e-8f + __ldcg(work);
  float f = inv_bc1 / denom;
  if (grad_limit > 0.f) {
    const float s = __ldcg(work + 1) / denom;               // max |gr|
    if (s > grad_limit) f *= grad_limit / s;
  }
  f *= lr;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) p[i] -= f * g1[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) { work[0] = 0.f; work[1] = 0.f; *ticket = 0u; }
  }
}

// Compiled variants: tile capacity TT -> staged-vertex capacity NV (then threads, min CTAs/SM)
#define TSB_V256 256, 256
#define TSB_V512 512, 256
#define TSB_V1024 1024, 640

cudaError_t launch_combine(const KParams &p, int n_vertices, const int32_t *slot_ptr, cudaStream_t stream);
int g_num_sms = 148;

template <int TT, int NV, int NT, int MINB>
cudaError_t launch_variant(const KParams &p0, int n_vertices, const int32_t *slot_ptr, cudaStream_t stream) {
  KParams p = p0;
  p.exp_flags = g_exp_flags;
  const int slots = MINB * g_num_sms;
  const int grid = p.n_tiles < slots ? p.n_tiles : slots;
  p.n_energy = grid;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(grid));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = Smem<TT, NV>::kBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl_tile ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = p.grad ? cudaLaunchKernelEx(&cfg, energy_grad_kernel<TT, NV, NT, MINB, true>, p)
                         : cudaLaunchKernelEx(&cfg, energy_grad_kernel<TT, NV, NT, MINB, false>, p);
  if (e != cudaSuccess || g_skip_combine) return e;
  return launch_combine(p, n_vertices, slot_ptr, stream);
}

// combine kernel, chained with programmatic dependent launch
cudaError_t launch_combine(const KParams &p, int n_vertices, const int32_t *slot_ptr, cudaStream_t stream) {
  constexpr int CNT = 256;
  const int nv = p.grad ? n_vertices : 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(nv > 0 ? (nv + CNT - 1) / CNT : 1));
  cfg.blockDim = dim3(CNT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, combine_kernel<CNT>, p, nv, slot_ptr);
}

template <int TT, int NV, int NT, int MINB>
cudaError_t prepare_variant() {
  const int smem = Smem<TT, NV>::kBytes;
  cudaError_t e = cudaFuncSetAttribute(energy_grad_kernel<TT, NV, NT, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(energy_grad_kernel<TT, NV, NT, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

inline int grid_for(int64_t count, int block) {
  int64_t g = (count + block - 1) / block;
  return int(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

int nvmax_for(int tile_tets) {
  switch (tile_tets) {
    case 256: return 256;
    case 512: return 256;
    case 1024: return 640;
  }
  return 0;
}

static int g_threads_512 = 256;

cudaError_t prepare_energy_grad(int tile_tets) {
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_num_sms = sms;
  }
  switch (tile_tets) {
    case 256: return prepare_variant<TSB_V256, 256, 3>();
    case 512: {
      cudaError_t e = prepare_variant<TSB_V512, 256, 2>();
      if (e != cudaSuccess) return e;
      return prepare_variant<TSB_V512, 512, 1>();
    }
    case 1024: return prepare_variant<TSB_V1024, 512, 1>();
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_energy_grad(const KParams &p, int tile_tets, int n_vertices, const int32_t *slot_ptr, cudaStream_t stream) {
  switch (tile_tets) {
    case 256: return launch_variant<TSB_V256, 256, 3>(p, n_vertices, slot_ptr, stream);
    case 512:
      return g_threads_512 == 512 ? launch_variant<TSB_V512, 512, 1>(p, n_vertices, slot_ptr, stream) : launch_variant<TSB_V512, 256, 2>(p, n_vertices, slot_ptr, stream);
    case 1024: return launch_variant<TSB_V1024, 512, 1>(p, n_vertices, slot_ptr, stream);
  }
  return cudaErrorInvalidValue;
}

void set_threads_512(int nt) { g_threads_512 = (nt == 512) ? 512 : 256; }
void set_skip_combine(int v) { g_skip_combine = v; }
void set_pdl_tile(int v) { g_pdl_tile = v; }
void set_exp_flags(int v) { g_exp_flags = v; }

cudaError_t launch_scale(const float *g, int64_t count, float gradH, const float *gradH_dev, float *out, cudaStream_t s) {
  scale_kernel<<<grid_for(count, 256), 256, 0, s>>>(g, count, gradH, gradH_dev, out);
  return cudaGetLastError();
}

cudaError_t launch_grad_limit(float *g, int64_t count, float thr, float s, float *work2, cudaStream_t st) {
  const int grid = grid_for(count, 256);
  absmax_kernel<<<grid, 256, 0, st>>>(g, count, work2);
  grad_limit_apply_kernel<<<grid, 256, 0, st>>>(g, count, thr, s, work2);
  return cudaGetLastError();
}

cudaError_t launch_adam_uniform(float *p, const float *grad, float *g1, float *g2, int64_t count, double lr, double b1,
                                double b2, int step, double grad_limit, float *work, cudaStream_t st) {
  const float inv_bc1 = float(1.0 / (1.0 - pow(b1, double(step))));   // optimizer.py:67
  const float inv_bc2 = float(1.0 / (1.0 - pow(b2, double(step))));   // optimizer.py:68
  const int grid = grid_for(count, 256);
  adam_uniform_moments_kernel<<<grid, 256, 0, st>>>(grad, g1, g2, count, float(b1), float(b2), float(1.0 - b1), float(1.0 - b2),
                                                    inv_bc1, inv_bc2, work);
  adam_uniform_apply_kernel<<<grid, 256, 0, st>>>(p, g1, count, float(lr), inv_bc1, float(grad_limit), work,
                                                  reinterpret_cast<unsigned int *>(work + 2));
  return cudaGetLastError();
}

}  // namespace tsb
