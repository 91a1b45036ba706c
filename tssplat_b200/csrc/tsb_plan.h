// Host-side plan for the fused energy+gradient kernel.
//
// Replaces what the reference gets from libpgo at construction
// (tssplat_ext/tet_spheres/tet_spheres.cpp:140-203: pgo_create_tet_gradient_matrix and
// pgo_create_tet_biharmonic_gradient_matrix, uploaded as two COO matrices): instead of sparse
// matrices we keep, per tet, the rest-shape inverse and the ids of the 8 vertices its smoothness
// stencil touches, grouped into tiles that one CTA processes out of shared memory.  Each tile's
// data is laid out as contiguous, 16-byte aligned blobs so the kernel can stage it with TMA bulk
// copies (cp.async.bulk) and never chases a pointer through global memory.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tsb {

// Header at the start of each tile's vertex blob (64 bytes).
struct TileHeader {
  int32_t ntet;      // tets in this tile (<= fill)
  int32_t nvert;     // vertices staged in shared memory (<= max_local_vertices)
  int32_t nrow;      // gather-table rows (a vertex with more than kRowCap entries spans several rows)
  int32_t ell_off;   // first entry in ell (multiple of 8 -> 16-byte aligned)
  int32_t nell;      // gather-table entries, padded to a multiple of 8
  int32_t pad[11];
};
static_assert(sizeof(TileHeader) == 64, "TileHeader must be 64 bytes");

// Gather-table rows hold at most kRowCap entries, so one hub vertex cannot serialise a warp; a
// vertex with in-tile degree d owns ceil(d / kRowCap) rows, each with its own scratch slot.
constexpr int kRowCap = 16;
inline int rows_cap(int tt, int nv) { return nv + 8 * tt / kRowCap; }           // rows per tile upper bound
inline int ell_cap(int tt, int nv) { return 8 * tt + 32 * kRowCap + rows_cap(tt, nv) + 64; }  // entries upper bound
// Vertex blob of one tile (NV = max_local_vertices, NR = rows_cap):
//   TileHeader | vlist int[NV] | X float[NV] | YZ float2[NV] | slot int[NR] | grp_ptr int[NR/32 + 4]
// `slot` (row order) is where the row's partial gradient goes in the float4 scratch array; the
// slots of one vertex are contiguous (ascending tile id, then row) so the combine kernel sums
// them in a fixed order.
inline int64_t vblob_bytes(int tt, int nv) { return 64 + int64_t(16) * nv + 4 * rows_cap(tt, nv) + 4 * (rows_cap(tt, nv) / 32 + 4); }
// Tet blob of one tile: idx8 (8 x u16)[TT] | B float[9*TT] (tet-major, 9 floats per tet)
inline int64_t tblob_bytes(int tt) { return int64_t(52) * tt; }

struct HostPlan {
  int32_t n = 0, nele = 0, tile_tets = 0, fill = 0, max_local_vertices = 0, n_tiles = 0, n_components = 0;
  int32_t laplacian_scale = 0, n_boundary_faces = 0, n_shared_vertices = 0, n_slots = 0;
  int64_t n_local_vertices = 0;

  std::vector<uint8_t> vblob;   // n_tiles * vblob_bytes(NV)
  std::vector<uint8_t> tblob;   // n_tiles * tblob_bytes(TT)
  // gather table (degree-sorted vertex order inside a tile): word offsets into the kernel's
  // [24][TT+4] output table; padding entries point at the table's zero column (offset TT);
  // layout [group][k/2][lane][2] so one 32-bit load fetches two entries of a lane
  std::vector<uint16_t> ell;
  std::vector<int32_t> slot_ptr;    // [n+1] scratch slots of each vertex (combine kernel CSR)
  std::vector<int32_t> tile_ell;    // [2*n_tiles] (ell_off, nell) per tile: lets the TMA producer size the copy
  std::vector<int32_t> tet_order;   // tile-order position -> original tet id
  std::vector<int32_t> tile_first;  // [n_tiles+1] position in tet_order of each tile's first tet
};

struct PlanOptions {
  int32_t tile_tets = 512;          // capacity TT of the compiled kernel variant
  int32_t max_local_vertices = 384; // capacity NV of the compiled kernel variant
  int32_t laplacian_scale = 0;
  int32_t balance_sms = 148;        // >0: pick the tile fill so the tile count is a multiple of this
};

// Returns 0 on success, TSB_E_* otherwise (message in err).
int build_plan(const float *rest_xyz, const int32_t *tets, int32_t n, int32_t nele,
               const PlanOptions &opt, HostPlan &plan, std::string &err);

}  // namespace tsb
