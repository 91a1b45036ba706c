// Host-side plan for the fused energy+gradient kernel.
//
// Replaces what the reference gets from libpgo at construction
// (tssplat_ext/tet_spheres/tet_spheres.cpp:140-203: pgo_create_tet_gradient_matrix and
// pgo_create_tet_biharmonic_gradient_matrix, uploaded as two COO matrices): instead of sparse
// matrices we keep, per tet, the rest-shape inverse and the ids of the 8 vertices its smoothness
// stencil touches, grouped into tiles that one CTA processes out of shared memory.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tsb {

// One CTA's work unit.  All offsets index the pooled arrays below.
struct TileDesc {
  int32_t ntet;      // tets in this tile (<= tile_tets)
  int32_t nvert;     // vertices staged in shared memory (<= max_local_vertices)
  int32_t vert_off;  // first entry in vlist / Xloc / dest
  int32_t ngrp;      // 32-wide vertex groups of the gather table (= ceil(nvert/32))
  int32_t grp_off;   // first entry in ell_grp_ptr (ngrp+1 entries, relative to ell_off)
  int32_t ell_off;   // first entry in ell
  int32_t cg_off;    // first entry in cg_list
  int32_t ncg;       // owner groups this tile contributes shared-vertex partials to
};

struct HostPlan {
  int32_t n = 0, nele = 0, tile_tets = 0, max_local_vertices = 0, n_tiles = 0, n_components = 0;
  int32_t laplacian_scale = 0, n_boundary_faces = 0, n_shared_vertices = 0, n_slots = 0;
  int64_t n_local_vertices = 0;

  std::vector<TileDesc> tiles;
  // per tet, tile-strided (tile t owns [t*tile_tets, (t+1)*tile_tets)):
  std::vector<uint16_t> idx8;   // 8 local vertex ids: own 0..3, opposite-of-face 0..3 (0xFFFF = boundary)
  std::vector<float> Bsoa;      // [tile][9][tile_tets]  rest inverse Dm^-1, row-major entries
  // per staged vertex, id-sorted inside a tile:
  std::vector<int32_t> vlist;   // global vertex id
  std::vector<float> Xloc;      // rest position (3 floats)
  // gather table (degree-sorted vertex order inside a tile):
  std::vector<int32_t> dest;    // >=0: global vertex id (tile is the only toucher); <0: -1-slot in scratch
  std::vector<uint16_t> ell;    // entries (tet_local*8 + slot), 0xFFFF = padding; [group][k][lane]
  std::vector<int32_t> ell_grp_ptr;
  // shared-vertex combine (last-arriver per owner tile):
  std::vector<int32_t> cg_list;
  std::vector<int32_t> need;        // [n_tiles] contributors per owner group (0 = no group)
  std::vector<int32_t> gsv_ptr;     // [n_tiles+1] shared vertices owned by each tile
  std::vector<int32_t> sv_vid;      // [n_shared] global vertex id
  std::vector<int32_t> sv_slot_ptr; // [n_shared+1] scratch slots (one per contributing tile, ascending tile id)
  std::vector<int32_t> tet_order;   // tile-order position -> original tet id
};

struct PlanOptions {
  int32_t tile_tets = 512;
  int32_t max_local_vertices = 512;
  int32_t laplacian_scale = 0;
};

// Returns 0 on success, TSB_E_* otherwise (message in err).
int build_plan(const float *rest_xyz, const int32_t *tets, int32_t n, int32_t nele,
               const PlanOptions &opt, HostPlan &plan, std::string &err);

}  // namespace tsb
