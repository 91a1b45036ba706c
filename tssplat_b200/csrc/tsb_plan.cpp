// Host plan builder: rest inverses, face adjacency, locality-ordered tiles, per-tile vertex
// staging lists, gather tables and the shared-vertex combine lists.  See tsb_plan.h.
#include "tsb_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "../../include/tssplat_b200.h"

namespace tsb {
namespace {

struct FaceKey {
  uint32_t a, b, c, owner;  // sorted vertex triple; owner = 4*tet + local face
  bool operator<(const FaceKey &o) const {
    if (a != o.a) return a < o.a;
    if (b != o.b) return b < o.b;
    return c < o.c;
  }
  bool same(const FaceKey &o) const { return a == o.a && b == o.b && c == o.c; }
};

// Face k of a tet is the face opposite to local vertex k.
const int kFace[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};

inline uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000FF;
  v = (v | (v << 8)) & 0x0300F00F;
  v = (v | (v << 4)) & 0x030C30C3;
  v = (v | (v << 2)) & 0x09249249;
  return v;
}

struct UnionFind {
  std::vector<int32_t> p;
  explicit UnionFind(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
  int find(int x) {
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
  }
  void unite(int a, int b) {
    a = find(a); b = find(b);
    if (a != b) p[std::max(a, b)] = std::min(a, b);
  }
};

bool invert3(const double *m, double *o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double d = m[0] * c00 + m[1] * c01 + m[2] * c02;
  if (d == 0.0 || !std::isfinite(d)) return false;
  const double id = 1.0 / d;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}

}  // namespace

int build_plan(const float *rest, const int32_t *tets, int32_t n, int32_t nele, const PlanOptions &opt,
               HostPlan &P, std::string &err) {
  if (!rest || !tets || n <= 0 || nele <= 0) { err = "null input or non-positive size"; return TSB_E_INVALID; }
  const int TT = opt.tile_tets, NVMAX = opt.max_local_vertices;
  if (TT < 32 || TT > 2048 || (TT % 32) != 0 || NVMAX < 32 || NVMAX > 0x7FFF || (NVMAX % 32) != 0) {
    err = "tile_tets must be a multiple of 32 in [32,2048]";
    return TSB_E_INVALID;
  }
  P = HostPlan();
  P.n = n; P.nele = nele; P.tile_tets = TT; P.max_local_vertices = NVMAX; P.laplacian_scale = opt.laplacian_scale;

  // ---- validate, rest inverses (fp64 -> fp32 like the reference: tet_spheres.cpp:43-45) --------
  std::vector<float> Binv(size_t(nele) * 9);
  for (int t = 0; t < nele; ++t) {
    const int32_t *v = tets + 4 * size_t(t);
    for (int k = 0; k < 4; ++k)
      if (v[k] < 0 || v[k] >= n) { err = "tet " + std::to_string(t) + " has a vertex index out of range"; return TSB_E_MESH; }
    if (v[0] == v[1] || v[0] == v[2] || v[0] == v[3] || v[1] == v[2] || v[1] == v[3] || v[2] == v[3]) {
      err = "tet " + std::to_string(t) + " repeats a vertex"; return TSB_E_MESH;
    }
    double Dm[9], Bi[9];
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) Dm[3 * r + k] = double(rest[3 * size_t(v[k + 1]) + r]) - double(rest[3 * size_t(v[0]) + r]);
    if (!invert3(Dm, Bi)) { err = "tet " + std::to_string(t) + " has zero rest volume"; return TSB_E_MESH; }
    for (int i = 0; i < 9; ++i) Binv[size_t(t) * 9 + i] = float(Bi[i]);
  }

  // ---- face adjacency -> opposite vertex across each face ---------------------------------------
  std::vector<int32_t> opp(size_t(nele) * 4, -1);  // global id of the neighbour's vertex not on the shared face
  UnionFind uf(n);
  {
    std::vector<FaceKey> fk(size_t(nele) * 4);
    for (int t = 0; t < nele; ++t) {
      const int32_t *v = tets + 4 * size_t(t);
      uf.unite(v[0], v[1]); uf.unite(v[0], v[2]); uf.unite(v[0], v[3]);
      for (int k = 0; k < 4; ++k) {
        uint32_t f[3] = {uint32_t(v[kFace[k][0]]), uint32_t(v[kFace[k][1]]), uint32_t(v[kFace[k][2]])};
        if (f[0] > f[1]) std::swap(f[0], f[1]);
        if (f[1] > f[2]) std::swap(f[1], f[2]);
        if (f[0] > f[1]) std::swap(f[0], f[1]);
        fk[4 * size_t(t) + k] = FaceKey{f[0], f[1], f[2], uint32_t(4 * t + k)};
      }
    }
    std::sort(fk.begin(), fk.end());
    const size_t nf = fk.size();
    for (size_t i = 0; i < nf;) {
      size_t j = i + 1;
      while (j < nf && fk[j].same(fk[i])) ++j;
      if (j - i > 2) { err = "non-manifold mesh: a face is shared by more than two tets"; return TSB_E_MESH; }
      if (j - i == 2) {
        const uint32_t o0 = fk[i].owner, o1 = fk[i + 1].owner;
        opp[o0] = tets[o1];  // tets[4*t1 + k1] is the vertex of t1 opposite the shared face
        opp[o1] = tets[o0];
      } else {
        ++P.n_boundary_faces;
      }
      i = j;
    }
  }

  // ---- locality order: (component, Morton code of the rest centroid inside the component bbox) --
  std::vector<int32_t> comp_of_vertex(n);
  {
    std::vector<int32_t> label(n, -1);
    int32_t nc = 0;
    for (int v = 0; v < n; ++v) {
      const int r = uf.find(v);
      if (label[r] < 0) label[r] = nc++;
      comp_of_vertex[v] = label[r];
    }
    P.n_components = nc;
  }
  const int NC = P.n_components;
  std::vector<float> lo(size_t(NC) * 3, 3.0e38f), hi(size_t(NC) * 3, -3.0e38f);
  for (int v = 0; v < n; ++v) {
    const int c = comp_of_vertex[v];
    for (int r = 0; r < 3; ++r) {
      lo[3 * size_t(c) + r] = std::min(lo[3 * size_t(c) + r], rest[3 * size_t(v) + r]);
      hi[3 * size_t(c) + r] = std::max(hi[3 * size_t(c) + r], rest[3 * size_t(v) + r]);
    }
  }
  std::vector<uint64_t> key(nele);
  for (int t = 0; t < nele; ++t) {
    const int32_t *v = tets + 4 * size_t(t);
    const int c = comp_of_vertex[v[0]];
    uint32_t q[3];
    for (int r = 0; r < 3; ++r) {
      const float cen = 0.25f * (rest[3 * size_t(v[0]) + r] + rest[3 * size_t(v[1]) + r] + rest[3 * size_t(v[2]) + r] + rest[3 * size_t(v[3]) + r]);
      const float ext = hi[3 * size_t(c) + r] - lo[3 * size_t(c) + r];
      float u = ext > 0.f ? (cen - lo[3 * size_t(c) + r]) / ext : 0.f;
      u = std::min(std::max(u, 0.f), 1.f);
      q[r] = uint32_t(u * 1023.f);
    }
    const uint32_t m = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    key[t] = (uint64_t(uint32_t(c)) << 32) | m;
  }
  P.tet_order.resize(nele);
  std::iota(P.tet_order.begin(), P.tet_order.end(), 0);
  std::stable_sort(P.tet_order.begin(), P.tet_order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });

  // ---- tiling: recursive coordinate bisection inside each component ----------------------------
  // The tile count is matched to the SM count (a whole number of waves), tiles are distributed over
  // components in proportion to their tet counts, and each component is split into boxy parts of
  // (almost) equal size; inside a tile tets keep the Morton order (locality of the warp's gathers).
  // A tile that would stage more than NVMAX vertices makes its component use one more tile.
  {
    // component -> [begin, end) in tet_order (tet_order is sorted by component first)
    std::vector<int32_t> comp_begin(size_t(NC) + 1, 0);
    for (int t = 0; t < nele; ++t) comp_begin[size_t(key[t] >> 32) + 1]++;
    for (int c = 0; c < NC; ++c) comp_begin[size_t(c) + 1] += comp_begin[c];
    int64_t target = (int64_t(nele) + TT - 1) / TT;
    if (opt.balance_sms > 0) {
      const int64_t per_round = int64_t(opt.balance_sms) * TT;
      const int64_t rounds = (int64_t(nele) + per_round - 1) / per_round;
      // do not go below a quarter-full tile: tiny meshes keep a handful of tiles
      target = std::min<int64_t>(rounds * opt.balance_sms, std::max<int64_t>(target, (int64_t(nele) + TT / 4 - 1) / (TT / 4)));
    }
    std::vector<float> cen(size_t(nele) * 3);
    for (int t = 0; t < nele; ++t) {
      const int32_t *v = tets + 4 * size_t(t);
      for (int r = 0; r < 3; ++r)
        cen[3 * size_t(t) + r] = 0.25f * (rest[3 * size_t(v[0]) + r] + rest[3 * size_t(v[1]) + r] + rest[3 * size_t(v[2]) + r] + rest[3 * size_t(v[3]) + r]);
    }
    std::vector<int32_t> stamp(n, -1);
    int stamp_id = 0;
    auto staged_vertices = [&](const int32_t *first, const int32_t *last) {
      ++stamp_id;
      int cnt = 0;
      for (const int32_t *it = first; it != last; ++it)
        for (int k = 0; k < 4; ++k) {
          const int32_t a0 = tets[4 * size_t(*it) + k], a1 = opp[4 * size_t(*it) + k];
          if (stamp[a0] != stamp_id) { stamp[a0] = stamp_id; ++cnt; }
          if (a1 >= 0 && stamp[a1] != stamp_id) { stamp[a1] = stamp_id; ++cnt; }
        }
      return cnt;
    };
    std::vector<int32_t> work;          // tets of one component, permuted in place by the bisection
    std::vector<std::pair<int32_t, int32_t>> parts;   // [begin,end) in `work`
    struct Job { int32_t b, e, k; };
    std::vector<Job> stack;
    std::vector<int32_t> new_order;
    new_order.reserve(nele);
    P.tile_first.clear();
    int max_fill = 0;
    for (int c = 0; c < NC; ++c) {
      const int cb = comp_begin[c], ce = comp_begin[size_t(c) + 1], nt = ce - cb;
      int k = int(std::max<int64_t>((nt + TT - 1) / TT, (int64_t(nt) * target + nele / 2) / nele));
      k = std::max(1, std::min(k, nt));
      for (;;) {
        work.assign(P.tet_order.begin() + cb, P.tet_order.begin() + ce);
        parts.clear();
        stack.clear();
        stack.push_back(Job{0, nt, k});
        while (!stack.empty()) {
          const Job j = stack.back();
          stack.pop_back();
          if (j.k == 1) { parts.emplace_back(j.b, j.e); continue; }
          float lo3[3] = {3e38f, 3e38f, 3e38f}, hi3[3] = {-3e38f, -3e38f, -3e38f};
          for (int i = j.b; i < j.e; ++i)
            for (int r = 0; r < 3; ++r) {
              lo3[r] = std::min(lo3[r], cen[3 * size_t(work[i]) + r]);
              hi3[r] = std::max(hi3[r], cen[3 * size_t(work[i]) + r]);
            }
          int ax = 0;
          for (int r = 1; r < 3; ++r) if (hi3[r] - lo3[r] > hi3[ax] - lo3[ax]) ax = r;
          const int k1 = j.k / 2, k2 = j.k - k1;
          const int n1 = int((int64_t(j.e - j.b) * k1 + j.k / 2) / j.k);
          std::nth_element(work.begin() + j.b, work.begin() + j.b + n1, work.begin() + j.e, [&](int32_t a, int32_t b2) {
            const float ca = cen[3 * size_t(a) + ax], cb2 = cen[3 * size_t(b2) + ax];
            return ca < cb2 || (ca == cb2 && a < b2);
          });
          stack.push_back(Job{j.b + n1, j.e, k2});
          stack.push_back(Job{j.b, j.b + n1, k1});
        }
        bool ok = true;
        for (const auto &pr : parts) {
          if (pr.second - pr.first > TT) { ok = false; break; }
          if (staged_vertices(work.data() + pr.first, work.data() + pr.second) > NVMAX) { ok = false; break; }
        }
        if (ok || k >= nt) break;
        k = std::min(nt, k + std::max(1, k / 8));
      }
      std::sort(parts.begin(), parts.end());
      for (const auto &pr : parts) {
        if (pr.second == pr.first) continue;
        std::sort(work.begin() + pr.first, work.begin() + pr.second, [&](int32_t a, int32_t b2) { return key[a] < key[b2] || (key[a] == key[b2] && a < b2); });
        P.tile_first.push_back(int32_t(new_order.size()));
        new_order.insert(new_order.end(), work.begin() + pr.first, work.begin() + pr.second);
        max_fill = std::max(max_fill, pr.second - pr.first);
      }
    }
    P.tile_first.push_back(nele);
    P.tet_order.swap(new_order);
    P.fill = std::min(TT, ((max_fill + 7) / 8) * 8);
  }
  const int NTILE = int(P.tile_first.size()) - 1;
  P.n_tiles = NTILE;
  const int64_t VB = vblob_bytes(TT, NVMAX), TB = tblob_bytes(TT);
  const int NR = rows_cap(TT, NVMAX);
  P.vblob.assign(size_t(NTILE) * VB, 0);
  P.tblob.assign(size_t(NTILE) * TB, 0);

  // ---- per tile: staged vertex list (id-sorted), local stencil ids, rest inverses ---------------
  std::vector<int32_t> local_of(n, -1);
  std::vector<std::vector<int32_t>> tile_verts(NTILE);
  std::vector<uint16_t> idx8(size_t(NTILE) * TT * 8, 0);   // kept for the gather-table pass below
  for (int tile = 0; tile < NTILE; ++tile) {
    const int p0 = P.tile_first[tile], p1 = P.tile_first[tile + 1];
    std::vector<int32_t> &vs = tile_verts[tile];
    vs.reserve(size_t(p1 - p0));
    for (int pos = p0; pos < p1; ++pos) {
      const int t = P.tet_order[pos];
      for (int k = 0; k < 4; ++k) {
        vs.push_back(tets[4 * size_t(t) + k]);
        if (opp[4 * size_t(t) + k] >= 0) vs.push_back(opp[4 * size_t(t) + k]);
      }
    }
    std::sort(vs.begin(), vs.end());
    vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
    for (size_t i = 0; i < vs.size(); ++i) local_of[vs[i]] = int32_t(i);
    P.n_local_vertices += int64_t(vs.size());
    uint8_t *tb = P.tblob.data() + size_t(tile) * TB;
    uint16_t *ids = reinterpret_cast<uint16_t *>(tb);
    float *Bt = reinterpret_cast<float *>(tb + size_t(16) * TT);
    for (int pos = p0; pos < p1; ++pos) {
      const int t = P.tet_order[pos], lt = pos - p0;
      uint16_t *d = ids + size_t(lt) * 8, *d2 = &idx8[(size_t(tile) * TT + lt) * 8];
      for (int k = 0; k < 4; ++k) {
        d[k] = uint16_t(local_of[tets[4 * size_t(t) + k]]);
        const int32_t o = opp[4 * size_t(t) + k];
        // boundary face: point at the tet's own opposite vertex (harmless gather), bit 15 clear;
        // interior face: neighbour's opposite vertex, bit 15 set
        d[4 + k] = o >= 0 ? uint16_t(0x8000u | uint16_t(local_of[o])) : d[k];
        d2[k] = d[k];
        d2[4 + k] = o >= 0 ? uint16_t(local_of[o]) : uint16_t(0xFFFF);
      }
      for (int i = 0; i < 9; ++i) Bt[size_t(lt) * 9 + i] = Binv[size_t(t) * 9 + i];
    }
  }

  // ---- in-tile degree of every staged vertex -> gather-table rows -> scratch slots ----------------
  // rows_of[tile][i] = ceil(deg / kRowCap) for the i-th staged vertex (id-sorted)
  std::vector<std::vector<int32_t>> tile_deg(NTILE);
  for (int tile = 0; tile < NTILE; ++tile) {
    const int nv = int(tile_verts[tile].size()), ntet = P.tile_first[tile + 1] - P.tile_first[tile];
    std::vector<int32_t> &dg = tile_deg[tile];
    dg.assign(nv, 0);
    for (int lt = 0; lt < ntet; ++lt) {
      const uint16_t *d = &idx8[(size_t(tile) * TT + lt) * 8];
      for (int s = 0; s < 8; ++s) if (d[s] != 0xFFFF) dg[d[s]]++;
    }
  }
  // slot_ptr[v]..slot_ptr[v+1]: scratch slots of vertex v, ordered by (tile, row)
  P.slot_ptr.assign(size_t(n) + 1, 0);
  std::vector<int32_t> touches(n, 0);
  for (int tile = 0; tile < NTILE; ++tile) {
    const std::vector<int32_t> &vs = tile_verts[tile];
    for (size_t i = 0; i < vs.size(); ++i) {
      P.slot_ptr[size_t(vs[i]) + 1] += (tile_deg[tile][i] + kRowCap - 1) / kRowCap;
      touches[vs[i]]++;
    }
  }
  for (int v = 0; v < n; ++v) {
    P.slot_ptr[size_t(v) + 1] += P.slot_ptr[v];
    P.n_shared_vertices += touches[v] > 1;
  }
  if (int64_t(P.slot_ptr[n]) > int64_t(0x7fffffff) / 4) { err = "too many scratch slots"; return TSB_E_INVALID; }
  P.n_slots = P.slot_ptr[n];
  std::vector<int32_t> next_slot(P.slot_ptr.begin(), P.slot_ptr.end() - 1);   // tiles visited in ascending order

  // ---- per tile: vertex blob, gather table ---------------------------------------------------------
  const int TTP = TT + 4;   // output-table row stride; column TT is the zero column
  struct Row { int32_t vert, chunk, len; };
  std::vector<Row> rows;
  std::vector<int32_t> grp_rel, fill_cnt, row_n;
  std::vector<uint16_t> row_ent;
  for (int tile = 0; tile < NTILE; ++tile) {
    const std::vector<int32_t> &vs = tile_verts[tile];
    const std::vector<int32_t> &dg = tile_deg[tile];
    const int nv = int(vs.size()), ntet = P.tile_first[tile + 1] - P.tile_first[tile];
    uint8_t *vb = P.vblob.data() + size_t(tile) * VB;
    TileHeader *hd = reinterpret_cast<TileHeader *>(vb);
    int32_t *vlist = reinterpret_cast<int32_t *>(vb + 64);
    float *Xx = reinterpret_cast<float *>(vb + 64 + size_t(4) * NVMAX);
    float *YZ = reinterpret_cast<float *>(vb + 64 + size_t(8) * NVMAX);
    int32_t *slot = reinterpret_cast<int32_t *>(vb + 64 + size_t(16) * NVMAX);
    int32_t *grp_ptr = reinterpret_cast<int32_t *>(vb + 64 + size_t(16) * NVMAX + size_t(4) * NR);
    for (int i = 0; i < nv; ++i) {
      vlist[i] = vs[i];
      Xx[i] = rest[3 * size_t(vs[i])];
      YZ[2 * i] = rest[3 * size_t(vs[i]) + 1];
      YZ[2 * i + 1] = rest[3 * size_t(vs[i]) + 2];
    }
    // rows, longest first (stable: vertex id, then chunk)
    rows.clear();
    for (int i = 0; i < nv; ++i) {
      const int nr = (dg[i] + kRowCap - 1) / kRowCap;
      for (int c = 0; c < nr; ++c) rows.push_back(Row{i, c, std::min(kRowCap, dg[i] - c * kRowCap)});
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.len > b.len; });
    const int nrow = int(rows.size());
    if (nrow > NR) { err = "internal: gather-table rows exceed capacity"; return TSB_E_INVALID; }
    const int ngrp = (nrow + 31) / 32;
    grp_rel.assign(size_t(ngrp) + 1, 0);
    for (int g = 0; g < ngrp; ++g) grp_rel[g + 1] = grp_rel[g] + 32 * ((rows[size_t(g) * 32].len + 1) & ~1);
    const int nell = ((grp_rel[ngrp] + 7) / 8) * 8;
    if (nell > ell_cap(TT, NVMAX)) { err = "internal: gather table exceeds capacity"; return TSB_E_INVALID; }
    while (P.ell.size() % 8) P.ell.push_back(uint16_t(TT));
    if (P.ell.size() + size_t(nell) > size_t(0x7fffffff)) { err = "gather table exceeds 32-bit offsets"; return TSB_E_INVALID; }
    hd->ntet = ntet; hd->nvert = nv; hd->nrow = nrow; hd->ell_off = int32_t(P.ell.size()); hd->nell = nell;
    P.tile_ell.push_back(hd->ell_off); P.tile_ell.push_back(nell);
    for (int g = 0; g <= ngrp; ++g) grp_ptr[g] = grp_rel[g];
    P.ell.resize(size_t(hd->ell_off) + size_t(nell), uint16_t(TT));
    // row index of (vertex, chunk 0); chunks of a vertex are NOT adjacent after sorting, so map each
    std::vector<std::vector<int32_t>> row_of(nv);
    for (int r = 0; r < nrow; ++r) {
      std::vector<int32_t> &ro = row_of[rows[r].vert];
      if (int(ro.size()) <= rows[r].chunk) ro.resize(size_t(rows[r].chunk) + 1, -1);
      ro[rows[r].chunk] = r;
    }
    // collect each row's entries (word offset of component 0 in the [24][TTP] table)
    row_ent.assign(size_t(nrow) * kRowCap, 0);
    row_n.assign(nrow, 0);
    fill_cnt.assign(nv, 0);
    for (int lt = 0; lt < ntet; ++lt) {
      const uint16_t *d = &idx8[(size_t(tile) * TT + lt) * 8];
      for (int s = 0; s < 8; ++s) {
        if (d[s] == 0xFFFF) continue;
        const int cnt = fill_cnt[d[s]]++;
        const int r = row_of[d[s]][cnt / kRowCap];
        row_ent[size_t(r) * kRowCap + row_n[r]++] = uint16_t(s * 3 * TTP + lt);
      }
    }
    // Order the entries of the 32 rows of a group so that, column by column, the 32 lanes of the
    // warp read distinct shared-memory banks (bank = word offset mod 32; the three components
    // are +TTP words apart, i.e. the same permutation shifted).  Greedy column-by-column
    // matching, rows with the fewest remaining entries first within a column.
    for (int g = 0; g < ngrp; ++g) {
      const int r0 = g * 32, r1 = std::min(nrow, r0 + 32);
      const int glen = (grp_rel[g + 1] - grp_rel[g]) / 32;
      const size_t base = size_t(hd->ell_off) + size_t(grp_rel[g]);
      uint8_t used_e[32][kRowCap] = {};
      for (int k = 0; k < glen; ++k) {
        uint32_t bank_used = 0;
        // two passes: first place rows that can take a free bank, then the rest
        int pending[32], np = 0;
        for (int r = r0; r < r1; ++r) {
          if (k >= row_n[r]) continue;   // this row is already exhausted -> padding (zero column)
          int pick = -1;
          for (int e = 0; e < row_n[r]; ++e) {
            if (used_e[r - r0][e]) continue;
            const int bank = row_ent[size_t(r) * kRowCap + e] & 31;
            if (!(bank_used >> bank & 1u)) { pick = e; break; }
          }
          if (pick < 0) { pending[np++] = r; continue; }
          used_e[r - r0][pick] = 1;
          bank_used |= 1u << (row_ent[size_t(r) * kRowCap + pick] & 31);
          P.ell[base + size_t(k >> 1) * 64 + size_t(r - r0) * 2 + (k & 1)] = row_ent[size_t(r) * kRowCap + pick];
        }
        for (int q = 0; q < np; ++q) {
          const int r = pending[q];
          int pick = -1;
          for (int e = 0; e < row_n[r] && pick < 0; ++e) if (!used_e[r - r0][e]) pick = e;
          used_e[r - r0][pick] = 1;
          P.ell[base + size_t(k >> 1) * 64 + size_t(r - r0) * 2 + (k & 1)] = row_ent[size_t(r) * kRowCap + pick];
        }
      }
    }
    // scratch slot of each row: vertex's slots are ordered by (tile, chunk)
    for (int i = 0; i < nv; ++i) {
      const int nr = int(row_of[i].size());
      for (int c = 0; c < nr; ++c) slot[row_of[i][c]] = next_slot[vs[i]] + c;
      next_slot[vs[i]] += nr;
    }
  }
  return TSB_OK;
}

}  // namespace tsb
