// tssplat-b200: setup-time surface extraction on the GPU  ("next" row (f)3).
//
// Replaces the reference's get_surface_vf (geometry/mesh_utils.py:5-35, numpy on the host; re-run by
// reset() / permute_surface_v(), geometry/tetmesh_geometry.py:164-170,369-371): the faces that belong to exactly one
// tet, listed in lexicographic order of their sorted vertex triple, each in the orientation its tet gives it, and
// re-indexed into the sorted list of surface vertex ids.
//
// Pipeline (all on the device, one stream): 4T face records -> two stable radix sorts (by the largest vertex, then
// by the (smallest, middle) pair: lexicographic order of the sorted triple for any vertex count < 2^31) -> a record
// is a surface face iff neither neighbour in sorted order has the same triple -> stream compaction (order kept) ->
// vertex flags -> exclusive scan = new vertex ids -> relabel.  The sort / select / scan primitives are CUB's (setup
// code, not the per-iteration path).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include "../../include/tssplat_b200.h"

namespace {

thread_local std::string g_setup_err;

int fail(int code, const std::string &msg) {
  g_setup_err = msg;
  return code;
}

// local vertices of face k (opposite local vertex k), in the orientation the tet gives it
__device__ __constant__ int kCorner[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};

__device__ __forceinline__ void sorted_triple(const int32_t *__restrict__ tets, uint32_t rec, uint32_t &a, uint32_t &b, uint32_t &c) {
  const int32_t *v = tets + 4 * size_t(rec >> 2);
  const int k = int(rec & 3u);
  a = uint32_t(v[kCorner[k][0]]); b = uint32_t(v[kCorner[k][1]]); c = uint32_t(v[kCorner[k][2]]);
  if (a > b) { const uint32_t t = a; a = b; b = t; }
  if (b > c) { const uint32_t t = b; b = c; c = t; }
  if (a > b) { const uint32_t t = a; a = b; b = t; }
}

// record id = k * T + t would be the reference's tie order (block k = faces opposite vertex k); ties only matter for
// faces shared by two tets, which are dropped, so rec = 4 t + k is used (coalesced tet reads)
__global__ void face_key_c_kernel(const int32_t *__restrict__ tets, int64_t nrec, uint32_t *key_c, uint32_t *rec_out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= nrec) return;
  uint32_t a, b, c;
  sorted_triple(tets, uint32_t(i), a, b, c);
  key_c[i] = c;
  rec_out[i] = uint32_t(i);
}

__global__ void face_key_ab_kernel(const int32_t *__restrict__ tets, const uint32_t *__restrict__ rec, int64_t nrec, uint64_t *key_ab) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= nrec) return;
  uint32_t a, b, c;
  sorted_triple(tets, rec[i], a, b, c);
  key_ab[i] = (uint64_t(a) << 32) | uint64_t(b);
}

__global__ void once_flag_kernel(const int32_t *__restrict__ tets, const uint32_t *__restrict__ rec, int64_t nrec, uint8_t *flag) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= nrec) return;
  uint32_t a, b, c, pa, pb, pc;
  sorted_triple(tets, rec[i], a, b, c);
  bool once = true;
  if (i > 0) { sorted_triple(tets, rec[i - 1], pa, pb, pc); once = once && !(pa == a && pb == b && pc == c); }
  if (i + 1 < nrec) { sorted_triple(tets, rec[i + 1], pa, pb, pc); once = once && !(pa == a && pb == b && pc == c); }
  flag[i] = once ? 1 : 0;
}

__global__ void mark_vertices_kernel(const int32_t *__restrict__ tets, const uint32_t *__restrict__ srec, int32_t nsf, int32_t *used) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsf) return;
  const uint32_t r = srec[i];
  const int32_t *v = tets + 4 * size_t(r >> 2);
  const int k = int(r & 3u);
  for (int j = 0; j < 3; ++j) used[v[kCorner[k][j]]] = 1;      // same value from every writer
}

__global__ void relabel_kernel(const int32_t *__restrict__ tets, const uint32_t *__restrict__ srec, int32_t nsf,
                               const int32_t *__restrict__ newid, int32_t *faces) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsf) return;
  const uint32_t r = srec[i];
  const int32_t *v = tets + 4 * size_t(r >> 2);
  const int k = int(r & 3u);
  for (int j = 0; j < 3; ++j) faces[3 * size_t(i) + j] = newid[v[kCorner[k][j]]];
}

__global__ void surface_vid_kernel(const int32_t *__restrict__ used, const int32_t *__restrict__ newid, int32_t n, int32_t *vid) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n && used[v]) vid[newid[v]] = v;
}

struct Scratch {            // frees everything on scope exit
  void *p[12] = {};
  int cnt = 0;
  cudaError_t get(void **out, size_t bytes) {
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
    if (e == cudaSuccess) p[cnt++] = *out;
    return e;
  }
  ~Scratch() { for (int i = 0; i < cnt; ++i) cudaFree(p[i]); }
};

inline unsigned blocks(int64_t count, int bs) { return unsigned((count + bs - 1) / bs); }

}  // namespace

extern "C" {

const char *tsb_setup_last_error(void) { return g_setup_err.c_str(); }

void tsb_free_host(void *p) { std::free(p); }

int tsb_surface_extract(const int32_t *tets_host, int32_t nele, int32_t n, int device, int32_t *nsv_out, int32_t *nsf_out,
                        int32_t **surface_vid_out, int32_t **surface_f_out) {
  if (!tets_host || !nsv_out || !nsf_out || !surface_vid_out || !surface_f_out || nele < 0 || n < 0)
    return fail(TSB_E_INVALID, "tsb_surface_extract: null pointer or negative size");
  *nsv_out = *nsf_out = 0;
  *surface_vid_out = *surface_f_out = nullptr;
  if (int64_t(nele) * 4 > int64_t(0x7FFFFFFF)) return fail(TSB_E_INVALID, "tsb_surface_extract: more than 2^29 tets");
  for (int64_t i = 0; i < int64_t(nele) * 4; ++i)
    if (tets_host[i] < 0 || tets_host[i] >= n) return fail(TSB_E_MESH, "tet " + std::to_string(i / 4) + " has a vertex index out of range");
  int ndev = 0, prev = -1;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(TSB_E_CUDA, "no CUDA device " + std::to_string(device) + " (tssplat_b200 has no CPU path)");
  }
  cudaGetDevice(&prev);
  if (prev != device && cudaSetDevice(device) != cudaSuccess) return fail(TSB_E_CUDA, "cannot select the CUDA device");
  struct Restore { int prev, dev; ~Restore() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); } } restore{prev, device};
  if (nele == 0) return TSB_OK;

  const int64_t nrec = int64_t(nele) * 4;
  Scratch sc;
  cudaStream_t st = nullptr;      // the default stream: a setup call, synchronous for the caller
  int32_t *d_tets = nullptr, *d_used = nullptr, *d_newid = nullptr, *d_faces = nullptr, *d_vid = nullptr, *d_count = nullptr;
  uint32_t *d_keyc[2] = {nullptr, nullptr}, *d_rec[2] = {nullptr, nullptr}, *d_srec = nullptr;
  uint64_t *d_keyab[2] = {nullptr, nullptr};
  uint8_t *d_flag = nullptr;
  void *d_tmp = nullptr;
  cudaError_t e = sc.get(reinterpret_cast<void **>(&d_tets), size_t(nrec) * 4);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
    e = sc.get(reinterpret_cast<void **>(&d_keyc[k]), size_t(nrec) * 4);
    if (e == cudaSuccess) e = sc.get(reinterpret_cast<void **>(&d_rec[k]), size_t(nrec) * 4);
    if (e == cudaSuccess) e = sc.get(reinterpret_cast<void **>(&d_keyab[k]), size_t(nrec) * 8);
  }
  if (e == cudaSuccess) e = sc.get(reinterpret_cast<void **>(&d_flag), size_t(nrec));
  if (e == cudaSuccess) e = sc.get(reinterpret_cast<void **>(&d_count), 16);
  if (e != cudaSuccess) return fail(TSB_E_NOMEM, std::string("tsb_surface_extract: cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemcpyAsync(d_tets, tets_host, size_t(nrec) * 4, cudaMemcpyHostToDevice, st);

  // CUB temporary storage: the largest of the four primitives
  size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  cub::DoubleBuffer<uint32_t> kc(d_keyc[0], d_keyc[1]), rc1(d_rec[0], d_rec[1]);
  cub::DoubleBuffer<uint64_t> kab(d_keyab[0], d_keyab[1]);
  int bits_c = 1;
  while (bits_c < 32 && (int64_t(1) << bits_c) < int64_t(n)) ++bits_c;
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, t1, kc, rc1, int(nrec), 0, bits_c, st);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, t2, kab, rc1, int(nrec), 0, 32 + bits_c, st);
  if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(nullptr, t3, d_rec[0], d_flag, d_rec[1], d_count, int(nrec), st);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, t4, static_cast<int32_t *>(nullptr), static_cast<int32_t *>(nullptr), n + 1, st);
  size_t tmp_bytes = t1 > t2 ? t1 : t2;
  tmp_bytes = tmp_bytes > t3 ? tmp_bytes : t3;
  tmp_bytes = tmp_bytes > t4 ? tmp_bytes : t4;
  if (e == cudaSuccess) e = sc.get(&d_tmp, tmp_bytes);
  if (e != cudaSuccess) return fail(TSB_E_CUDA, std::string("tsb_surface_extract: setup: ") + cudaGetErrorString(e));

  // sort the 4T records by their sorted vertex triple (a, b, c): stable sort by c, then by (a, b)
  face_key_c_kernel<<<blocks(nrec, 256), 256, 0, st>>>(d_tets, nrec, kc.Current(), rc1.Current());
  e = cub::DeviceRadixSort::SortPairs(d_tmp, t1, kc, rc1, int(nrec), 0, bits_c, st);
  if (e == cudaSuccess) {
    face_key_ab_kernel<<<blocks(nrec, 256), 256, 0, st>>>(d_tets, rc1.Current(), nrec, kab.Current());
    e = cub::DeviceRadixSort::SortPairs(d_tmp, t2, kab, rc1, int(nrec), 0, 32 + bits_c, st);
  }
  // faces that occur once, in sorted order
  uint32_t *sorted_rec = rc1.Current();
  d_srec = rc1.Alternate();
  if (e == cudaSuccess) {
    once_flag_kernel<<<blocks(nrec, 256), 256, 0, st>>>(d_tets, sorted_rec, nrec, d_flag);
    e = cub::DeviceSelect::Flagged(d_tmp, t3, sorted_rec, d_flag, d_srec, d_count, int(nrec), st);
  }
  int32_t nsf = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&nsf, d_count, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(TSB_E_CUDA, std::string("tsb_surface_extract: sort: ") + cudaGetErrorString(e));

  // surface vertices: flags -> exclusive scan -> ids in increasing vertex order (the reference's np.unique)
  e = sc.get(reinterpret_cast<void **>(&d_used), size_t(n + 1) * 4);
  if (e == cudaSuccess) e = sc.get(reinterpret_cast<void **>(&d_newid), size_t(n + 1) * 4);
  if (e == cudaSuccess) e = sc.get(reinterpret_cast<void **>(&d_faces), size_t(nsf) * 12);
  if (e != cudaSuccess) return fail(TSB_E_NOMEM, std::string("tsb_surface_extract: cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemsetAsync(d_used, 0, size_t(n + 1) * 4, st);
  int32_t nsv = 0;
  if (e == cudaSuccess && nsf > 0) mark_vertices_kernel<<<blocks(nsf, 256), 256, 0, st>>>(d_tets, d_srec, nsf, d_used);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_tmp, t4, d_used, d_newid, n + 1, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&nsv, d_newid + n, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(TSB_E_CUDA, std::string("tsb_surface_extract: scan: ") + cudaGetErrorString(e));
  e = sc.get(reinterpret_cast<void **>(&d_vid), size_t(nsv) * 4);
  if (e != cudaSuccess) return fail(TSB_E_NOMEM, std::string("tsb_surface_extract: cudaMalloc: ") + cudaGetErrorString(e));
  if (nsf > 0) relabel_kernel<<<blocks(nsf, 256), 256, 0, st>>>(d_tets, d_srec, nsf, d_newid, d_faces);
  if (n > 0) surface_vid_kernel<<<blocks(n, 256), 256, 0, st>>>(d_used, d_newid, n, d_vid);
  e = cudaGetLastError();

  int32_t *h_vid = static_cast<int32_t *>(std::malloc(size_t(nsv ? nsv : 1) * 4));
  int32_t *h_f = static_cast<int32_t *>(std::malloc(size_t(nsf ? nsf : 1) * 12));
  if (!h_vid || !h_f) { std::free(h_vid); std::free(h_f); return fail(TSB_E_NOMEM, "tsb_surface_extract: host allocation failed"); }
  if (e == cudaSuccess && nsv) e = cudaMemcpyAsync(h_vid, d_vid, size_t(nsv) * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && nsf) e = cudaMemcpyAsync(h_f, d_faces, size_t(nsf) * 12, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    std::free(h_vid); std::free(h_f);
    return fail(TSB_E_CUDA, std::string("tsb_surface_extract: ") + cudaGetErrorString(e));
  }
  *nsv_out = nsv; *nsf_out = nsf;
  *surface_vid_out = h_vid; *surface_f_out = h_f;
  return TSB_OK;
}

}  // extern "C"
