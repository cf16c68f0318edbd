// knn.cuh -- exact k-NN traversal of the implicit AABB tree (device functions).
//
// Replaces KdTreeFLANN::nearestKSearch -> findNeighbors -> searchLevel + KNNResultSet
// (third_party/nano_gicp/include/nano_gicp/nanoflann.hpp:140-152;
//  .../impl/nanoflann_impl.hpp:1229-1250, 1354-1418, 151-214).
//
// Exactness contract:
//   * point distance is the reference's fp32 expression ((dx*dx) + dy*dy) + dz*dz with
//     diff = query - point and NO fma contraction (nanoflann_impl.hpp:441-449), spelled with
//     __fsub_rn/__fmul_rn/__fadd_rn so nvcc cannot fuse it;
//   * the box lower bound uses the SAME operation sequence on the clamped per-axis gap, so by
//     monotonicity of IEEE rounding it never exceeds the fp32 distance of any point inside
//     the box; subtrees are skipped only when bound > current worst (strict), so candidates at
//     exactly the worst distance are still examined;
//   * ties on d2 resolve to the lower ORIGINAL index (the deterministic rule of App. A.3).
#pragma once
#include "internal.cuh"

namespace b200 {

__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float box_dist2_rn(float qx, float qy, float qz, float4 lo, float4 hi) {
  float dx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
  float dy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
  float dz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Result set of K entries kept ascending in registers (fully unrolled; no dynamic indexing).
template <int K>
struct KnnSet {
  float d[K];
  int p[K];  // sorted POSITION in the cloud (original index is pts[p].w)
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < K; j++) {
      d[j] = 3.402823466e+38f;
      p[j] = -1;
    }
  }
  __device__ __forceinline__ float worst() const { return d[K - 1]; }
};

__device__ __forceinline__ int orig_index(const float4* __restrict__ pts, int pos) {
  return pos < 0 ? 0x7FFFFFFF : __float_as_int(pts[pos].w);
}

// candidate (cd, cp with original index co) against slot (d, p): strictly better?
__device__ __forceinline__ bool cand_better(float cd, int co, float d, int p, const float4* __restrict__ pts) {
  if (cd < d) return true;
  if (cd > d) return false;
  return co < orig_index(pts, p);  // exact tie (rare): compare original indices
}

template <int K>
__device__ __forceinline__ void knn_insert(KnnSet<K>& s, float cd, int cp, int co, const float4* __restrict__ pts) {
  // single pass: carry the displaced entry down the ascending list
#pragma unroll
  for (int j = 0; j < K; j++) {
    if (cand_better(cd, co, s.d[j], s.p[j], pts)) {
      float td = s.d[j];
      int tp = s.p[j];
      s.d[j] = cd;
      s.p[j] = cp;
      cd = td;
      cp = tp;
      co = orig_index(pts, cp);
    }
  }
}

// Exact K-NN of (qx,qy,qz) in cloud c.  One thread per query; Morton-sorted queries keep
// neighbouring lanes on neighbouring paths (coherent loads, low divergence).
template <int K>
__device__ __forceinline__ void knn_search(const CloudDev& c, float qx, float qy, float qz, KnnSet<K>& res) {
  const float4* __restrict__ pts = c.pts;
  const float4* __restrict__ boxes = c.boxes;
  const int nlp = c.nlp;
  int stack_id[MAX_STACK];
  float stack_d[MAX_STACK];
  int sp = 0;
  int id = 1;
  float dnode = 0.f;
  for (;;) {
    // descend from `id` while it is an internal node worth visiting
    bool alive = !(dnode > res.worst());
    while (alive && id < nlp) {
      const int c0 = 2 * id;
      float4 lo0 = __ldg(&boxes[2 * c0]), hi0 = __ldg(&boxes[2 * c0 + 1]);
      float4 lo1 = __ldg(&boxes[2 * c0 + 2]), hi1 = __ldg(&boxes[2 * c0 + 3]);
      float d0 = box_dist2_rn(qx, qy, qz, lo0, hi0);
      float d1 = box_dist2_rn(qx, qy, qz, lo1, hi1);
      int nid = c0, fid = c0 + 1;
      float nd = d0, fd = d1;
      if (d1 < d0) {
        nid = c0 + 1; fid = c0; nd = d1; fd = d0;
      }
      const float w = res.worst();
      if (!(fd > w)) {
        stack_id[sp] = fid;
        stack_d[sp] = fd;
        sp++;
      }
      id = nid;
      dnode = nd;
      alive = !(nd > w);
    }
    if (alive) {  // leaf
      const int base = (id - nlp) * LEAF;
#pragma unroll
      for (int j = 0; j < LEAF; j++) {
        float4 p = __ldg(&pts[base + j]);
        float d2 = dist2_rn(qx, qy, qz, p.x, p.y, p.z);
        if (cand_better(d2, __float_as_int(p.w), res.d[K - 1], res.p[K - 1], pts))
          knn_insert<K>(res, d2, base + j, __float_as_int(p.w), pts);
      }
    }
    // pop
    bool found = false;
    while (sp > 0) {
      sp--;
      if (!(stack_d[sp] > res.worst())) {
        id = stack_id[sp];
        dnode = stack_d[sp];
        found = true;
        break;
      }
    }
    if (!found) break;
  }
}

}  // namespace b200
