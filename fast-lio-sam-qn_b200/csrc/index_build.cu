// index_build.cu -- batched spatial-index build (replaces the serial nanoflann kd-tree build,
// third_party/nano_gicp/include/nano_gicp/impl/nanoflann_impl.hpp:1199-1211, 867-1012, which the
// reference runs twice per pair plus PCL's hidden third FLANN build).
//
// B200-first design: instead of a serial top-down pointer tree, every cloud becomes a
// Morton-sorted point array with a linear BVH on top (Karras 2012 radix tree, built with one
// thread per node, no recursion).  All clouds of a batch are built together: blockIdx.y is the
// cloud.  Steps: bbox (atomic min/max) -> 30-bit Morton keys -> stable LSD radix sort (4 x 8 bit;
// stable => the layout, and with it every later reduction order, is deterministic) -> gather into
// float4 (w carries the original index) -> radix-tree topology -> bottom-up AABBs with arrival
// flags, writing the two-children-per-node records the traversal reads (internal.cuh).
// A mid-count split over the same Morton order was measured first and discarded: ranges that
// straddle octant boundaries give huge overlapping boxes (195 node + 71 leaf visits per 15-NN
// query vs 29 + 7 with prefix splits on the 100k KITTI-shaped scan).
#include "internal.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------
__global__ void k_bbox_init(const CloudDev* clouds) {
  const CloudDev& c = clouds[blockIdx.x];
  if (threadIdx.x < 3) ((int*)c.bbox)[threadIdx.x] = f2ord(INFINITY);
  else if (threadIdx.x < 6) ((int*)c.bbox)[threadIdx.x] = f2ord(-INFINITY);
}

__global__ void __launch_bounds__(256) k_bbox(const CloudDev* clouds) {
  const CloudDev& c = clouds[blockIdx.y];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    const float* p = c.raw + (size_t)i * c.raw_stride;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      float v = p[d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  if ((threadIdx.x & 31) == 0 && blockIdx.x * blockDim.x < c.n) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      atomicMin(&((int*)c.bbox)[d], f2ord(mn[d]));
      atomicMax(&((int*)c.bbox)[3 + d], f2ord(mx[d]));
    }
  }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) k_morton(const CloudDev* clouds) {
  const CloudDev& c = clouds[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const int* bb = (const int*)c.bbox;
  float lo[3] = {ord2f(bb[0]), ord2f(bb[1]), ord2f(bb[2])};
  float ext = fmaxf(fmaxf(ord2f(bb[3]) - lo[0], ord2f(bb[4]) - lo[1]), ord2f(bb[5]) - lo[2]);
  float scale = ext > 0.f ? 1023.0f / ext : 0.f;  // cubic cells: one scale for all axes
  const float* p = c.raw + (size_t)i * c.raw_stride;
  uint32_t q[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    float f = (p[d] - lo[d]) * scale;
    int v = (int)f;
    q[d] = (uint32_t)min(max(v, 0), 1023);
  }
  c.keys[0][i] = (expand10(q[2]) << 2) | (expand10(q[1]) << 1) | expand10(q[0]);
  c.vals[0][i] = (uint32_t)i;
}

// ---- stable LSD radix sort, one 8-bit digit per pass --------------------------------
// pass p reads keys[p&1], writes keys[(p+1)&1].
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const CloudDev* clouds, int pass) {
  const CloudDev& c = clouds[blockIdx.y];
  const int base = blockIdx.x * SORT_TILE;
  if (base >= c.n) return;
  __shared__ uint32_t h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* keys = c.keys[pass & 1];
  const int shift = pass * RADIX_BITS;
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; j++) {
    int i = base + j * SORT_THREADS + threadIdx.x;
    if (i < c.n) atomicAdd(&h[(keys[i] >> shift) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  const int ntiles = (c.n + SORT_TILE - 1) / SORT_TILE;
  c.hist[threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// exclusive scan of the digit-major (digit, tile) table: one block per cloud
__global__ void __launch_bounds__(1024) k_sort_scan(const CloudDev* clouds) {
  const CloudDev& c = clouds[blockIdx.x];
  const int ntiles = (c.n + SORT_TILE - 1) / SORT_TILE;
  const int total = RADIX * ntiles;
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < total ? c.hist[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warp_sums[threadIdx.x];
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_sums[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
    if (i < total) c.hist[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const CloudDev* clouds, int pass) {
  const CloudDev& c = clouds[blockIdx.y];
  const int base = blockIdx.x * SORT_TILE;
  if (base >= c.n) return;
  constexpr int NW = SORT_THREADS / 32;
  __shared__ uint32_t cnt[NW][RADIX];
  for (int j = threadIdx.x; j < NW * RADIX; j += SORT_THREADS) (&cnt[0][0])[j] = 0;
  __syncthreads();
  const uint32_t* keys = c.keys[pass & 1];
  const uint32_t* vals = c.vals[pass & 1];
  uint32_t* okeys = c.keys[(pass + 1) & 1];
  uint32_t* ovals = c.vals[(pass + 1) & 1];
  const int shift = pass * RADIX_BITS;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  // warp w owns the contiguous run [base + w*32*ITEMS, +32*ITEMS): round r, lane l -> key r*32+l,
  // so (warp, round, lane) order == key order and the ranks below are stable.
  uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rk[SORT_ITEMS];
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; r++) {
    int i = base + (w * SORT_ITEMS + r) * 32 + lane;
    bool ok = i < c.n;
    key[r] = ok ? keys[i] : 0xFFFFFFFFu;
    val[r] = ok ? vals[i] : 0u;
    uint32_t dig = ok ? ((key[r] >> shift) & (RADIX - 1)) : RADIX;  // RADIX = "invalid"
    uint32_t m = __match_any_sync(0xffffffffu, dig);
    uint32_t old = ok ? cnt[w][dig] : 0u;
    __syncwarp();
    if (ok && (m & lt) == 0) cnt[w][dig] = old + __popc(m);
    __syncwarp();
    rk[r] = old + __popc(m & lt);
  }
  __syncthreads();
  {  // per digit: exclusive prefix over warps + the global (digit, tile) offset
    const int ntiles = (c.n + SORT_TILE - 1) / SORT_TILE;
    uint32_t run = c.hist[threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < NW; ww++) {
      uint32_t t = cnt[ww][threadIdx.x];
      cnt[ww][threadIdx.x] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; r++) {
    int i = base + (w * SORT_ITEMS + r) * 32 + lane;
    if (i < c.n) {
      uint32_t dig = (key[r] >> shift) & (RADIX - 1);
      uint32_t dst = cnt[w][dig] + rk[r];
      okeys[dst] = key[r];
      ovals[dst] = val[r];
    }
  }
}

// sorted float4 array (w = original index) and the inverse permutation
__global__ void __launch_bounds__(256) k_gather(const CloudDev* clouds, int final_buf) {
  const CloudDev& c = clouds[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  uint32_t o = c.vals[final_buf][i];
  const float* p = c.raw + (size_t)o * c.raw_stride;
  c.pts[i] = make_float4(p[0], p[1], p[2], __int_as_float((int)o));
  c.rank[o] = i;
}

// ---- Karras radix tree over the sorted keys -----------------------------------------
__device__ __forceinline__ int delta(const uint32_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint32_t x = keys[i] ^ keys[j];
  return x ? __clz(x) : 32 + __clz((uint32_t)i ^ (uint32_t)j);  // equal keys: fall back to the index
}

__global__ void __launch_bounds__(256) k_lbvh_topology(const CloudDev* clouds, int kbuf) {
  const CloudDev& c = clouds[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = c.n;
  if (i >= n - 1) return;
  const uint32_t* __restrict__ keys = c.keys[kbuf];
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int first = min(i, j), last = max(i, j);
  const int left_leaf = first == gamma, right_leaf = last == gamma + 1;
  c.info[i] = make_int4(first, last, left_leaf | (right_leaf << 1), gamma);
  if (left_leaf) c.parent_leaf[gamma] = i; else c.parent_node[gamma] = i;
  if (right_leaf) c.parent_leaf[gamma + 1] = i; else c.parent_node[gamma + 1] = i;
}

// Bottom-up AABBs.  Work items: every point whose parent spans > LEAF points (a 1-point leaf) and every internal
// node that is the root of a collapsed leaf (<= LEAF points, parent > LEAF): it reduces its <= 8 points directly.
// Each item then climbs; the second arrival at a node owns it, merges the children's boxes and writes the
// two-children traversal record.  (Starting a climb from every single point, through the tiny sub-trees, cost 2 atomics
// per node of the full radix tree and made this the slowest build kernel.)
__global__ void __launch_bounds__(256) k_lbvh_aabb(const CloudDev* clouds) {
  const CloudDev& c = clouds[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = c.n;
  if (n <= LEAF || t >= 2 * n - 1) return;
  int node;
  if (t < n) {  // point t
    node = c.parent_leaf[t];
    const int4 pi = c.info[node];
    if (pi.y - pi.x + 1 <= LEAF) return;  // lives inside a collapsed leaf
  } else {      // internal node t - n
    const int i = t - n;
    const int4 inf = c.info[i];
    if (i == 0 || inf.y - inf.x + 1 > LEAF) return;
    const int par = c.parent_node[i];
    const int4 pi = c.info[par];
    if (pi.y - pi.x + 1 <= LEAF) return;  // an ancestor is the collapsed-leaf root
    float lo0 = INFINITY, lo1 = INFINITY, lo2 = INFINITY, hi0 = -INFINITY, hi1 = -INFINITY, hi2 = -INFINITY;
    for (int p = inf.x; p <= inf.y; p++) {
      const float4 q = c.pts[p];
      lo0 = fminf(lo0, q.x); hi0 = fmaxf(hi0, q.x);
      lo1 = fminf(lo1, q.y); hi1 = fmaxf(hi1, q.y);
      lo2 = fminf(lo2, q.z); hi2 = fmaxf(hi2, q.z);
    }
    c.nbox[2 * i] = make_float4(lo0, lo1, lo2, 0.f);
    c.nbox[2 * i + 1] = make_float4(hi0, hi1, hi2, 0.f);
    __threadfence();
    node = par;
  }
  for (;;) {
    if (atomicAdd(&c.flags[node], 1u) == 0u) return;
    __threadfence();
    const int4 inf = c.info[node];
    const int g = inf.w;
    float4 lo[2], hi[2];
    int ref[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int ch = g + k;
      if ((inf.z >> k) & 1) {
        float4 q = c.pts[ch];
        lo[k] = q;
        hi[k] = q;
        ref[k] = leaf_ref(ch, 1);
      } else {
        lo[k] = __ldcg(&c.nbox[2 * ch]);
        hi[k] = __ldcg(&c.nbox[2 * ch + 1]);
        const int4 ci = c.info[ch];
        const int cnt = ci.y - ci.x + 1;
        ref[k] = cnt <= LEAF ? leaf_ref(ci.x, cnt) : ch;
      }
    }
    c.nbox[2 * node] = make_float4(fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.f);
    c.nbox[2 * node + 1] = make_float4(fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.f);
    c.tnodes[4 * node + 0] = make_float4(lo[0].x, lo[0].y, lo[0].z, __int_as_float(ref[0]));
    c.tnodes[4 * node + 1] = make_float4(hi[0].x, hi[0].y, hi[0].z, 0.f);
    c.tnodes[4 * node + 2] = make_float4(lo[1].x, lo[1].y, lo[1].z, __int_as_float(ref[1]));
    c.tnodes[4 * node + 3] = make_float4(hi[1].x, hi[1].y, hi[1].z, 0.f);
    if (node == 0) return;
    __threadfence();
    node = c.parent_node[node];
  }
}

// stable LSD radix sort of (keys[0], vals[0]) of every descriptor; result in keys[npass & 1] / vals[npass & 1].
// Only the n / keys / vals / hist fields of the descriptors are used (also by the voxel grid, assemble.cu).
int launch_radix_sort(const CloudDev* d_clouds, int count, int max_n, int npass, cudaStream_t s) {
  const int ntiles = (max_n + SORT_TILE - 1) / SORT_TILE;
  for (int p = 0; p < npass; p++) {
    k_sort_hist<<<dim3(ntiles, count), SORT_THREADS, 0, s>>>(d_clouds, p);
    k_sort_scan<<<count, 1024, 0, s>>>(d_clouds);
    k_sort_scatter<<<dim3(ntiles, count), SORT_THREADS, 0, s>>>(d_clouds, p);
  }
  return 3 * npass;
}

// ------------------------------------------------------------------------------------
// host launcher: builds `count` clouds whose descriptors are already in device memory.
// Returns the number of kernel launches issued.
int launch_index_build(const CloudDev* d_clouds, int count, int max_n, cudaStream_t s) {
  int launches = 0;
  k_bbox_init<<<count, 32, 0, s>>>(d_clouds); launches++;
  {
    int gx = min((max_n + 255) / 256, 296);
    k_bbox<<<dim3(gx, count), 256, 0, s>>>(d_clouds); launches++;
  }
  k_morton<<<dim3((max_n + 255) / 256, count), 256, 0, s>>>(d_clouds); launches++;
  const int npass = 4;  // 30-bit keys
  launches += launch_radix_sort(d_clouds, count, max_n, npass, s);
  k_gather<<<dim3((max_n + 255) / 256, count), 256, 0, s>>>(d_clouds, npass & 1); launches++;
  if (max_n > 1) {
    k_lbvh_topology<<<dim3((max_n + 254) / 256, count), 256, 0, s>>>(d_clouds, npass & 1); launches++;
    k_lbvh_aabb<<<dim3((2 * max_n + 254) / 256, count), 256, 0, s>>>(d_clouds); launches++;
  }
  return launches;
}

}  // namespace b200
