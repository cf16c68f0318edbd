// internal.cuh -- device-side data layout shared by the kernels of libb200reg.so.
//
// HBM layout of one cloud (all arrays in MORTON-SORTED order, position p):
//   pts   float4[P_pad]   (x, y, z, __int_as_float(original index)); padded to a whole number
//                         of leaves with +inf points (d2 = inf never enters a result set)
//   boxes float4[2*2*NLp] implicit complete binary tree over the leaves, heap ids 1..2*NLp-1;
//                         id -> (lo.xyz, hi.xyz) at boxes[2*id], boxes[2*id+1]; leaves are ids
//                         NLp..2*NLp-1, leaf l covers pts[l*LEAF, (l+1)*LEAF)
//   cov   double[6*P]     regularised covariance, symmetric 3x3 (xx,xy,xz,yy,yz,zz), 48 B/point
//   rank  int[P]          original index -> sorted position
// See DESIGN.md "Data layout in HBM".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int LEAF = 8;            // points per leaf: 8 x 16 B = one 128-B line
constexpr int SORT_THREADS = 256;  // radix sort tile = SORT_THREADS * SORT_ITEMS keys
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int STEP_THREADS = 128;  // threads per block of the per-point kernels
constexpr int NRED = 28;           // 21 (H upper) + 6 (b) + 1 (err)
constexpr int MAX_STACK = 24;      // >= tree depth + 1 (2^23 leaves * 8 pts = 67M points)

struct CloudDev {
  const float* raw;    // device copy of the caller's records (xyz at stride)
  int raw_stride;      // in floats
  int n;               // points
  int nl;              // leaves = ceil(n / LEAF)
  int nlp;             // leaves padded to a power of two
  int depth;           // log2(nlp)
  float4* pts;         // [nl * LEAF]
  float4* boxes;       // [2 * 2 * nlp]
  double* cov;         // [6 * n] (valid once has_cov)
  int* rank;           // [n]
  uint32_t* keys[2];   // sort ping-pong
  uint32_t* vals[2];
  uint32_t* hist;      // [RADIX * ntiles]
  uint32_t* flags;     // [nlp] tree build arrival flags
  float* bbox;         // [6] ordered-int encoded min/max
};

// phase of the per-pair LM state machine
enum Phase : int { PH_LINEARIZE = 0, PH_TRIAL = 1, PH_FITNESS = 2, PH_DONE = 3 };

struct GicpParamsDev {
  int max_iterations;
  int lm_max_iterations;
  double max_corr_dist2;
  double transformation_eps;
  double rotation_eps;
  double lm_init_lambda_factor;
  double icp_score_thr;
};

struct PairState {
  // current estimate x0 and trial xi = delta * x0 (row-major R, t)
  double R[9], t[3];
  double Rt[9], tt[3];
  double dR[9], dt[3];  // delta of the current trial
  double d[6];
  double H[36], b[6];
  double y0;
  double lambda, nu;
  double fitness;
  float Tf[12];  // x0.cast<float>() rows (r0 r1 r2 t)
  int phase;
  int outer_it;  // index of the current outer iteration
  int inner_it;  // LM trials done in this outer iteration
  int converged;
  int lm_failed;
  int n_lin, n_err;
  int nr_iterations;
  unsigned int arrive;  // block arrival counter for the last-block reduction
  int pad;
};

struct PairDev {
  CloudDev src, tgt;
  int* corr;       // [src.n] sorted target position or -1 (per sorted source position)
  float* sqd;      // [src.n]
  double* mahal;   // [6 * src.n]
  double* partial; // [nblocks * NRED]
};

// ---- ordered-int encoding of floats for atomicMin/atomicMax ------------------------
__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

}  // namespace b200
