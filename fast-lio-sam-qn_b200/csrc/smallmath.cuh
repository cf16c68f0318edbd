// smallmath.cuh -- tiny fp64 device helpers shared by gicp.cu and quatro.cu.
#pragma once
#include <cuda_runtime.h>

namespace b200 {

// eigenvector of the smallest eigenvalue of a symmetric 3x3 (cyclic Jacobi, fp64)
__device__ __forceinline__ void sym3_smallest_evec(double a00, double a01, double a02, double a11, double a12,
                                                   double a22, double n[3]) {
  double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-34 * dg || off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; pq++) {
      const int p = pq == 2 ? 1 : 0;
      const int q = pq == 0 ? 1 : 2;
      const int r = 3 - p - q;
      double apq = A[p][q];
      if (apq == 0.0) continue;
      double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      A[p][p] -= t * apq;
      A[q][q] += t * apq;
      A[p][q] = A[q][p] = 0.0;
      double arp = A[r][p], arq = A[r][q];
      A[r][p] = A[p][r] = c * arp - s * arq;
      A[r][q] = A[q][r] = s * arp + c * arq;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double vp = V[k][p], vq = V[k][q];
        V[k][p] = c * vp - s * vq;
        V[k][q] = s * vp + c * vq;
      }
    }
  }
  int m = 0;
  double best = A[0][0];
  if (A[1][1] < best) { best = A[1][1]; m = 1; }
  if (A[2][2] < best) { m = 2; }
  n[0] = m == 0 ? V[0][0] : (m == 1 ? V[0][1] : V[0][2]);
  n[1] = m == 0 ? V[1][0] : (m == 1 ? V[1][1] : V[1][2]);
  n[2] = m == 0 ? V[2][0] : (m == 1 ? V[2][1] : V[2][2]);
}

}  // namespace b200
