// Synthetic KITTI-shaped scan generator.
//
// Not part of the registration hot path: this only manufactures inputs for the
// tests and bench.py (there is no dataset on the GPU box).  A virtual HDL-64E
// (64 beams, elevation +2.0 .. -24.8 deg linear, `scan_line: 64` in the
// reference's third_party/fastlio_config_launch/kitti.yaml:10) sits 1.73 m above
// the ground and ray-casts into a procedural street scene: ground plane, two rows
// of box buildings, 40 vertical cylinders (poles / trunks), 12 car-sized boxes.
// Range gate [2, 80] m (kitti.yaml:13 `blind: 2`), Gaussian range noise
// sigma = 0.02 m, then a uniform-random subsample to EXACTLY n points.
//
// Build: g++ -O2 -fopenmp -shared -fPIC synth.cpp -o libb200synth.so
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

namespace {

struct Box {  // oriented (yaw only) box, bottom at z0
  double cx, cy, z0, hx, hy, h, c, s;
};
struct Cyl {
  double cx, cy, r, h;
};
struct Scene {
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
};

double urand(std::mt19937_64& g, double a, double b) {
  // 53-bit uniform, written out so that it does not depend on libstdc++'s
  // uniform_real_distribution implementation.
  return a + (b - a) * ((g() >> 11) * (1.0 / 9007199254740992.0));
}

double nrand(std::mt19937_64& g) {  // Box-Muller, one sample per call
  double u1 = urand(g, 0.0, 1.0), u2 = urand(g, 0.0, 1.0);
  if (u1 < 1e-300) u1 = 1e-300;
  return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
}

Scene make_scene(uint64_t seed) {
  std::mt19937_64 g(seed * 0x9E3779B97F4A7C15ull + 12345);
  Scene sc;
  // two rows of buildings along the street (x axis), 8-15 m from the centre line
  for (int side = -1; side <= 1; side += 2) {
    double x = -130.0;
    while (x < 130.0) {
      double len = urand(g, 8.0, 22.0);
      double gap = urand(g, 0.5, 6.0);
      double depth = urand(g, 6.0, 12.0);
      double setback = urand(g, 8.0, 15.0);
      double h = urand(g, 6.0, 15.0);
      Box b;
      b.cx = x + 0.5 * len;
      b.cy = side * (setback + 0.5 * depth);
      b.z0 = 0.0;
      b.hx = 0.5 * len;
      b.hy = 0.5 * depth;
      b.h = h;
      double yaw = urand(g, -0.05, 0.05);
      b.c = std::cos(yaw);
      b.s = std::sin(yaw);
      sc.boxes.push_back(b);
      x += len + gap;
    }
  }
  // 12 car-sized boxes
  for (int i = 0; i < 12; i++) {
    Box b;
    b.cx = urand(g, -60.0, 60.0);
    double side = urand(g, 0.0, 1.0) < 0.5 ? -1.0 : 1.0;
    b.cy = side * urand(g, 3.0, 6.5);
    b.z0 = 0.0;
    b.hx = 0.5 * urand(g, 3.8, 4.8);
    b.hy = 0.5 * urand(g, 1.6, 1.9);
    b.h = urand(g, 1.4, 1.8);
    double yaw = urand(g, -0.15, 0.15);
    b.c = std::cos(yaw);
    b.s = std::sin(yaw);
    sc.boxes.push_back(b);
  }
  // 40 vertical cylinders
  for (int i = 0; i < 40; i++) {
    Cyl c;
    c.cx = urand(g, -100.0, 100.0);
    double side = urand(g, 0.0, 1.0) < 0.5 ? -1.0 : 1.0;
    c.cy = side * urand(g, 4.5, 7.8);
    c.r = urand(g, 0.15, 0.4);
    c.h = urand(g, 3.0, 9.0);
    sc.cyls.push_back(c);
  }
  return sc;
}

// first hit along o + t d, t in (0, inf); returns t or +inf
double cast(const Scene& sc, const double o[3], const double d[3]) {
  double best = INFINITY;
  if (d[2] < -1e-12) {
    double t = -o[2] / d[2];
    if (t > 0) best = t;
  }
  for (const Box& b : sc.boxes) {
    // ray into box frame
    double ox = o[0] - b.cx, oy = o[1] - b.cy;
    double lx = b.c * ox + b.s * oy, ly = -b.s * ox + b.c * oy, lz = o[2] - b.z0 - 0.5 * b.h;
    double dx = b.c * d[0] + b.s * d[1], dy = -b.s * d[0] + b.c * d[1], dz = d[2];
    double tmin = -INFINITY, tmax = INFINITY;
    const double lo[3] = {-b.hx, -b.hy, -0.5 * b.h}, hi[3] = {b.hx, b.hy, 0.5 * b.h};
    const double oo[3] = {lx, ly, lz}, dd[3] = {dx, dy, dz};
    bool miss = false;
    for (int a = 0; a < 3 && !miss; a++) {
      if (std::fabs(dd[a]) < 1e-12) {
        if (oo[a] < lo[a] || oo[a] > hi[a]) miss = true;
      } else {
        double t1 = (lo[a] - oo[a]) / dd[a], t2 = (hi[a] - oo[a]) / dd[a];
        if (t1 > t2) std::swap(t1, t2);
        tmin = std::max(tmin, t1);
        tmax = std::min(tmax, t2);
        if (tmin > tmax) miss = true;
      }
    }
    if (miss || tmax <= 0) continue;
    double t = tmin > 0 ? tmin : tmax;
    if (t > 0 && t < best) best = t;
  }
  for (const Cyl& c : sc.cyls) {
    double ox = o[0] - c.cx, oy = o[1] - c.cy;
    double A = d[0] * d[0] + d[1] * d[1];
    if (A < 1e-14) continue;
    double B = ox * d[0] + oy * d[1];
    double C = ox * ox + oy * oy - c.r * c.r;
    double disc = B * B - A * C;
    if (disc < 0) continue;
    double t = (-B - std::sqrt(disc)) / A;
    if (t <= 0) continue;
    double z = o[2] + t * d[2];
    if (z < 0 || z > c.h) continue;
    if (t < best) best = t;
  }
  return best;
}

}  // namespace

extern "C" {

// pose: row-major 4x4 double, sensor pose in the world frame (sensor -> world).
// out: n x 4 floats (x, y, z, intensity) in the SENSOR frame.
// Returns the number of points written (== n) or a negative error.
int b200synth_scan(uint64_t scene_seed, uint64_t scan_seed, const double* pose, int n, float* out) {
  if (n <= 0 || !pose || !out) return -1;
  Scene sc = make_scene(scene_seed);
  const double R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
  const double o[3] = {pose[3], pose[7], pose[11]};
  const int beams = 64;
  const double el_top = 2.0 * M_PI / 180.0, el_bot = -24.8 * M_PI / 180.0;
  std::vector<float> hits;  // xyz in the sensor frame, noise-free range stored separately
  std::vector<double> ranges;
  int az = std::max(16, (int)std::ceil(1.25 * n / beams));
  for (int attempt = 0; attempt < 12; attempt++) {
    hits.clear();
    ranges.clear();
    std::vector<double> tt((size_t)az * beams);
#pragma omp parallel for schedule(static)
    for (int a = 0; a < az; a++) {
      double phi = 2.0 * M_PI * (a + 0.5) / az;
      for (int b = 0; b < beams; b++) {
        double el = el_top + (el_bot - el_top) * b / (beams - 1);
        double ds[3] = {std::cos(el) * std::cos(phi), std::cos(el) * std::sin(phi), std::sin(el)};
        double dw[3] = {R[0] * ds[0] + R[1] * ds[1] + R[2] * ds[2], R[3] * ds[0] + R[4] * ds[1] + R[5] * ds[2],
                        R[6] * ds[0] + R[7] * ds[1] + R[8] * ds[2]};
        tt[(size_t)a * beams + b] = cast(sc, o, dw);
      }
    }
    size_t cnt = 0;
    for (double t : tt)
      if (t >= 2.0 && t <= 80.0) cnt++;
    if ((int)cnt >= n) {
      std::mt19937_64 g(scan_seed * 0xD1B54A32D192ED03ull + 777);
      // gather hits (serial: the noise stream must not depend on the thread count)
      std::vector<float> all;
      all.reserve(cnt * 3);
      for (int a = 0; a < az; a++) {
        double phi = 2.0 * M_PI * (a + 0.5) / az;
        for (int b = 0; b < beams; b++) {
          double t = tt[(size_t)a * beams + b];
          if (!(t >= 2.0 && t <= 80.0)) continue;
          double el = el_top + (el_bot - el_top) * b / (beams - 1);
          double r = t + 0.02 * nrand(g);
          all.push_back((float)(r * std::cos(el) * std::cos(phi)));
          all.push_back((float)(r * std::cos(el) * std::sin(phi)));
          all.push_back((float)(r * std::sin(el)));
        }
      }
      // uniform subsample to exactly n (partial Fisher-Yates)
      std::vector<uint32_t> perm(cnt);
      std::iota(perm.begin(), perm.end(), 0u);
      for (size_t i = 0; i < (size_t)n; i++) {
        size_t j = i + (size_t)(g() % (cnt - i));
        std::swap(perm[i], perm[j]);
      }
      for (int i = 0; i < n; i++) {
        uint32_t p = perm[i];
        out[4 * i + 0] = all[3 * p + 0];
        out[4 * i + 1] = all[3 * p + 1];
        out[4 * i + 2] = all[3 * p + 2];
        out[4 * i + 3] = (float)urand(g, 0.0, 1.0);
      }
      return n;
    }
    az = (int)std::ceil(az * 1.3) + 8;
  }
  return -2;
}

}  // extern "C"
