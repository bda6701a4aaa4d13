// Deterministic synthetic LiDAR scans for tests and benchmarks.
//
// The reference ships no datasets; its replay drivers stream KITTI velodyne
// scans as packed float32 x,y,z,intensity (python_scripts/kitti_singlerobot_processor.py:164-185).
// This generator produces clouds of the same layout and of VLP-16 / HDL-64
// (KITTI-shape) / OS1-128 shape by casting rays into a procedural street scene
// along a smooth trajectory.  Everything is a pure function of (seed, scan_idx).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t hash3(uint64_t a, uint64_t b, uint64_t c) { return splitmix64(splitmix64(splitmix64(a) ^ b) ^ c); }
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  double next() { s = splitmix64(s); return u01(s); }
  double uni(double a, double b) { return a + (b - a) * next(); }
};

struct Box { double lo[3], hi[3]; };
struct Cyl { double cx, cy, r, z0, z1; };

constexpr double kGroundZ = -1.73;  // KITTI sensor height
constexpr double kBlock = 100.0;

void scene_block(uint64_t seed, long b, std::vector<Box>& boxes, std::vector<Cyl>& cyls) {
  Rng rng(hash3(seed, 0xB10C, (uint64_t)(b + (1l << 40))));
  const double x_begin = b * kBlock, x_end = x_begin + kBlock;
  for (int side = -1; side <= 1; side += 2) {
    double x = x_begin + rng.uni(0.0, 6.0);
    while (x < x_end) {
      double w = rng.uni(6.0, 18.0), depth = rng.uni(8.0, 20.0), h = rng.uni(6.0, 15.0), dist = rng.uni(7.0, 20.0);
      double x1 = std::fmin(x + w, x_end);
      Box bx;
      bx.lo[0] = x; bx.hi[0] = x1;
      if (side > 0) { bx.lo[1] = dist; bx.hi[1] = dist + depth; } else { bx.lo[1] = -dist - depth; bx.hi[1] = -dist; }
      bx.lo[2] = kGroundZ; bx.hi[2] = kGroundZ + h;
      boxes.push_back(bx);
      x = x1 + rng.uni(1.0, 6.0);
    }
  }
  for (int i = 0; i < 16; ++i) {
    Cyl c;
    c.cx = rng.uni(x_begin, x_end);
    c.cy = (rng.next() < 0.5 ? -1.0 : 1.0) * rng.uni(4.0, 7.5);
    c.r = rng.uni(0.15, 0.4);
    c.z0 = kGroundZ; c.z1 = kGroundZ + rng.uni(4.0, 9.0);
    cyls.push_back(c);
  }
  for (int i = 0; i < 6; ++i) {
    double cx = rng.uni(x_begin + 3.0, x_end - 3.0);
    double cy = (rng.next() < 0.5 ? -1.0 : 1.0) * rng.uni(2.8, 5.0);
    Box bx;
    bx.lo[0] = cx - 2.1; bx.hi[0] = cx + 2.1;
    bx.lo[1] = cy - 0.9; bx.hi[1] = cy + 0.9;
    bx.lo[2] = kGroundZ; bx.hi[2] = kGroundZ + 1.5;
    boxes.push_back(bx);
  }
  for (int i = 0; i < 5; ++i) {  // fences / hedges across the verge: surfaces facing along the road
    double cx = rng.uni(x_begin, x_end), side = (rng.next() < 0.5 ? -1.0 : 1.0);
    double y0 = rng.uni(5.0, 7.0), y1 = y0 + rng.uni(2.0, 6.0);
    Box bx;
    bx.lo[0] = cx - 0.15; bx.hi[0] = cx + 0.15;
    bx.lo[1] = side > 0 ? y0 : -y1; bx.hi[1] = side > 0 ? y1 : -y0;
    bx.lo[2] = kGroundZ; bx.hi[2] = kGroundZ + rng.uni(1.0, 2.5);
    boxes.push_back(bx);
  }
}

struct SensorModel { int rings, cols; double el_top_deg, el_bot_deg, max_range; };
SensorModel sensor_model(int sensor) {
  switch (sensor) {
    case 0: return {16, 1800, 15.0, -15.0, 100.0};    // VLP-16: 28,800 rays
    case 1: return {64, 1900, 2.0, -24.8, 120.0};     // HDL-64 (KITTI shape): 121,600 rays
    case 2: return {128, 2048, 22.5, -22.5, 120.0};   // OS1-128: 262,144 rays
    default: return {128, 8192, 22.5, -22.5, 120.0};  // OS1-128 "1M": 1,048,576 rays
  }
}

void trajectory_pose(uint64_t seed, int scan_idx, double* R /*row-major*/, double* t) {
  double s = 0.0;
  for (int j = 0; j < scan_idx; ++j) s += 0.35 + 0.2 * u01(hash3(seed, 0x57E9, (uint64_t)j));  // 0.35-0.55 m/scan: inside the 1 m-voxel convergence basin
  const double L = 180.0, A = 1.5;
  double y = A * std::sin(2 * M_PI * s / L);
  double dyds = A * 2 * M_PI / L * std::cos(2 * M_PI * s / L);
  double yaw = std::atan(dyds);
  double z = 0.02 * std::sin(0.7 * s);
  double roll = 0.005 * std::sin(1.3 * s), pitch = 0.005 * std::cos(0.9 * s);
  double cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch), cy = std::cos(yaw), sy = std::sin(yaw);
  // R = Rz(yaw) Ry(pitch) Rx(roll)
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
  t[0] = s; t[1] = y; t[2] = z;
}

}  // namespace

extern "C" {

int b2r_synth_num_rays(int sensor) {
  SensorModel m = sensor_model(sensor);
  return m.rings * m.cols;
}

// Sensor pose of scan `scan_idx` in the world frame, 4x4 row-major double.
void b2r_synth_pose(uint64_t seed, int scan_idx, double* T16) {
  double R[9], t[3];
  trajectory_pose(seed, scan_idx, R, t);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T16[r * 4 + c] = R[r * 3 + c];
    T16[r * 4 + 3] = t[r];
  }
  T16[12] = T16[13] = T16[14] = 0.0;
  T16[15] = 1.0;
}

// Writes up to num_rays points (sensor frame, float32 x,y,z,intensity) and returns the count.
int b2r_synth_scan(int sensor, uint64_t seed, int scan_idx, float* out_xyzi) {
  const SensorModel m = sensor_model(sensor);
  double R[9], t[3];
  trajectory_pose(seed, scan_idx, R, t);
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
  long b0 = (long)std::floor((t[0] - m.max_range) / kBlock), b1 = (long)std::floor((t[0] + m.max_range) / kBlock);
  for (long b = b0; b <= b1; ++b) scene_block(seed, b, boxes, cyls);
  const int nrays = m.rings * m.cols;
  std::vector<float> tmp((size_t)nrays * 4);
  std::vector<uint8_t> ok(nrays, 0);
  const double sigma = 0.02;
#pragma omp parallel for schedule(static)
  for (int ray = 0; ray < nrays; ++ray) {
    int col = ray / m.rings, ring = ray % m.rings;  // column-major firing order, like a spinning sensor
    double el = (m.el_top_deg + (m.el_bot_deg - m.el_top_deg) * (m.rings > 1 ? (double)ring / (m.rings - 1) : 0.0)) * M_PI / 180.0;
    double az = 2 * M_PI * (double)col / m.cols;
    double ds[3] = {std::cos(el) * std::cos(az), std::cos(el) * std::sin(az), std::sin(el)};
    double d[3] = {R[0] * ds[0] + R[1] * ds[1] + R[2] * ds[2], R[3] * ds[0] + R[4] * ds[1] + R[5] * ds[2],
                   R[6] * ds[0] + R[7] * ds[1] + R[8] * ds[2]};
    double best = m.max_range;
    bool hit = false;
    if (d[2] < -1e-9) {
      double tt = (kGroundZ - t[2]) / d[2];
      if (tt > 0.5 && tt < best) { best = tt; hit = true; }
    }
    for (const Box& bx : boxes) {
      double t0 = 0.0, t1 = best;
      bool miss = false;
      for (int a = 0; a < 3 && !miss; ++a) {
        if (std::fabs(d[a]) < 1e-12) {
          if (t[a] < bx.lo[a] || t[a] > bx.hi[a]) miss = true;
        } else {
          double ta = (bx.lo[a] - t[a]) / d[a], tb = (bx.hi[a] - t[a]) / d[a];
          if (ta > tb) { double s = ta; ta = tb; tb = s; }
          if (ta > t0) t0 = ta;
          if (tb < t1) t1 = tb;
          if (t0 > t1) miss = true;
        }
      }
      if (!miss && t0 > 0.5 && t0 < best) { best = t0; hit = true; }
    }
    for (const Cyl& c : cyls) {
      double ox = t[0] - c.cx, oy = t[1] - c.cy;
      double a = d[0] * d[0] + d[1] * d[1];
      if (a < 1e-12) continue;
      double bq = ox * d[0] + oy * d[1], cq = ox * ox + oy * oy - c.r * c.r;
      double disc = bq * bq - a * cq;
      if (disc < 0) continue;
      double tt = (-bq - std::sqrt(disc)) / a;
      if (tt > 0.5 && tt < best) {
        double z = t[2] + tt * d[2];
        if (z >= c.z0 && z <= c.z1) { best = tt; hit = true; }
      }
    }
    if (!hit) continue;
    uint64_t h = hash3(seed ^ 0xD0D0, (uint64_t)scan_idx, (uint64_t)ray);
    if (u01(h) < 0.02) continue;  // dropout
    double u1 = u01(splitmix64(h ^ 1)), u2 = u01(splitmix64(h ^ 2));
    double g = std::sqrt(-2.0 * std::log(u1 + 1e-300)) * std::cos(2 * M_PI * u2);
    double range = best + sigma * g;
    tmp[(size_t)ray * 4 + 0] = (float)(ds[0] * range);
    tmp[(size_t)ray * 4 + 1] = (float)(ds[1] * range);
    tmp[(size_t)ray * 4 + 2] = (float)(ds[2] * range);
    tmp[(size_t)ray * 4 + 3] = (float)u01(splitmix64(h ^ 3));
    ok[ray] = 1;
  }
  int n = 0;
  for (int ray = 0; ray < nrays; ++ray)
    if (ok[ray]) { std::memcpy(out_xyzi + (size_t)n * 4, &tmp[(size_t)ray * 4], 16); ++n; }
  return n;
}

}  // extern "C"
