// PCD keyframe files for the C++ host side: the reference stores every keyframe cloud with
// pcl::io::savePCDFileBinary and reads it back with pcl::io::loadPCDFile
// (/root/reference/src/mrg_slam/keyframe.cpp:108-110,195-197).  Header-only, no PCL needed; the layout rules are those of
// mrg_slam_b200/pcd.py (PCL describes the struct padding of PointXYZI as `_` fields; DATA binary or ascii).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "registration.hpp"

namespace b2r {

inline bool save_pcd_binary(const std::string& path, const PointCloud& cloud) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  const size_t n = cloud.size();
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z _ intensity _\nSIZE 4 4 4 1 4 1\nTYPE F F F U F U\n"
       "COUNT 1 1 1 4 1 12\nWIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  static_assert(sizeof(PointXYZI) == 32, "PointXYZI must keep pcl's 32-byte layout");
  f.write(reinterpret_cast<const char*>(cloud.points.data()), (std::streamsize)(n * sizeof(PointXYZI)));
  return (bool)f;
}

inline bool load_pcd(const std::string& path, PointCloud& cloud) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::vector<std::string> fields, types;
  std::vector<int> sizes, counts;
  size_t n = 0, width = 0, height = 1;
  std::string mode, line;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string key, v;
    ss >> key;
    if (key == "FIELDS") while (ss >> v) fields.push_back(v);
    else if (key == "SIZE") while (ss >> v) sizes.push_back(std::stoi(v));
    else if (key == "TYPE") while (ss >> v) types.push_back(v);
    else if (key == "COUNT") while (ss >> v) counts.push_back(std::stoi(v));
    else if (key == "WIDTH") ss >> width;
    else if (key == "HEIGHT") ss >> height;
    else if (key == "POINTS") ss >> n;
    else if (key == "DATA") { ss >> mode; break; }
  }
  if (fields.empty() || sizes.size() != fields.size() || types.size() != fields.size()) return false;
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (n == 0) n = width * height;
  std::map<std::string, int> want{{"x", 0}, {"y", 1}, {"z", 2}, {"intensity", 3}};
  cloud.points.assign(n, PointXYZI());
  auto put = [&](PointXYZI& p, int which, float v) {
    if (which == 0) p.x = v; else if (which == 1) p.y = v; else if (which == 2) p.z = v; else p.intensity = v;
  };
  if (mode == "binary") {
    size_t rec = 0;
    for (size_t i = 0; i < fields.size(); ++i) rec += (size_t)sizes[i] * counts[i];
    std::vector<char> buf(rec * n);
    f.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)f.gcount() != buf.size()) return false;
    size_t off = 0;
    for (size_t i = 0; i < fields.size(); ++i) {
      auto it = want.find(fields[i]);
      if (it != want.end() && counts[i] == 1) {
        for (size_t k = 0; k < n; ++k) {
          const char* src = buf.data() + k * rec + off;
          float v = 0.f;
          if (types[i] == "F" && sizes[i] == 4) std::memcpy(&v, src, 4);
          else if (types[i] == "F" && sizes[i] == 8) { double d; std::memcpy(&d, src, 8); v = (float)d; }
          else if (sizes[i] == 1) v = types[i] == "I" ? (float)*(const int8_t*)src : (float)*(const uint8_t*)src;
          else if (sizes[i] == 2) { uint16_t u; std::memcpy(&u, src, 2); v = types[i] == "I" ? (float)(int16_t)u : (float)u; }
          else if (sizes[i] == 4) { uint32_t u; std::memcpy(&u, src, 4); v = types[i] == "I" ? (float)(int32_t)u : (float)u; }
          put(cloud.points[k], it->second, v);
        }
      }
      off += (size_t)sizes[i] * counts[i];
    }
    return true;
  }
  if (mode == "ascii") {
    for (size_t k = 0; k < n; ++k) {
      if (!std::getline(f, line)) return false;
      std::istringstream ss(line);
      for (size_t i = 0; i < fields.size(); ++i)
        for (int c = 0; c < counts[i]; ++c) {
          double d;
          ss >> d;
          auto it = want.find(fields[i]);
          if (it != want.end() && counts[i] == 1) put(cloud.points[k], it->second, (float)d);
        }
    }
    return true;
  }
  return false;  // binary_compressed: not written by the reference
}

}  // namespace b2r
