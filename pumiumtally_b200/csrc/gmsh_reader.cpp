// Gmsh .msh ingest (ASCII, format 2.2 and 4.1): nodes + 4-node tetrahedra (element type 4).
//
// The reference only reads Omega_h ".osh" directories (PumiTallyImpl.cpp:553-568) and asks users
// to convert their Gmsh mesh with `msh2osh` first (README.md:115-126).  Reading the .msh file
// directly removes that step and the Omega_h dependency for the common workflow; all other
// element types in the file (points, lines, triangles of the boundary) are ignored.
#include <cstdint>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "tet_mesh.hpp"

namespace ptb {
namespace {

constexpr long long kMaxCount = 2000000000LL;  // ids are 32-bit in the engine

bool seek_section(std::istream &f, const std::string &name) {
  std::string line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line == name) return true;
  }
  return false;
}

// nodes of each Gmsh element type we may have to skip in format 2.2
int nodes_of_type(int t) {
  static const int n[] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1, 8, 20, 15, 13};
  return (t >= 1 && t <= 19) ? n[t] : -1;
}

}  // namespace

bool read_gmsh_mesh(const std::string &path, std::vector<double> *coords, std::vector<int32_t> *t2v,
                    std::string *err) {
  std::ifstream f(path);
  if (!f) { *err = "cannot open " + path; return false; }
  if (!seek_section(f, "$MeshFormat")) { *err = "not a Gmsh file (no $MeshFormat): " + path; return false; }
  double version = 0;
  int file_type = 0, data_size = 0;
  f >> version >> file_type >> data_size;
  if (file_type != 0) { *err = "binary Gmsh files are not supported; export ASCII (-format msh2 or msh4 without -bin)"; return false; }
  std::unordered_map<long long, int32_t> id2idx;
  coords->clear();
  t2v->clear();
  if (!seek_section(f, "$Nodes")) { *err = "Gmsh file has no $Nodes"; return false; }
  if (version < 4.0) {
    long long n = 0;
    f >> n;
    if (!f || n < 0 || n > kMaxCount) { *err = "implausible node count in $Nodes"; return false; }
    coords->reserve(size_t(3) * n);
    for (long long i = 0; i < n; ++i) {
      long long id; double x, y, z;
      f >> id >> x >> y >> z;
      id2idx[id] = int32_t(i);
      coords->push_back(x); coords->push_back(y); coords->push_back(z);
    }
    if (!f) { *err = "truncated $Nodes"; return false; }
    if (!seek_section(f, "$Elements")) { *err = "Gmsh file has no $Elements"; return false; }
    long long ne = 0;
    f >> ne;
    if (!f || ne < 0 || ne > kMaxCount) { *err = "implausible element count in $Elements"; return false; }
    for (long long e = 0; e < ne; ++e) {
      long long id; int type, ntags;
      f >> id >> type >> ntags;
      if (!f || ntags < 0 || ntags > 64) { *err = "damaged element record in $Elements"; return false; }
      for (int k = 0; k < ntags; ++k) { long long tag; f >> tag; }
      const int nn = nodes_of_type(type);
      if (nn < 0) { *err = "unsupported Gmsh element type " + std::to_string(type); return false; }
      long long v[27];
      for (int k = 0; k < nn; ++k) f >> v[k];
      if (type == 4)
        for (int k = 0; k < 4; ++k) {
          auto it = id2idx.find(v[k]);
          if (it == id2idx.end()) { *err = "tet references unknown node " + std::to_string(v[k]); return false; }
          t2v->push_back(it->second);
        }
    }
    if (!f) { *err = "truncated $Elements"; return false; }
  } else {
    long long nblocks = 0, n = 0, mintag = 0, maxtag = 0;
    f >> nblocks >> n >> mintag >> maxtag;
    if (!f || nblocks < 0 || n < 0 || nblocks > kMaxCount || n > kMaxCount) { *err = "implausible counts in $Nodes"; return false; }
    coords->reserve(size_t(3) * n);
    for (long long b = 0; b < nblocks; ++b) {
      int edim, etag, parametric; long long nb;
      f >> edim >> etag >> parametric >> nb;
      if (!f || nb < 0 || nb > n) { *err = "implausible node block in $Nodes"; return false; }
      std::vector<long long> ids(nb);
      for (long long i = 0; i < nb; ++i) f >> ids[i];
      for (long long i = 0; i < nb; ++i) {
        double x, y, z;
        f >> x >> y >> z;
        if (parametric) { double u; for (int k = 0; k < edim; ++k) f >> u; }
        id2idx[ids[i]] = int32_t(coords->size() / 3);
        coords->push_back(x); coords->push_back(y); coords->push_back(z);
      }
    }
    if (!f) { *err = "truncated $Nodes"; return false; }
    if (!seek_section(f, "$Elements")) { *err = "Gmsh file has no $Elements"; return false; }
    long long ne = 0;
    f >> nblocks >> ne >> mintag >> maxtag;
    if (!f || nblocks < 0 || ne < 0 || nblocks > kMaxCount || ne > kMaxCount) { *err = "implausible counts in $Elements"; return false; }
    for (long long b = 0; b < nblocks; ++b) {
      int edim, etag, type; long long nb;
      f >> edim >> etag >> type >> nb;
      if (!f || nb < 0 || nb > ne) { *err = "implausible element block in $Elements"; return false; }
      const int nn = nodes_of_type(type);
      if (nn < 0) { *err = "unsupported Gmsh element type " + std::to_string(type); return false; }
      for (long long i = 0; i < nb; ++i) {
        long long id, v[27];
        f >> id;
        for (int k = 0; k < nn; ++k) f >> v[k];
        if (type == 4)
          for (int k = 0; k < 4; ++k) {
            auto it = id2idx.find(v[k]);
            if (it == id2idx.end()) { *err = "tet references unknown node " + std::to_string(v[k]); return false; }
            t2v->push_back(it->second);
          }
      }
    }
    if (!f) { *err = "truncated $Elements"; return false; }
  }
  if (t2v->empty()) { *err = "Gmsh file contains no 4-node tetrahedra: " + path; return false; }
  return true;
}

}  // namespace ptb
