// Gmsh .msh ingest (format 2.2 and 4.1, ASCII or binary): nodes + 4-node tetrahedra (element type 4).
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

// ---- binary flavour ------------------------------------------------------------------------------
// 2.2: "$Nodes\nN\n" then N x {int id, double xyz[3]};  "$Elements\nM\n" then blocks
//      {int type, int count, int ntags} followed by count x {int id, int tags[ntags], int nodes[nn]}.
// 4.1: "$Nodes\n" then size_t {blocks, nodes, min, max}; per block {int dim, int tag, int parametric,
//      size_t n}, size_t ids[n], double xyz[n][3 (+dim if parametric)];  "$Elements\n" then size_t
//      {blocks, elements, min, max}; per block {int dim, int tag, int type, size_t n}, n x {size_t id,
//      size_t nodes[nn]}.  Both start with the int 1 after the format line (byte-order check).
template <typename T>
bool get(std::istream &f, T *v, size_t n = 1) {
  return bool(f.read(reinterpret_cast<char *>(v), std::streamsize(n * sizeof(T))));
}

bool seek_line(std::istream &f, const std::string &name) {  // positions the stream right after "name\n"
  return seek_section(f, name);
}

bool read_gmsh_binary(const std::string &path, double version, int data_size, std::vector<double> *coords,
                      std::vector<int32_t> *t2v, std::string *err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { *err = "cannot open " + path; return false; }
  if (data_size != 8) { *err = "binary Gmsh file with data size != 8"; return false; }
  if (!seek_line(f, "$MeshFormat")) { *err = "not a Gmsh file (no $MeshFormat): " + path; return false; }
  std::string fmt_line;
  std::getline(f, fmt_line);
  int one = 0;
  if (!get(f, &one) || one != 1) { *err = "binary Gmsh file written with the other byte order"; return false; }
  std::unordered_map<long long, int32_t> id2idx;
  if (!seek_line(f, "$Nodes")) { *err = "Gmsh file has no $Nodes"; return false; }
  if (version < 4.0) {
    long long n = 0;
    f >> n;
    f.ignore(1);  // the newline after the count
    if (!f || n < 0 || n > kMaxCount) { *err = "implausible node count in $Nodes"; return false; }
    coords->reserve(size_t(3) * n);
    for (long long i = 0; i < n; ++i) {
      int id;
      double x[3];
      if (!get(f, &id) || !get(f, x, 3)) { *err = "truncated $Nodes"; return false; }
      id2idx[id] = int32_t(i);
      coords->insert(coords->end(), x, x + 3);
    }
    if (!seek_line(f, "$Elements")) { *err = "Gmsh file has no $Elements"; return false; }
    long long ne = 0;
    f >> ne;
    f.ignore(1);
    if (!f || ne < 0 || ne > kMaxCount) { *err = "implausible element count in $Elements"; return false; }
    for (long long done = 0; done < ne;) {
      int head[3];  // type, number of elements that follow, number of tags
      if (!get(f, head, 3)) { *err = "truncated $Elements"; return false; }
      const int nn = nodes_of_type(head[0]);
      if (nn < 0 || head[1] < 1 || head[1] > ne - done || head[2] < 0 || head[2] > 64) { *err = "damaged element block in $Elements"; return false; }
      std::vector<int> rec(size_t(1 + head[2] + nn));
      for (int k = 0; k < head[1]; ++k) {
        if (!get(f, rec.data(), rec.size())) { *err = "truncated $Elements"; return false; }
        if (head[0] == 4)
          for (int q = 0; q < 4; ++q) {
            auto it = id2idx.find(rec[size_t(1 + head[2] + q)]);
            if (it == id2idx.end()) { *err = "tet references unknown node"; return false; }
            t2v->push_back(it->second);
          }
      }
      done += head[1];
    }
  } else {
    size_t h[4];
    if (!get(f, h, 4) || h[0] > size_t(kMaxCount) || h[1] > size_t(kMaxCount)) { *err = "implausible counts in $Nodes"; return false; }
    coords->reserve(3 * h[1]);
    for (size_t b = 0; b < h[0]; ++b) {
      int ent[3];
      size_t nb = 0;
      if (!get(f, ent, 3) || !get(f, &nb) || nb > h[1] || ent[0] < 0 || ent[0] > 3) { *err = "damaged node block in $Nodes"; return false; }
      std::vector<size_t> ids(nb);
      if (nb && !get(f, ids.data(), nb)) { *err = "truncated $Nodes"; return false; }
      const size_t per = 3 + (ent[2] ? size_t(ent[0]) : 0);
      std::vector<double> xyz(per * nb);
      if (nb && !get(f, xyz.data(), xyz.size())) { *err = "truncated $Nodes"; return false; }
      for (size_t i = 0; i < nb; ++i) {
        id2idx[(long long)ids[i]] = int32_t(coords->size() / 3);
        coords->insert(coords->end(), xyz.begin() + long(per * i), xyz.begin() + long(per * i + 3));
      }
    }
    if (!seek_line(f, "$Elements")) { *err = "Gmsh file has no $Elements"; return false; }
    if (!get(f, h, 4) || h[0] > size_t(kMaxCount) || h[1] > size_t(kMaxCount)) { *err = "implausible counts in $Elements"; return false; }
    for (size_t b = 0; b < h[0]; ++b) {
      int ent[3];
      size_t nb = 0;
      if (!get(f, ent, 3) || !get(f, &nb) || nb > h[1]) { *err = "damaged element block in $Elements"; return false; }
      const int nn = nodes_of_type(ent[2]);
      if (nn < 0) { *err = "unsupported Gmsh element type " + std::to_string(ent[2]); return false; }
      std::vector<size_t> rec(size_t(1 + nn));
      for (size_t i = 0; i < nb; ++i) {
        if (!get(f, rec.data(), rec.size())) { *err = "truncated $Elements"; return false; }
        if (ent[2] == 4)
          for (int q = 0; q < 4; ++q) {
            auto it = id2idx.find((long long)rec[size_t(1 + q)]);
            if (it == id2idx.end()) { *err = "tet references unknown node"; return false; }
            t2v->push_back(it->second);
          }
      }
    }
  }
  if (t2v->empty()) { *err = "Gmsh file contains no 4-node tetrahedra: " + path; return false; }
  return true;
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
  coords->clear();
  t2v->clear();
  if (file_type != 0) return read_gmsh_binary(path, version, data_size, coords, t2v, err);
  std::unordered_map<long long, int32_t> id2idx;
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
