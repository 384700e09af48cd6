#include "vtk_writer.hpp"

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sys/stat.h>

namespace ptb {
namespace {

bool make_dir(const std::string &p, std::string *err) {
  if (mkdir(p.c_str(), 0777) == 0 || errno == EEXIST) return true;
  *err = "cannot create directory " + p + ": " + std::strerror(errno);
  return false;
}

struct Block {
  const void *data;
  uint64_t bytes;
};

}  // namespace

bool write_vtk_dataset(const std::string &path, const HostMesh &m, const std::vector<double> &flux,
                       const std::vector<double> &volume, int rank, int nranks, std::string *err) {
  if (!make_dir(path, err) || !make_dir(path + "/pieces", err)) return false;
  const uint64_t V = uint64_t(m.nverts), E = uint64_t(m.ntets);
  std::vector<int32_t> offsets(E);
  for (uint64_t e = 0; e < E; ++e) offsets[e] = int32_t(4 * (e + 1));
  std::vector<uint8_t> types(E, 10);  // VTK_TETRA
  // connectivity in the caller's element order (the cell data arrays already are)
  const std::vector<int32_t> conn = m.to_original(m.t2v.data(), 4);
  const Block blocks[6] = {{m.coords.data(), V * 24}, {conn.data(), E * 16},
                           {offsets.data(), E * 4},   {types.data(), E},
                           {flux.data(), E * 8},      {volume.data(), E * 8}};
  uint64_t off[6], acc = 0;
  for (int i = 0; i < 6; ++i) { off[i] = acc; acc += 8 + blocks[i].bytes; }

  const std::string piece = path + "/pieces/piece_" + std::to_string(rank) + ".vtu";
  std::ofstream f(piece, std::ios::binary);
  if (!f) { *err = "cannot open " + piece; return false; }
  f << "<?xml version=\"1.0\"?>\n"
    << "<VTKFile type=\"UnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
    << "<UnstructuredGrid>\n"
    << "<Piece NumberOfPoints=\"" << V << "\" NumberOfCells=\"" << E << "\">\n"
    << "<Points>\n<DataArray type=\"Float64\" Name=\"coordinates\" NumberOfComponents=\"3\" format=\"appended\" offset=\"" << off[0] << "\"/>\n</Points>\n"
    << "<Cells>\n"
    << "<DataArray type=\"Int32\" Name=\"connectivity\" format=\"appended\" offset=\"" << off[1] << "\"/>\n"
    << "<DataArray type=\"Int32\" Name=\"offsets\" format=\"appended\" offset=\"" << off[2] << "\"/>\n"
    << "<DataArray type=\"UInt8\" Name=\"types\" format=\"appended\" offset=\"" << off[3] << "\"/>\n"
    << "</Cells>\n"
    << "<CellData>\n"
    << "<DataArray type=\"Float64\" Name=\"flux\" NumberOfComponents=\"1\" format=\"appended\" offset=\"" << off[4] << "\"/>\n"
    << "<DataArray type=\"Float64\" Name=\"volume\" NumberOfComponents=\"1\" format=\"appended\" offset=\"" << off[5] << "\"/>\n"
    << "</CellData>\n"
    << "</Piece>\n</UnstructuredGrid>\n<AppendedData encoding=\"raw\">\n_";
  for (int i = 0; i < 6; ++i) {
    f.write(reinterpret_cast<const char *>(&blocks[i].bytes), 8);
    f.write(reinterpret_cast<const char *>(blocks[i].data), std::streamsize(blocks[i].bytes));
  }
  f << "\n</AppendedData>\n</VTKFile>\n";
  if (!f) { *err = "write failed: " + piece; return false; }
  f.close();

  if (rank == 0) {
    std::ofstream p(path + "/pieces.pvtu");
    if (!p) { *err = "cannot open " + path + "/pieces.pvtu"; return false; }
    p << "<?xml version=\"1.0\"?>\n"
      << "<VTKFile type=\"PUnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
      << "<PUnstructuredGrid GhostLevel=\"0\">\n"
      << "<PPoints>\n<PDataArray type=\"Float64\" Name=\"coordinates\" NumberOfComponents=\"3\"/>\n</PPoints>\n"
      << "<PCellData>\n"
      << "<PDataArray type=\"Float64\" Name=\"flux\" NumberOfComponents=\"1\"/>\n"
      << "<PDataArray type=\"Float64\" Name=\"volume\" NumberOfComponents=\"1\"/>\n"
      << "</PCellData>\n";
    for (int r = 0; r < nranks; ++r) p << "<Piece Source=\"pieces/piece_" << r << ".vtu\"/>\n";
    p << "</PUnstructuredGrid>\n</VTKFile>\n";
  }
  return true;
}

}  // namespace ptb
