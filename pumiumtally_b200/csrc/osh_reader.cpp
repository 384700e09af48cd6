// Omega_h ".osh" mesh directory reader -- placeholder until the stream format
// is implemented.
#include "tet_mesh.hpp"

namespace ptb {

bool read_osh_mesh(const std::string &dir, std::vector<double> *, std::vector<int32_t> *,
                   std::string *err) {
  *err = "Omega_h .osh ingest is not available in this build (" + dir +
         "); convert the mesh with pumiumtally_b200.mesh.save_raw_mesh or use box:nx,ny,nz";
  return false;
}

}  // namespace ptb
