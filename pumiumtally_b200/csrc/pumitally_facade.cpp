// pumitally::PumiTally -- the C++ class the OpenMC fork links against,
// implemented over the C ABI (include/pumitally_c.h).  Mirrors the reference
// facade's behaviour: forward each call, accumulate wall-clock time per phase,
// print the [TIME] block after WriteTallyResults
// (reference: src/pumitally/PumiTally.cpp:16-60, PumiTallyImpl.cpp:22-29).
#include "pumitally/PumiTally.h"

#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <stdexcept>

#include "pumitally/PumiTallyExtras.h"
#include "pumitally_c.h"

namespace pumitally {

// reference: TallyTimes, PumiTallyImpl.h:18-27
struct TallyTimes {
  double initialization_time = 0.0;
  double total_time_to_tally = 0.0;
  double vtk_file_write_time = 0.0;
  void PrintTimes() const {
    printf("\n");
    printf("[TIME] Initialization time     : %f seconds\n", initialization_time);
    printf("[TIME] Total time to tally     : %f seconds\n", total_time_to_tally);
    printf("[TIME] VTK file write time     : %f seconds\n", vtk_file_write_time);
    printf("[TIME] Total PUMI-Tally time   : %f seconds\n",
           initialization_time + total_time_to_tally + vtk_file_write_time);
  }
};

struct PumiTallyImpl {
  pumitally_engine *engine = nullptr;
  TallyTimes tally_times;
  ~PumiTallyImpl() { pumitally_destroy(engine); }
};

// The reference header gives no access to what is behind pimpl_ and must stay as it is (it IS the ABI),
// so the additive calls of PumiTallyExtras.h find the engine of a PumiTally object through this table.
namespace {
std::mutex registry_mutex;
std::map<const PumiTally *, pumitally_engine *> &registry() {
  static std::map<const PumiTally *, pumitally_engine *> r;
  return r;
}
}  // namespace

pumitally_engine *engine_of(const PumiTally &tally) {
  std::lock_guard<std::mutex> lk(registry_mutex);
  auto it = registry().find(&tally);
  return it == registry().end() ? nullptr : it->second;
}

namespace {
struct ScopedTimer {
  double &acc;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit ScopedTimer(double &a) : acc(a) {}
  ~ScopedTimer() {
    acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
};
}  // namespace

PumiTally::PumiTally(const std::string &mesh_filename, const int32_t num_particles, int &argc,
                     char **&argv)
    : pimpl_(std::make_unique<PumiTallyImpl>()) {
  pimpl_->engine = pumitally_create(mesh_filename.c_str(), num_particles, &argc, &argv);
  // The reference prints and carries on (PumiTallyImpl.cpp:558-565); a tally object without an
  // engine would silently turn every later call into a no-op and write an empty result, so this
  // implementation refuses to exist instead.
  if (!pimpl_->engine)
    throw std::runtime_error("pumitally::PumiTally: engine construction failed (mesh \"" + mesh_filename +
                             "\" unreadable, or no usable CUDA device; this library has no CPU fallback)");
  std::lock_guard<std::mutex> lk(registry_mutex);
  registry()[this] = pimpl_->engine;
}

PumiTally::~PumiTally() {
  {
    std::lock_guard<std::mutex> lk(registry_mutex);
    registry().erase(this);
  }
  pimpl_.reset(nullptr);
}

void PumiTally::CopyInitialPosition(double *init_particle_positions, const std::int32_t size) const {
  ScopedTimer t(pimpl_->tally_times.initialization_time);
  if (pumitally_copy_initial_position(pimpl_->engine, init_particle_positions, size))
    throw std::runtime_error("pumitally::PumiTally::CopyInitialPosition failed (see the message above)");
}

void PumiTally::MoveToNextLocation(double *particle_origin, double *particle_destinations,
                                   int8_t *flying, double *weights, const std::int32_t size) const {
  ScopedTimer t(pimpl_->tally_times.total_time_to_tally);
  if (pumitally_move_to_next_location(pimpl_->engine, particle_origin, particle_destinations, flying,
                                      weights, size))
    throw std::runtime_error("pumitally::PumiTally::MoveToNextLocation failed (see the message above)");
  // kernels of the last particle range may still be in flight here; the time
  // they take is charged to the next call that waits on them, as with the
  // reference's un-fenced Kokkos launches (PumiTallyImpl.cpp:146-148).
}

void PumiTally::WriteTallyResults() const {
  {
    ScopedTimer t(pimpl_->tally_times.vtk_file_write_time);
    // lost walks: get_stats prints the reference's "Not all particles are found" line
    // (PumiTallyImpl.cpp:455-458) -- a result with lost particles must not pass silently
    pumitally_stats st;
    if (pumitally_get_stats(pimpl_->engine, &st) == 0 && st.lost)
      fprintf(stderr, "[pumitally] WARNING: %llu walks were cut short by the crossing limit or had unusable "
                      "(non-finite) inputs; the tally misses their contributions\n", (unsigned long long)st.lost);
    if (pumitally_write_tally_results(pimpl_->engine))
      throw std::runtime_error("pumitally::PumiTally::WriteTallyResults failed (see the message above)");
  }
  pimpl_->tally_times.PrintTimes();
}

}  // namespace pumitally
