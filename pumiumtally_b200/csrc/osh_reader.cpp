// Omega_h ".osh" mesh ingest without Omega_h.
//
// The reference constructs its mesh with Omega_h::binary::read(mesh_filename, world)
// (PumiTallyImpl.cpp:553-568; the test writes one with Omega_h::binary::write,
// test_pumi_tally_impl_methods.cpp:45-46).  Omega_h itself is an un-vendored dependency
// (scorec-v11.0.0, deps.yml:79) whose sources are not available here, and no sample .osh
// exists in the reference tree.  The stream layout below is therefore a RESTATEMENT FROM
// THE PUBLISHED FORMAT AS REMEMBERED, not checked against a file written by Omega_h:
//
//   mesh.osh/nparts   text integer (must be 1: one picpart holding the whole mesh, which is
//                     what the reference runs on, PumiTallyImpl.cpp:530-539)
//   mesh.osh/version  text integer (stream version; files older than 4 carry it in the stream)
//   mesh.osh/0.osh    little-endian stream:
//       u8[2] magic = a1 1a
//       [i32 version]            only when there is no version file
//       i8  is_compressed        arrays are zlib streams when set
//       meta: [i8 family (version >= 7)] i8 dim, i32 comm_size, i32 comm_rank, i8 parting,
//             i32 nghost_layers, i8 have_hints, [i32 naxes, f64[3*naxes]]
//       i32 nverts
//       for d = 1..dim: array<i32> down(d -> d-1); for d > 1 also array<i8> alignment codes
//       for d = 0..dim: i32 ntags, tags (string name, i8 ncomps, i8 type, array), [owners]
//       ... (class sets, parents: not needed)
//   array<T> = i32 count, then count*sizeof(T) raw bytes, or (compressed) i64 nbytes + zlib data
//   string   = i32 length + bytes
//
// Only what the walk needs is consumed: the three downward adjacencies and the vertex tag
// "coordinates".  Tet -> vertex sets come from composing tet->tri->edge->vert as SETS, so the
// alignment codes are skipped and nothing depends on Omega_h's template conventions; the
// vertices of a tet are emitted in ascending id (the engine derives face planes and their
// outward orientation from geometry, so local vertex order carries no meaning).  Element ids
// (the order of tets in the file) are preserved: they are what the tally is indexed by.
//
// The tag header differs between Omega_h versions (older ones carried two transfer/output
// flag bytes, the SCOREC fork adds a class-id list); the reader probes the known layouts and
// accepts the first one whose array header is consistent (count == nverts*ncomps and, when
// compressed, a zlib stream that inflates to exactly that size).  Anything inconsistent is
// reported as an error -- there is no silent guess.  The same goes for the two header details
// this restatement is least sure of (whether the stream repeats the version after the magic
// bytes, and from which version on the meta block starts with the mesh family): the expected
// reading is tried first, the alternatives after it, and a reading only counts if the whole
// stream up to the coordinates parses consistently under it.
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "tet_mesh.hpp"

namespace ptb {
namespace {

static_assert(__BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__, ".osh streams are little-endian; so must the host be");

constexpr unsigned char kMagic[2] = {0xa1, 0x1a};
constexpr int kTypeI8 = 0, kTypeI32 = 2, kTypeI64 = 3, kTypeF64 = 5;  // Omega_h_Type

struct Fail : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct Cursor {
  const unsigned char *p;
  size_t n, at = 0;
  bool compressed = false;

  void need(size_t k) const {
    if (k > n - at) throw Fail("unexpected end of .osh stream at byte " + std::to_string(at));
  }
  template <typename T>
  T get() {
    need(sizeof(T));
    T v;
    std::memcpy(&v, p + at, sizeof(T));
    at += sizeof(T);
    return v;
  }
  std::string str() {
    const int32_t len = get<int32_t>();
    if (len < 0 || len > 4096) throw Fail("implausible string length " + std::to_string(len));
    need(size_t(len));
    std::string s(reinterpret_cast<const char *>(p + at), size_t(len));
    at += size_t(len);
    return s;
  }
  // Reads array<T>; `expect` < 0 accepts any count.  out == nullptr skips the payload
  // (a compressed payload is still bounds-checked but not inflated).
  template <typename T>
  int64_t array(int64_t expect, std::vector<T> *out) {
    const int32_t count = get<int32_t>();
    if (count < 0 || (expect >= 0 && count != expect))
      throw Fail("array of " + std::to_string(count) + " entries where " + std::to_string(expect) +
                 " were expected (byte " + std::to_string(at - 4) + ")");
    const size_t bytes = size_t(count) * sizeof(T);
    if (compressed) {
      const int64_t cbytes = get<int64_t>();
      if (cbytes < 0 || uint64_t(cbytes) > uint64_t(compressBound(uLong(bytes))) + 64)
        throw Fail("implausible compressed size " + std::to_string(cbytes));
      need(size_t(cbytes));
      if (out) {
        out->resize(size_t(count));
        uLongf dst = uLongf(bytes);
        unsigned char dummy = 0;
        Bytef *dp = bytes ? reinterpret_cast<Bytef *>(out->data()) : &dummy;
        const int rc = uncompress(dp, &dst, p + at, uLong(cbytes));
        if (rc != Z_OK || dst != bytes) throw Fail("zlib stream does not inflate to the declared size");
      }
      at += size_t(cbytes);
    } else {
      need(bytes);
      if (out) {
        out->resize(size_t(count));
        if (bytes) std::memcpy(out->data(), p + at, bytes);
      }
      at += bytes;
    }
    return count;
  }
  void skip_typed_array(int type, int64_t expect) {
    switch (type) {
      case kTypeI8: array<int8_t>(expect, nullptr); break;
      case kTypeI32: array<int32_t>(expect, nullptr); break;
      case kTypeI64: array<int64_t>(expect, nullptr); break;
      case kTypeF64: array<double>(expect, nullptr); break;
      default: throw Fail("unknown tag type " + std::to_string(type));
    }
  }
};

bool read_text_int(const std::string &path, long *v) {
  std::ifstream f(path);
  return bool(f >> *v);
}

bool slurp(const std::string &path, std::vector<unsigned char> *buf) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return false;
  const std::streamsize n = f.tellg();
  f.seekg(0);
  buf->resize(size_t(n));
  return n == 0 || bool(f.read(reinterpret_cast<char *>(buf->data()), n));
}

// One vertex tag.  Returns true when it was "coordinates" (stored into *coords).
bool read_vertex_tag(Cursor &c, int64_t nverts, std::vector<double> *coords) {
  const std::string name = c.str();
  const int ncomps = c.get<int8_t>();
  const int type = c.get<int8_t>();
  if (ncomps < 1) throw Fail("tag '" + name + "' has " + std::to_string(ncomps) + " components");
  const int64_t expect = nverts * ncomps;
  const bool want = name == "coordinates";
  if (want && (type != kTypeF64 || ncomps != 3))
    throw Fail("'coordinates' is not a 3-component f64 tag (dim-3 meshes only)");
  // Layouts of what sits between the header and the array, newest first.
  enum { kClassIds, kDirect, kTwoFlagBytes };
  const size_t start = c.at;
  std::string why;
  for (int layout : {kDirect, kClassIds, kTwoFlagBytes}) {
    c.at = start;
    try {
      if (layout == kClassIds) {
        const int32_t nclass = c.get<int32_t>();
        if (nclass < 0 || nclass > (1 << 24)) throw Fail("implausible class-id count");
        if (nclass > 0) c.array<int32_t>(nclass, nullptr);
      } else if (layout == kTwoFlagBytes) {
        c.get<int8_t>();
        c.get<int8_t>();
      }
      if (want) {
        c.array<double>(expect, coords);
      } else if (c.compressed) {
        // validate by inflating: a wrong layout must not be accepted on a lucky header
        switch (type) {
          case kTypeI8: { std::vector<int8_t> t; c.array<int8_t>(expect, &t); break; }
          case kTypeI32: { std::vector<int32_t> t; c.array<int32_t>(expect, &t); break; }
          case kTypeI64: { std::vector<int64_t> t; c.array<int64_t>(expect, &t); break; }
          case kTypeF64: { std::vector<double> t; c.array<double>(expect, &t); break; }
          default: throw Fail("unknown tag type " + std::to_string(type));
        }
      } else {
        c.skip_typed_array(type, expect);
      }
      return want;
    } catch (const Fail &f) {
      why = f.what();
    }
  }
  throw Fail("cannot parse vertex tag '" + name + "': " + why);
}

// `version_in_stream`: an int32 version follows the magic bytes (bare streams, old directories).
// The meta block starts with the mesh family from version 7 on; `flip_family_rule` tries the other
// reading.  The caller tries the expected combination first; every later check (dimension,
// communicator, array sizes, the coordinates tag, four distinct vertices per tet) has to pass for a
// combination to be accepted.
void parse_stream(Cursor &c, long version, bool version_in_stream, bool flip_family_rule,
                  std::vector<double> *coords, std::vector<int32_t> *t2v) {
  c.need(2);
  if (c.p[0] != kMagic[0] || c.p[1] != kMagic[1]) throw Fail("not an Omega_h binary stream (magic bytes)");
  c.at = 2;
  if (version_in_stream) version = c.get<int32_t>();
  if (version < 1 || version > 64) throw Fail("implausible stream version " + std::to_string(version));
  const int comp = c.get<int8_t>();
  if (comp != 0 && comp != 1) throw Fail("implausible compression flag");
  c.compressed = comp != 0;
  if ((version >= 7) != flip_family_rule) {
    const int family = c.get<int8_t>();
    if (family != 0) throw Fail("mesh family is not simplex");
  }
  const int dim = c.get<int8_t>();
  if (dim != 3) throw Fail("mesh dimension is " + std::to_string(dim) + "; the tally needs a tet mesh (dim 3)");
  const int32_t comm_size = c.get<int32_t>();
  const int32_t comm_rank = c.get<int32_t>();
  if (comm_size != 1 || comm_rank != 0)
    throw Fail("stream was written by rank " + std::to_string(comm_rank) + " of " + std::to_string(comm_size) +
               "; a single-part mesh is required");
  const int parting = c.get<int8_t>();
  const int32_t nghost = c.get<int32_t>();
  if (parting < 0 || parting > 2 || nghost < 0) throw Fail("implausible partition metadata");
  if (c.get<int8_t>() != 0) {  // recursive-inertial-bisection hints
    const int32_t naxes = c.get<int32_t>();
    if (naxes < 0 || naxes > 64) throw Fail("implausible hint count");
    for (int i = 0; i < 3 * naxes; ++i) c.get<double>();
  }
  const int64_t nverts = c.get<int32_t>();
  if (nverts < 4) throw Fail("mesh has " + std::to_string(nverts) + " vertices");

  std::vector<int32_t> ev2v, fe2e, rf2f;
  const int64_t ne2 = c.array<int32_t>(-1, &ev2v);
  if (ne2 % 2) throw Fail("edge->vertex array has odd length");
  const int64_t nf3 = c.array<int32_t>(-1, &fe2e);
  if (nf3 % 3) throw Fail("triangle->edge array length is not a multiple of 3");
  c.array<int8_t>(nf3, nullptr);
  const int64_t nr4 = c.array<int32_t>(-1, &rf2f);
  if (nr4 % 4 || nr4 == 0) throw Fail("tet->triangle array length is not a positive multiple of 4");
  c.array<int8_t>(nr4, nullptr);
  const int64_t nedges = ne2 / 2, ntris = nf3 / 3, ntets = nr4 / 4;
  for (int32_t v : ev2v) if (v < 0 || v >= nverts) throw Fail("edge references vertex out of range");
  for (int32_t e : fe2e) if (e < 0 || e >= nedges) throw Fail("triangle references edge out of range");
  for (int32_t f : rf2f) if (f < 0 || f >= ntris) throw Fail("tet references triangle out of range");

  const int32_t ntags = c.get<int32_t>();
  if (ntags < 1 || ntags > 4096) throw Fail("implausible vertex tag count " + std::to_string(ntags));
  bool have = false;
  for (int i = 0; i < ntags && !have; ++i) have = read_vertex_tag(c, nverts, coords);
  if (!have) throw Fail("no 'coordinates' tag on the vertices");

  // tet -> vertex set through tet -> tri -> edge -> vert
  t2v->resize(size_t(ntets) * 4);
  for (int64_t r = 0; r < ntets; ++r) {
    int32_t v[24];
    int n = 0;
    for (int f = 0; f < 4; ++f) {
      const int32_t tri = rf2f[4 * r + f];
      for (int e = 0; e < 3; ++e) {
        const int32_t edge = fe2e[3 * size_t(tri) + e];
        v[n++] = ev2v[2 * size_t(edge)];
        v[n++] = ev2v[2 * size_t(edge) + 1];
      }
    }
    std::sort(v, v + n);
    n = int(std::unique(v, v + n) - v);
    if (n != 4) throw Fail("tet " + std::to_string(r) + " touches " + std::to_string(n) + " vertices");
    for (int k = 0; k < 4; ++k) (*t2v)[4 * size_t(r) + k] = v[k];
  }
}

}  // namespace

// `path` is the .osh directory (the reference's calling convention, PumiTally.h:40-47) or a
// bare stream file that carries its own version word.
bool read_osh_mesh(const std::string &path, std::vector<double> *coords, std::vector<int32_t> *t2v,
                   std::string *err) {
  std::string stream_file = path;
  long version = -1;
  std::ifstream probe(path + "/nparts");
  if (probe) {
    long nparts = 0;
    if (!(probe >> nparts)) { *err = path + "/nparts is unreadable"; return false; }
    if (nparts != 1) {
      *err = path + " holds " + std::to_string(nparts) +
             " parts; the tally loads the whole mesh on every rank -- write it from a single rank";
      return false;
    }
    if (!read_text_int(path + "/version", &version)) version = -1;
    stream_file = path + "/0.osh";
  } else if (std::ifstream(path + "/0.osh")) {
    stream_file = path + "/0.osh";
  }
  std::vector<unsigned char> buf;
  if (!slurp(stream_file, &buf)) { *err = "cannot read " + stream_file; return false; }
  // layout hypotheses, most likely first: the version comes from the version file when there is one
  // (else from the stream), family byte by the version rule; then the alternatives
  const bool have_file_version = version >= 0;
  std::string first_error;
  for (int attempt = 0; attempt < 4; ++attempt) {
    const bool flip_version_place = attempt >= 2, flip_family = attempt & 1;
    if (flip_version_place && !have_file_version) break;  // without a version file the stream must carry it
    const bool version_in_stream = have_file_version == flip_version_place;
    Cursor c{buf.data(), buf.size()};
    try {
      parse_stream(c, version, version_in_stream, flip_family, coords, t2v);
      return true;
    } catch (const Fail &f) {
      if (first_error.empty()) first_error = f.what();
    }
  }
  *err = "Omega_h mesh " + path + ": " + first_error;
  return false;
}

}  // namespace ptb
