// Compact (vertex-indexed) tet layout and the edge-function form of the exit test.
//
// The plane records of walk_core.cuh cost 128 B per tet: a 1 M-tet table is as large as the L2 and
// every second crossing goes to HBM.  Here the walk reads, per crossing,
//     one 32-byte TetLinks record   (who is behind each face, which vertex comes with it)
//   + one 32-byte VertexRec         (the single vertex the next tet does not share with this one)
// so the data a walk touches is 32 B per tet + 32 B per vertex (c2: 32 MB + 5.6 MB) and stays
// L2-resident; the other three vertices of the next tet are already in registers.
//
// Exit test.  With P_i = v_i - o (o = fixed ray origin, u = target - o) and the edge functions
//     m(i,j) = u . (P_i x P_j)
// the line pierces a triangle (i,j,k) iff m(i,j), m(j,k), m(k,i) have one sign; their sum is
// u . n (n = the triangle's normal in that orientation) and the crossing parameter is
//     t = P_i . (P_j x P_k) / (m(i,j) + m(j,k) + m(k,i)).
// The walk keeps the face it entered through as an ordered triple (a,b,c) with all three edge
// functions > 0.  In the next tet only the three edge functions of the NEW vertex d are
// evaluated, s_r = m(d, r) = (u x P_d) . P_r, and their signs alone pick the exit face:
//     (a,b,d) iff s_a >= 0 > s_b      (b,c,d) iff s_b >= 0 > s_c      (c,a,d) iff s_c >= 0 > s_a
// The exit triple inherits the sign of one edge function from the entry triple and of two from
// the s_r, so every edge is classified once per ray and all tets around it see the same answer:
// the walk is watertight by construction, with no canonical vertex ordering and no division
// except the one for t.  This replaces the arithmetic of the reference's external tracer
// (PumiTallyImpl.cpp:454; contract PumiTallyImpl.h:74-85) like the plane form does; tally,
// clipping and advance are the shared advance() of walk_core.cuh.
//
// The first tet of a ray has no entry face: all four vertices are read (TetStart, one 128-byte
// line streamed past the L2) and the face whose three edge functions are all > 0 is the exit.
//
// Degenerate rays.  An edge function is exactly zero when the ray and a mesh edge are coplanar
// (axis-parallel tracks in a structured mesh, a track inside a face plane, a particle sitting on
// an edge).  Sign tests cannot order such a ray consistently, so both step functions report it
// (return false, nothing committed) and the kernel walks the REST OF THAT RAY with the plane
// records of walk_core.cuh, whose ties are resolved by bit-identical quotients on both sides of
// a face.  Generic rays never take that path; it costs nothing but the presence of the table.
#pragma once
#include "tet_mesh.hpp"
#include "walk_core.cuh"

namespace ptb {

// Registers of the edge-function walk on top of Ray.  Ray::entry holds the slots of the four
// roles in the current tet, 2 bits each (roles 0,1,2 = entry triple a,b,c; role 3 = d), or -1
// on the first tet of a ray.
struct EdgeRay {
  double ax, ay, az, bx, by, bz, cx, cy, cz;  // entry-face vertices minus the ray origin, ordered
                                              // so that their three edge functions are > 0
  int32_t dv;                                 // vertex id of role 3 in the current tet
};

// |edge function| <= kEdgeTol * (sum of the magnitudes of its terms) counts as zero.  The terms of q = u x P
// carry their own rounding, hence the generous factor; a generic ray is 1e-13 away from an edge with
// probability ~1e-13 per crossing, and then merely finishes on the plane records.
constexpr double kEdgeTol = 1024.0 * 2.220446049250313e-16;

PTB_HD uint32_t sel4(int k, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return k == 0 ? a : (k == 1 ? b : (k == 2 ? c : d));
}

// Leave the current tet through face k; s0,s1,s2 = slots of the three vertices that stay.
// Returns the next tet (-1 = hull), its role->slot word and the vertex to fetch for role 3.
PTB_HD void cross_face(const TetLinks &L, int k, int s0, int s1, int s2, int32_t &next, int32_t &roles,
                       int32_t &dv) {
  const uint32_t a = sel4(k, L.nbr[0], L.nbr[1], L.nbr[2], L.nbr[3]);
  const uint32_t b = sel4(k, L.opp[0], L.opp[1], L.opp[2], L.opp[3]);
  const uint32_t map6 = (a >> 30) | ((b >> 28) << 2);
  const int n0 = (int)((map6 >> (2 * (s0 - (s0 > k)))) & 3u);
  const int n1 = (int)((map6 >> (2 * (s1 - (s1 > k)))) & 3u);
  const int n2 = (int)((map6 >> (2 * (s2 - (s2 > k)))) & 3u);
  roles = n0 | (n1 << 2) | (n2 << 4) | ((6 - n0 - n1 - n2) << 6);
  next = (int32_t)(((a & kIdMask) + 1u) & kIdMask) - 1;
  dv = (int32_t)(b & kVertMask);
}

// Crossing parameter of the exit triple now held in g: with n = (b-a) x (c-a), t = n.a / n.u
// (n.u = the sum of the triple's edge functions > 0; n.a = a . (b x c)).  Only the signs of the
// edge functions steer the walk, so t needs no agreement between neighbouring tets.  Returns
// false when rounding left n.u <= 0 (face parallel to the ray within rounding).
PTB_HD bool edge_exit_parameter(const Ray &r, const EdgeRay &g, double &texit) {
  const double e1x = g.bx - g.ax, e1y = g.by - g.ay, e1z = g.bz - g.az;
  const double e2x = g.cx - g.ax, e2y = g.cy - g.ay, e2z = g.cz - g.az;
  const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
  const double den = nx * r.ux + ny * r.uy + nz * r.uz;
  const double num = nx * g.ax + ny * g.ay + nz * g.az;
  texit = (num < den) ? num / den : __builtin_huge_val();
  return den > 0.0;
}

// Tet after the first: vertex d = (dx,dy,dz) is the only new one.
PTB_HD bool edge_step(const Ray &r, EdgeRay &g, const TetLinks &L, double dx, double dy, double dz,
                      double &texit, int32_t &next, int32_t &roles) {
  const double px = dx - r.ox, py = dy - r.oy, pz = dz - r.oz;
  const double qx = r.uy * pz - r.uz * py, qy = r.uz * px - r.ux * pz, qz = r.ux * py - r.uy * px;
  const double sa = qx * g.ax + qy * g.ay + qz * g.az;
  const double sb = qx * g.bx + qy * g.by + qz * g.bz;
  const double sc = qx * g.cx + qy * g.cy + qz * g.cz;
  // role whose vertex is left behind (the exit face is the one opposite it).  An edge function that is
  // zero within the rounding of its own terms (the compiler may or may not fuse the multiply-adds, so
  // "exactly zero" is not a portable test), or one of the two sign patterns that cannot occur
  // geometrically (+++ / ---), hands the ray to the planes.
  const double aqx = fabs(qx), aqy = fabs(qy), aqz = fabs(qz);
  const double ta = kEdgeTol * (aqx * fabs(g.ax) + aqy * fabs(g.ay) + aqz * fabs(g.az));
  const double tb = kEdgeTol * (aqx * fabs(g.bx) + aqy * fabs(g.by) + aqz * fabs(g.bz));
  const double tc = kEdgeTol * (aqx * fabs(g.cx) + aqy * fabs(g.cy) + aqz * fabs(g.cz));
  const bool pa = sa > 0.0, pb = sb > 0.0, pc = sc > 0.0;
  if (!(fabs(sa) > ta) || !(fabs(sb) > tb) || !(fabs(sc) > tc) || (pa == pb && pb == pc)) return false;
  const int z = pa ? (pb ? 0 : 2) : (pc ? 1 : 0);
  const int e = r.entry;
  int s0 = e & 3, s1 = (e >> 2) & 3, s2 = (e >> 4) & 3;
  const int s3 = (e >> 6) & 3;
  int k;
  EdgeRay h = g;
  if (z == 2) {         // (a,b,d): edge functions m(a,b) inherited, m(b,d) = -sb, m(d,a) = sa
    k = s2; s2 = s3;
    h.cx = px; h.cy = py; h.cz = pz;
  } else if (z == 0) {  // (d,b,c): m(d,b) = sb, m(b,c) inherited, m(c,d) = -sc
    k = s0; s0 = s3;
    h.ax = px; h.ay = py; h.az = pz;
  } else {              // (a,d,c): m(a,d) = -sa, m(d,c) = sc, m(c,a) inherited
    k = s1; s1 = s3;
    h.bx = px; h.by = py; h.bz = pz;
  }
  if (!edge_exit_parameter(r, h, texit)) return false;
  g = h;
  cross_face(L, k, s0, s1, s2, next, roles, g.dv);
  return true;
}

// First tet of a ray: v = the four vertices in slot order (positively oriented).
PTB_HD bool edge_first(const Ray &r, EdgeRay &g, const TetLinks &L, const double (&v)[12], double &texit,
                       int32_t &next, int32_t &roles) {
  double p[4][3], q[4][3];
PTB_UNROLL
  for (int i = 0; i < 4; ++i) {
    p[i][0] = v[3 * i] - r.ox; p[i][1] = v[3 * i + 1] - r.oy; p[i][2] = v[3 * i + 2] - r.oz;
    q[i][0] = r.uy * p[i][2] - r.uz * p[i][1];
    q[i][1] = r.uz * p[i][0] - r.ux * p[i][2];
    q[i][2] = r.ux * p[i][1] - r.uy * p[i][0];
  }
#define PTB_M(i, j) (q[i][0] * p[j][0] + q[i][1] * p[j][1] + q[i][2] * p[j][2])
  const double m01 = PTB_M(0, 1), m02 = PTB_M(0, 2), m03 = PTB_M(0, 3), m12 = PTB_M(1, 2), m13 = PTB_M(1, 3),
               m23 = PTB_M(2, 3);
#undef PTB_M
  // outward faces of a positively oriented tet: 0:(1,2,3) 1:(0,3,2) 2:(0,1,3) 3:(0,2,1); the exit
  // is the one whose edge functions are all > 0 (the entry face has them all < 0).  No such face:
  // a zero (coplanar edge, u = 0) or a start point outside the tet by rounding -> planes.
  int ia, ib, ic, f;
  if (m12 > 0.0 && m23 > 0.0 && m13 < 0.0)      { f = 0; ia = 1; ib = 2; ic = 3; }
  else if (m03 > 0.0 && m23 < 0.0 && m02 < 0.0) { f = 1; ia = 0; ib = 3; ic = 2; }
  else if (m01 > 0.0 && m13 > 0.0 && m03 < 0.0) { f = 2; ia = 0; ib = 1; ic = 3; }
  else if (m02 > 0.0 && m12 < 0.0 && m01 < 0.0) { f = 3; ia = 0; ib = 2; ic = 1; }
  else return false;
  {  // any edge function within rounding of zero: planes (see edge_step)
    double t[4];
PTB_UNROLL
    for (int i = 0; i < 4; ++i) t[i] = fabs(q[i][0]) + fabs(q[i][1]) + fabs(q[i][2]);
    double pm[4];
PTB_UNROLL
    for (int i = 0; i < 4; ++i) pm[i] = fmax(fabs(p[i][0]), fmax(fabs(p[i][1]), fabs(p[i][2])));
    if (!(fabs(m01) > kEdgeTol * t[0] * pm[1]) || !(fabs(m02) > kEdgeTol * t[0] * pm[2]) ||
        !(fabs(m03) > kEdgeTol * t[0] * pm[3]) || !(fabs(m12) > kEdgeTol * t[1] * pm[2]) ||
        !(fabs(m13) > kEdgeTol * t[1] * pm[3]) || !(fabs(m23) > kEdgeTol * t[2] * pm[3]))
      return false;
  }
  auto pick = [&](int i, int d) { return i == 0 ? p[0][d] : (i == 1 ? p[1][d] : (i == 2 ? p[2][d] : p[3][d])); };
  g.ax = pick(ia, 0); g.ay = pick(ia, 1); g.az = pick(ia, 2);
  g.bx = pick(ib, 0); g.by = pick(ib, 1); g.bz = pick(ib, 2);
  g.cx = pick(ic, 0); g.cy = pick(ic, 1); g.cz = pick(ic, 2);
  if (!edge_exit_parameter(r, g, texit)) return false;
  cross_face(L, f, ia, ib, ic, next, roles, g.dv);
  return true;
}

}  // namespace ptb
