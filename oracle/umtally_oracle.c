/*
 * umtally_oracle.c -- CPU restatement of the reference track-length tally path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * engine; nothing in the product (pumiumtally_b200/) may call, link or import
 * it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it.
 *
 * What it restates (all citations are into /root/reference):
 *   - two-phase move (relocate at weight 0, then fly and tally):
 *         src/pumitally/PumiTallyImpl.cpp:66-149
 *   - localisation = the same walk with tallying off, from centroid of tet 0:
 *         PumiTallyImpl.cpp:195-221, 492-528
 *   - per-iteration callback order EvaluateFlux -> prev_x -> VacuumBC ->
 *     UpdateCurrentElement:  PumiTallyImpl.cpp:297-316
 *   - tally  flux[elem] += |x - prev_x| * w  for in_flight==1 && !done:
 *         PumiTallyImpl.cpp:352-380
 *   - vacuum boundary: clip dest to hull point, particle stays in last tet:
 *         PumiTallyImpl.cpp:256-286, 243-254
 *   - prev_x seeding + search driver: PumiTallyImpl.cpp:433-459
 *   - flying[] reset to 0 on return: PumiTallyImpl.cpp:169-172
 *   - volume normalisation: PumiTallyImpl.cpp:382-409
 *
 * PARITY PINNING.  The walk arithmetic itself (ParticleTracer::search,
 * find-exit-face) lives in the un-vendored dependency Fuad-HH/pumi-pic, branch
 * `make_search_class` (pinned by branch name only:
 * .github/actions/install-deps/action.yml:122-136) and is NOT under
 * /root/reference, so the reference path cannot be compiled here.  This
 * restatement follows the tracer's contract as the reference uses it
 * (PumiTallyImpl.h:74-85; last_exit==-1 <=> reached destination, next_elems==-1
 * <=> boundary face) with a standard fp64 segment/tet-face traversal.  It is
 * pinned against every known answer the reference's own test holds for this
 * path (test/test_pumi_tally_impl_methods.cpp:83,152-169,221-282,354-388) in
 * tests/test_oracle_golden.py, and cross-checked against the independent
 * brute-force integrator at the bottom of this file (no adjacency, no walk).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct um_oracle {
  int nverts, ntets, nptcls;
  double *coords; /* [V*3] */
  int *t2v;       /* [E*4] */
  int *t2t;       /* [E*4] neighbour across the face opposite local vertex f; -1 = hull */
  /* particle structure members (PumiTallyImpl.h:39-41): origin, dest, in_flight, weight */
  double *orig, *dest, *wgt;
  short *in_flight;
  /* tracer arrays (callback signature PumiTallyImpl.h:74-85) */
  int *elem_ids, *next_elems, *inter_faces, *last_exit, *ptcl_done, *entry_face;
  double *inter_points, *prev_x;
  double *flux;
  /* staging buffers (PumiTallyImpl.cpp:36-41) */
  double *pos_buf, *wgt_buf;
  signed char *fly_buf;
  int initialized;
  long iter_count;
  int looplimit;
  int per_particle; /* 0 = reference-shaped global loop, 1 = walk each particle to completion */
  int strict_exit;  /* 1 = every face with n.u > 0 is an exit candidate (the rule before the parallel-face
                     * tolerance was added); 0 = default, faces parallel to the segment are skipped */
  /* statistics (not in the reference; used for the segments/s metric) */
  long long n_segments;  /* tally contributions with in_flight==1 during weighted phases */
  long long n_crossings; /* all walk iterations of all particles, any phase */
  long long n_tracks;    /* flying particles in weighted phases */
  long long n_lost;
} um_oracle;

/* ------------------------------------------------------------------ geometry */

static inline void sub3(const double *a, const double *b, double *o) {
  o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2];
}
static inline double dot3(const double *a, const double *b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* Outward plane of face f (opposite local vertex f) of tet e: n.x = c on the
 * face, n.x < c inside.  The three face vertices are taken in ascending global
 * id so the two tets sharing a face evaluate the identical expression. */
static void face_plane(const um_oracle *o, int e, int f, double *n, double *c) {
  int v[3], k = 0;
  for (int i = 0; i < 4; ++i)
    if (i != f) v[k++] = o->t2v[4 * e + i];
  if (v[0] > v[1]) { int t = v[0]; v[0] = v[1]; v[1] = t; }
  if (v[1] > v[2]) { int t = v[1]; v[1] = v[2]; v[2] = t; }
  if (v[0] > v[1]) { int t = v[0]; v[0] = v[1]; v[1] = t; }
  const double *A = o->coords + 3 * v[0], *B = o->coords + 3 * v[1], *C = o->coords + 3 * v[2];
  double ab[3], ac[3];
  sub3(B, A, ab);
  sub3(C, A, ac);
  cross3(ab, ac, n);
  *c = dot3(n, A);
  const double *P = o->coords + 3 * o->t2v[4 * e + f];
  if (dot3(n, P) - *c > 0.0) { /* opposite vertex must be on the negative side */
    n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2];
    *c = -*c;
  }
}

/* The tracer's per-iteration "find exit face" step (reference: the external
 * ParticleTracer::search called at PumiTallyImpl.cpp:454; contract in
 * PumiTallyImpl.h:74-85).  Segment O->D is the particle's (origin, dest) pair,
 * fixed for the whole search.  Returns local exit face or -1 when D lies in
 * tet e; xpt receives the intersection point (or D). */
static int find_exit_face(const um_oracle *o, int e, int entry, const double *O, const double *D,
                          double *xpt) {
  double u[3];
  sub3(D, O, u);
  double tbest = INFINITY;
  int fbest = -1;
  for (int f = 0; f < 4; ++f) {
    if (f == entry) continue; /* never leave through the face just entered */
    double n[3], c;
    face_plane(o, e, f, n, &c);
    double den = dot3(n, u);
    /* a face the segment is parallel to (within 1e-12 rad) is no exit candidate: for a track that runs
     * inside a face plane or along an edge, den is rounding noise and num/den an arbitrary number.  The
     * reference tracer carries a tolerance for the same purpose (constructor argument 1e-8,
     * PumiTallyImpl.cpp:51); generic tracks are unaffected.  Same rule as scan_face() in the CUDA path --
     * which makes the behaviour on such degenerate tracks self-referential: it was chosen, not derived
     * from the (absent) tracer source.  The rule it replaced (den > 0, um_oracle_set_exit_rule(o, 1)) is
     * kept so that tests can assert both give identical results on every generic workload. */
    if (o->strict_exit ? den > 0.0 : den > 1e-12 * (fabs(u[0]) + fabs(u[1]) + fabs(u[2]))) {
      double t = (c - dot3(n, O)) / den;
      if (t < tbest) { tbest = t; fbest = f; }
    }
  }
  if (fbest < 0 || tbest >= 1.0) {
    xpt[0] = D[0]; xpt[1] = D[1]; xpt[2] = D[2];
    return -1;
  }
  if (tbest < 0.0) tbest = 0.0;
  xpt[0] = O[0] + tbest * u[0];
  xpt[1] = O[1] + tbest * u[1];
  xpt[2] = O[2] + tbest * u[2];
  return fbest;
}

/* ---------------------------------------------------------------- adjacency */

typedef struct { int a, b, c, tet, face; } face_key;
static int face_cmp(const void *x, const void *y) {
  const face_key *p = (const face_key *)x, *q = (const face_key *)y;
  if (p->a != q->a) return p->a < q->a ? -1 : 1;
  if (p->b != q->b) return p->b < q->b ? -1 : 1;
  if (p->c != q->c) return p->c < q->c ? -1 : 1;
  return 0;
}

static int build_adjacency(um_oracle *o) {
  size_t nf = (size_t)o->ntets * 4;
  face_key *keys = (face_key *)malloc(nf * sizeof(face_key));
  if (!keys) return -1;
  for (int e = 0; e < o->ntets; ++e)
    for (int f = 0; f < 4; ++f) {
      int v[3], k = 0;
      for (int i = 0; i < 4; ++i)
        if (i != f) v[k++] = o->t2v[4 * e + i];
      if (v[0] > v[1]) { int t = v[0]; v[0] = v[1]; v[1] = t; }
      if (v[1] > v[2]) { int t = v[1]; v[1] = v[2]; v[2] = t; }
      if (v[0] > v[1]) { int t = v[0]; v[0] = v[1]; v[1] = t; }
      face_key *q = &keys[4 * (size_t)e + f];
      q->a = v[0]; q->b = v[1]; q->c = v[2]; q->tet = e; q->face = f;
    }
  qsort(keys, nf, sizeof(face_key), face_cmp);
  for (size_t i = 0; i < nf; ++i) o->t2t[i] = -1;
  int bad = 0;
  for (size_t i = 0; i < nf;) {
    size_t j = i + 1;
    while (j < nf && face_cmp(&keys[i], &keys[j]) == 0) ++j;
    if (j - i == 2) {
      o->t2t[4 * keys[i].tet + keys[i].face] = keys[i + 1].tet;
      o->t2t[4 * keys[i + 1].tet + keys[i + 1].face] = keys[i].tet;
    } else if (j - i > 2) {
      bad = 1;
    }
    i = j;
  }
  free(keys);
  return bad;
}

/* ------------------------------------------------------------------ lifecycle */

static void seed_in_element0(um_oracle *o) {
  /* PumiTallyImpl.cpp:492-528: every particle starts at the centroid of tet 0 */
  double c[3] = {0, 0, 0};
  for (int i = 0; i < 4; ++i)
    for (int d = 0; d < 3; ++d) c[d] += o->coords[3 * o->t2v[i] + d];
  for (int d = 0; d < 3; ++d) c[d] /= 4.0; /* Omega_h::average = sum / n */
  for (int p = 0; p < o->nptcls; ++p) {
    for (int d = 0; d < 3; ++d) o->orig[3 * p + d] = c[d];
    o->in_flight[p] = 1;
    o->elem_ids[p] = 0; /* PumiTallyImpl.cpp:472-475: all particles parked in element 0 */
  }
}

um_oracle *um_oracle_create(const double *coords, int nverts, const int *tet2vert, int ntets,
                            int nptcls) {
  um_oracle *o = (um_oracle *)calloc(1, sizeof(um_oracle));
  o->nverts = nverts; o->ntets = ntets; o->nptcls = nptcls;
  o->coords = (double *)malloc(sizeof(double) * 3 * (size_t)nverts);
  o->t2v = (int *)malloc(sizeof(int) * 4 * (size_t)ntets);
  o->t2t = (int *)malloc(sizeof(int) * 4 * (size_t)ntets);
  memcpy(o->coords, coords, sizeof(double) * 3 * (size_t)nverts);
  memcpy(o->t2v, tet2vert, sizeof(int) * 4 * (size_t)ntets);
  if (build_adjacency(o)) fprintf(stderr, "[oracle] ERROR: non-manifold face in mesh\n");
  size_t n = (size_t)nptcls;
  o->orig = (double *)calloc(3 * n, sizeof(double));
  o->dest = (double *)calloc(3 * n, sizeof(double));
  o->wgt = (double *)calloc(n, sizeof(double));
  o->in_flight = (short *)calloc(n, sizeof(short));
  o->elem_ids = (int *)calloc(n, sizeof(int));
  o->next_elems = (int *)calloc(n, sizeof(int));
  o->inter_faces = (int *)calloc(n, sizeof(int));
  o->last_exit = (int *)calloc(n, sizeof(int));
  o->ptcl_done = (int *)calloc(n, sizeof(int));
  o->entry_face = (int *)calloc(n, sizeof(int));
  o->inter_points = (double *)calloc(3 * n, sizeof(double));
  o->prev_x = (double *)calloc(3 * n, sizeof(double));
  o->flux = (double *)calloc((size_t)ntets, sizeof(double));
  o->pos_buf = (double *)calloc(3 * n, sizeof(double));
  o->wgt_buf = (double *)calloc(n, sizeof(double));
  o->fly_buf = (signed char *)calloc(n, 1);
  o->looplimit = ntets + 16;
  o->per_particle = 0;
  seed_in_element0(o);
  return o;
}

void um_oracle_destroy(um_oracle *o) {
  if (!o) return;
  free(o->coords); free(o->t2v); free(o->t2t); free(o->orig); free(o->dest); free(o->wgt);
  free(o->in_flight); free(o->elem_ids); free(o->next_elems); free(o->inter_faces);
  free(o->last_exit); free(o->ptcl_done); free(o->entry_face); free(o->inter_points);
  free(o->prev_x); free(o->flux); free(o->pos_buf); free(o->wgt_buf); free(o->fly_buf);
  free(o);
}

void um_oracle_set_mode(um_oracle *o, int per_particle) { o->per_particle = per_particle; }
void um_oracle_set_exit_rule(um_oracle *o, int strict) { o->strict_exit = strict; }

/* -------------------------------------------------- search, reference-shaped */

/* One tracer iteration for particle p followed by the reference functor body
 * for that particle.  Splitting by particle instead of by kernel is legal
 * because every kernel in the sequence only touches slot p of each array. */
static inline void iterate_particle(um_oracle *o, int p, int initial, int weighted,
                                    long long *segs, long long *cross) {
  const double *O = o->orig + 3 * p, *D = o->dest + 3 * p;
  double *x = o->inter_points + 3 * p;
  int e = o->elem_ids[p];
  /* tracer: exit face, intersection point, neighbour */
  int f = find_exit_face(o, e, o->entry_face[p], O, D, x);
  o->last_exit[p] = f;
  int nxt = (f >= 0) ? o->t2t[4 * e + f] : e;
  o->next_elems[p] = nxt;
  ++*cross;
  /* functor (PumiTallyImpl.cpp:297-316) */
  if (!initial) {
    /* EvaluateFlux (PumiTallyImpl.cpp:352-380) */
    if (o->in_flight[p] == 1 && !o->ptcl_done[p]) {
      double *px = o->prev_x + 3 * p;
      double d[3];
      sub3(x, px, d);
      double len = sqrt(dot3(d, d)); /* Omega_h::norm */
      double contribution = len * o->wgt[p];
#ifdef _OPENMP
#pragma omp atomic
#endif
      o->flux[e] += contribution;
      if (weighted) ++*segs;
    }
    /* UpdatePreviousXPoints (PumiTallyImpl.cpp:322-333) */
    o->prev_x[3 * p] = x[0]; o->prev_x[3 * p + 1] = x[1]; o->prev_x[3 * p + 2] = x[2];
  }
  /* ApplyVacuumBC (PumiTallyImpl.cpp:256-286) */
  if (!o->ptcl_done[p]) {
    int reached = (f == -1);
    int hit = (nxt == -1) && (e != -1);
    if (reached || hit) o->ptcl_done[p] = 1;
    if (hit) {
      o->inter_faces[p] = f;
      o->dest[3 * p] = x[0]; o->dest[3 * p + 1] = x[1]; o->dest[3 * p + 2] = x[2];
    }
  }
  /* UpdateCurrentElement (PumiTallyImpl.cpp:243-254) */
  if (o->in_flight[p] && nxt != -1) {
    if (nxt != e) {
      /* remember which local face of the new tet we came through */
      int ef = -1;
      for (int k = 0; k < 4; ++k)
        if (o->t2t[4 * nxt + k] == e) ef = k;
      o->entry_face[p] = ef;
    }
    o->elem_ids[p] = nxt;
  }
}

/* SearchAndRebuild (PumiTallyImpl.cpp:433-459).  `weighted` only steers the
 * segment statistics. */
static int search(um_oracle *o, int initial, int weighted) {
  int n = o->nptcls;
  if (!initial) /* UpdatePreviousXPoints(ptcls): prev_x <- origin (PumiTallyImpl.cpp:335-350) */
    memcpy(o->prev_x, o->orig, sizeof(double) * 3 * (size_t)n);
  for (int p = 0; p < n; ++p) { o->ptcl_done[p] = 0; o->entry_face[p] = -1; }
  long long segs = 0, cross = 0;
  int found_all = 1;
  if (o->per_particle) {
    long long lost = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : segs, cross, lost)
#endif
    for (int p = 0; p < n; ++p) {
      int loops = 0;
      while (!o->ptcl_done[p]) {
        iterate_particle(o, p, initial, weighted, &segs, &cross);
        if (++loops > o->looplimit) { ++lost; break; }
      }
    }
    if (lost) { found_all = 0; o->n_lost += lost; }
  } else {
    int loops = 0;
    for (;;) {
      long long remaining = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : segs, cross, remaining)
#endif
      for (int p = 0; p < n; ++p) {
        if (o->ptcl_done[p]) continue; /* tracer kernels skip done particles */
        iterate_particle(o, p, initial, weighted, &segs, &cross);
        if (!o->ptcl_done[p]) ++remaining;
      }
      if (!remaining) break;
      if (++loops > o->looplimit) { found_all = 0; o->n_lost += remaining; break; }
    }
  }
  o->n_segments += segs;
  o->n_crossings += cross;
  /* tracer commit: origin <- dest (semantics documented by the dead
   * CommitParticlePositions, PumiTallyImpl.cpp:418-431; pinned by
   * test_pumi_tally_impl_methods.cpp:243-251, 323-346) */
  memcpy(o->orig, o->dest, sizeof(double) * 3 * (size_t)n);
  if (!found_all)
    printf("ERROR: Not all particles are found. May need more loops in search\n");
  return found_all;
}

/* ---------------------------------------------------------------- public API */

/* CopyInitialPositionToBuffer + MoveToInitialLocation
 * (PumiTallyImpl.cpp:54-64, 195-221).  size = 3 * num_particles. */
void um_oracle_copy_initial_position(um_oracle *o, const double *xyz, int size) {
  if (size != 3 * o->nptcls) { fprintf(stderr, "[oracle] size != 3N\n"); return; }
  if (o->initialized) { fprintf(stderr, "[oracle] CopyInitialPosition called twice\n"); return; }
  memcpy(o->pos_buf, xyz, sizeof(double) * (size_t)size);
  for (int p = 0; p < o->nptcls; ++p) {
    for (int d = 0; d < 3; ++d) o->dest[3 * p + d] = o->pos_buf[3 * p + d];
    o->in_flight[p] = 1;
  }
  search(o, 1, 0);
  o->initialized = 1;
}

/* MoveToNextLocation (PumiTallyImpl.cpp:66-149).  size = 3 * num_particles. */
void um_oracle_move_to_next_location(um_oracle *o, const double *origin, const double *dest,
                                     signed char *flying, const double *weights, int size) {
  int n = o->nptcls;
  if (size != 3 * n) { fprintf(stderr, "[oracle] size != 3N\n"); return; }
  if (!o->initialized) { fprintf(stderr, "[oracle] move before CopyInitialPosition\n"); return; }
  /* phase 1: relocate to caller's origin with weight 0 (lines 73-112) */
  memcpy(o->pos_buf, origin, sizeof(double) * 3 * (size_t)n);
  memcpy(o->fly_buf, flying, (size_t)n);
  for (int p = 0; p < n; ++p) flying[p] = 0; /* lines 169-172 */
  for (int p = 0; p < n; ++p) {
    o->in_flight[p] = (short)(unsigned char)o->fly_buf[p];
    const double *src = (o->in_flight[p] == 1) ? o->pos_buf + 3 * p : o->orig + 3 * p;
    for (int d = 0; d < 3; ++d) o->dest[3 * p + d] = src[d];
    o->wgt[p] = 0.0;
  }
  search(o, 0, 0);
  /* phase 2: fly to destination and tally with weight (lines 119-145) */
  memcpy(o->pos_buf, dest, sizeof(double) * 3 * (size_t)n);
  memcpy(o->wgt_buf, weights, sizeof(double) * (size_t)n);
  for (int p = 0; p < n; ++p) o->wgt[p] = o->wgt_buf[p];
  for (int p = 0; p < n; ++p) {
    const double *src = (o->in_flight[p] == 1) ? o->pos_buf + 3 * p : o->orig + 3 * p;
    for (int d = 0; d < 3; ++d) o->dest[3 * p + d] = src[d];
    if (o->in_flight[p] == 1) o->n_tracks++;
  }
  o->iter_count++;
  search(o, 0, 1);
}

/* NormalizeFlux (PumiTallyImpl.cpp:382-409): flux / tet volume. */
void um_oracle_normalized_flux(const um_oracle *o, double *out_flux, double *out_volume) {
  for (int e = 0; e < o->ntets; ++e) {
    const double *a = o->coords + 3 * o->t2v[4 * e], *b = o->coords + 3 * o->t2v[4 * e + 1];
    const double *c = o->coords + 3 * o->t2v[4 * e + 2], *d = o->coords + 3 * o->t2v[4 * e + 3];
    double ab[3], ac[3], ad[3], cr[3];
    sub3(b, a, ab); sub3(c, a, ac); sub3(d, a, ad);
    cross3(ab, ac, cr);
    double vol = fabs(dot3(cr, ad)) / 6.0;
    if (out_volume) out_volume[e] = vol;
    if (out_flux) out_flux[e] = o->flux[e] / vol;
  }
}

/* Places particles [0, count) at given positions in given tets without walking there (test support
 * for drivers that hand particles between picparts; not part of the reference interface). */
void um_oracle_set_state(um_oracle *o, const double *xyz, const int *elems, int count) {
  for (int p = 0; p < count && p < o->nptcls; ++p) {
    for (int d = 0; d < 3; ++d) o->orig[3 * p + d] = xyz[3 * p + d];
    o->elem_ids[p] = elems[p];
  }
  o->initialized = 1;
}

int um_oracle_ntets(const um_oracle *o) { return o->ntets; }
int um_oracle_nptcls(const um_oracle *o) { return o->nptcls; }
const double *um_oracle_flux(const um_oracle *o) { return o->flux; }
const int *um_oracle_elem_ids(const um_oracle *o) { return o->elem_ids; }
const double *um_oracle_positions(const um_oracle *o) { return o->orig; }
const int *um_oracle_adjacency(const um_oracle *o) { return o->t2t; }
long long um_oracle_n_segments(const um_oracle *o) { return o->n_segments; }
long long um_oracle_n_crossings(const um_oracle *o) { return o->n_crossings; }
long long um_oracle_n_tracks(const um_oracle *o) { return o->n_tracks; }
long long um_oracle_n_lost(const um_oracle *o) { return o->n_lost; }
void um_oracle_reset_flux(um_oracle *o) {
  memset(o->flux, 0, sizeof(double) * (size_t)o->ntets);
  o->n_segments = o->n_crossings = o->n_tracks = 0;
}
void um_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int um_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------- brute-force cross-check
 * Independent of adjacency and of the walk: clip each weighted segment
 * [a,b] against every tet (four half-spaces from the vertex coordinates,
 * oriented by the tet's own signed volume) and add the clipped length * w.
 * O(N*E) -- small cases only.  Also returns, per segment, the largest
 * parameter t still inside the mesh (the vacuum-BC clip point for a convex
 * mesh) and the tet that contains that end point's approach. */
void um_bruteforce_tally(const double *coords, const int *t2v, int ntets, const double *a,
                         const double *b, const double *w, int nseg, double *flux,
                         double *t_last, int *elem_last) {
  for (int s = 0; s < nseg; ++s) {
    const double *A = a + 3 * s, *B = b + 3 * s;
    double u[3];
    sub3(B, A, u);
    double L = sqrt(dot3(u, u));
    double tl = 0.0;
    int el = -1;
    for (int e = 0; e < ntets; ++e) {
      const double *V[4];
      for (int i = 0; i < 4; ++i) V[i] = coords + 3 * t2v[4 * e + i];
      double lo = 0.0, hi = 1.0;
      for (int f = 0; f < 4 && lo <= hi; ++f) {
        const double *P = V[(f + 1) & 3], *Q = V[(f + 2) & 3], *R = V[(f + 3) & 3];
        double pq[3], pr[3], n[3];
        sub3(Q, P, pq); sub3(R, P, pr);
        cross3(pq, pr, n);
        double pin[3];
        sub3(V[f], P, pin);
        if (dot3(n, pin) > 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
        double pa[3];
        sub3(A, P, pa);
        double g0 = dot3(n, pa), g1 = dot3(n, u); /* g(t) = g0 + t*g1 <= 0 inside */
        if (g1 > 0) { double t = -g0 / g1; if (t < hi) hi = t; }
        else if (g1 < 0) { double t = -g0 / g1; if (t > lo) lo = t; }
        else if (g0 > 0) { hi = -1.0; }
      }
      if (hi > lo) {
        flux[e] += (hi - lo) * L * w[s];
        if (hi > tl) { tl = hi; el = e; }
      }
    }
    if (t_last) t_last[s] = tl;
    if (elem_last) elem_last[s] = el;
  }
}
