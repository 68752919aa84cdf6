// See homme_driver.hpp. Host-side stand-in for HOMME's Fortran driver.
#include "homme_driver.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace {

constexpr int NP = 4;
constexpr int NPSQ = 16;
constexpr double DD_PI = 3.141592653589793238462643383279;
// src/share/physical_constants.F90:58-70
constexpr double REARTH = 6.376e6;
constexpr double OMEGA = 7.292e-5;
constexpr double GRAV = 9.80616;
constexpr double RGAS = 287.04;
constexpr double P0 = 100000.0;
constexpr double DIST_THRESHOLD = 1.0e-9;  // src/share/coordinate_systems_mod.F90

std::string g_last_error;

// C position codes used by the dycore (mpi_cxx_f90_interface.cpp:47-49): S,N,W,E,SW,SE,NW,NE.
enum { POS_S = 0, POS_N, POS_W, POS_E, POS_SW, POS_SE, POS_NW, POS_NE };
// Fortran codes (src/share/control_mod.F90:190-197): W=1,E=2,S=3,N=4,SW=5,SE=6,NW=7,NE=8.
int c_pos_to_f90(int cpos) {
  static const int t[8] = {3, 4, 1, 2, 5, 6, 7, 8};
  return t[cpos];
}

// ---- GLL quadrature and derivative matrix (np = 4) ----------------------------------------
struct Gll {
  long double x[NP], w[NP];
  double dvv[NP][NP];  // memory image of F90 deriv%Dvv: dvv[l][i] = dl_i/dx (x_l)
};

long double legendre(int n, long double x) {
  long double p0 = 1, p1 = x;
  if (n == 0) return p0;
  for (int k = 2; k <= n; ++k) {
    long double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1; p1 = p2;
  }
  return p1;
}

Gll make_gll() {
  Gll g;
  // Gauss-Lobatto points for np=4: +-1, +-sqrt(1/5); weights 2/(n(n-1) P_{n-1}(x)^2)
  g.x[0] = -1.0L; g.x[1] = -sqrtl(0.2L); g.x[2] = sqrtl(0.2L); g.x[3] = 1.0L;
  for (int i = 0; i < NP; ++i) {
    long double p = legendre(NP - 1, g.x[i]);
    g.w[i] = 2.0L / (NP * (NP - 1) * p * p);
  }
  // derivative_mod_base.F90:440-475 : Dvv_F(j,i) = L(x_i)/L(x_j)/(x_i-x_j), corners -+np(np-1)/4.
  // Memory image (column-major F90 (j,i) == row-major [i][j]) is the standard collocation
  // matrix d l_j/dx at x_i.
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) {
      long double d;
      if (i != j)
        d = legendre(NP - 1, g.x[i]) / legendre(NP - 1, g.x[j]) / (g.x[i] - g.x[j]);
      else if (i == 0)
        d = -(long double)(NP * (NP - 1)) / 4.0L;
      else if (i == NP - 1)
        d = (long double)(NP * (NP - 1)) / 4.0L;
      else
        d = 0.0L;
      g.dvv[i][j] = (double)d;
    }
  return g;
}

// ---- HOMME's space-filling curve over an n x n face ------------------------------------------
// spacecurve_mod.F90:39-1040. n = 2^a 3^b 5^c is traversed by nested Hilbert (2 x 2), meandering Peano (3 x 3)
// and Cinco (5 x 5) refinements, the factors of 2 finest and the factors of 5 coarsest (Factor :901-974,
// map :994-1009). Each refinement visits its sub-cells in a fixed order; a sub-cell is itself a curve
// described by a major axis / direction and a "joiner" (the unit step that leaves it). The reference spells
// the 4 + 9 + 25 sub-cells out as code; here they are rows of one table relative to the parent's frame:
//   {A, D, JA, JD}: major axis = (ma + A) mod 2, major direction = D * md,
//                   joiner = JA == 2 ? the parent's own joiner : axis (ma + JA) mod 2, direction JD * md.
struct SubCell { signed char A, D, JA, JD; };
const SubCell HILBERT[4] = {{1, 1, 1, 1}, {0, 1, 0, 1}, {0, 1, 1, -1}, {1, -1, 2, 0}};
const SubCell PEANO[9] = {{1, 1, 1, 1}, {1, 1, 1, 1}, {0, 1, 0, 1}, {0, 1, 0, 1}, {0, 1, 1, -1},
                          {0, -1, 0, -1}, {1, -1, 1, -1}, {1, -1, 0, 1}, {0, 1, 2, 0}};
const SubCell CINCO[25] = {{0, 1, 0, 1}, {0, 1, 0, 1}, {1, 1, 1, 1}, {1, 1, 1, 1}, {1, 1, 0, -1},
                           {1, -1, 1, -1}, {0, -1, 0, -1}, {0, -1, 1, 1}, {1, 1, 1, 1}, {1, 1, 1, 1},
                           {0, 1, 0, 1}, {0, 1, 1, -1}, {1, -1, 0, 1}, {1, 1, 1, 1}, {0, 1, 0, 1},
                           {0, 1, 0, 1}, {0, 1, 1, -1}, {0, -1, 0, -1}, {1, -1, 1, -1}, {1, -1, 0, 1},
                           {0, 1, 1, -1}, {0, -1, 0, -1}, {1, -1, 1, -1}, {1, -1, 0, 1}, {0, 1, 2, 0}};

struct FaceCurve {
  int n = 0;
  std::vector<int> factors;   // finest first
  std::vector<int> order;     // order[i + n * j] = visit number of cell (i, j) ("Mesh(i+1, j+1)")
  int pos[2] = {0, 0}, count = 0;
  void gen(int level, int ma, int md, int ja, int jd) {   // GenCurve :885-899
    const int type = factors[level - 1];
    const SubCell* t = type == 2 ? HILBERT : type == 3 ? PEANO : CINCO;
    for (int c = 0; c < type * type; ++c) {
      const int lma = (ma + t[c].A) & 1, lmd = t[c].D * md;
      const int lja = t[c].JA == 2 ? ja : (ma + t[c].JA) & 1, ljd = t[c].JA == 2 ? jd : t[c].JD * md;
      if (level > 1) gen(level - 1, lma, lmd, lja, ljd);
      else {  // IncrementCurve :771-783
        order[pos[0] + n * pos[1]] = count++;
        pos[lja] += ljd;
      }
    }
  }
};
bool factor_235(int n, std::vector<int>& f) {
  f.clear();
  for (int p : {2, 3, 5})
    while (n % p == 0) { f.push_back(p); n /= p; }
  return n == 1 && !f.empty();
}
// Mesh(ne, ne) of CubeTopology (cube_mod.F90:1457-1523): the curve itself when ne factors into 2, 3, 5, else the
// curve of the next power of two sampled at the ne x ne cell centres.
std::vector<int> face_curve(int ne) {
  FaceCurve c;
  if (ne == 1) return {0};
  if (factor_235(ne, c.factors)) {
    c.n = ne;
    c.order.assign((size_t)ne * ne, 0);
    c.gen((int)c.factors.size(), 0, 1, 0, 1);
    return c.order;
  }
  int ne2 = 1;
  while (ne2 < ne) ne2 *= 2;
  factor_235(ne2, c.factors);
  c.n = ne2;
  c.order.assign((size_t)ne2 * ne2, 0);
  c.gen((int)c.factors.size(), 0, 1, 0, 1);
  std::vector<int> to_small((size_t)ne2 * ne2, -1);   // Mesh2_map
  for (int j = 1; j <= ne; ++j)
    for (int i = 1; i <= ne; ++i) {
      int i2 = (int)std::lround(((i - 0.5) / ne) * ne2 + 0.5), j2 = (int)std::lround(((j - 0.5) / ne) * ne2 + 0.5);
      i2 = std::min(std::max(i2, 1), ne2);
      j2 = std::min(std::max(j2, 1), ne2);
      to_small[(i2 - 1) + (size_t)ne2 * (j2 - 1)] = (i - 1) + ne * (j - 1);
    }
  std::vector<int> where((size_t)ne2 * ne2);          // sfcij: visit number -> cell of the big mesh
  for (int k = 0; k < ne2 * ne2; ++k) where[c.order[k]] = k;
  std::vector<int> mesh((size_t)ne * ne, 0);
  int idx = 0;
  for (int k = 0; k < ne2 * ne2; ++k)
    if (to_small[where[k]] >= 0) mesh[to_small[where[k]]] = idx++;
  return mesh;
}

struct Vec3i {
  int x, y, z;
  bool operator<(const Vec3i& o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
  bool operator==(const Vec3i& o) const { return x == o.x && y == o.y && z == o.z; }
};

// lattice point (X,Y in [-ne,ne]) on face f (1..6) -> integer point on the cube surface
Vec3i face_to_cube(int f, int X, int Y, int ne) {
  switch (f) {
    case 1: return {ne, X, Y};
    case 2: return {-X, ne, Y};
    case 3: return {-ne, -X, Y};
    case 4: return {X, -ne, Y};
    case 5: return {Y, X, -ne};
    default: return {-Y, X, ne};
  }
}
void face_to_cart(int f, double X, double Y, double c[3]) {
  switch (f) {
    case 1: c[0] = 1; c[1] = X; c[2] = Y; break;
    case 2: c[0] = -X; c[1] = 1; c[2] = Y; break;
    case 3: c[0] = -1; c[1] = -X; c[2] = Y; break;
    case 4: c[0] = X; c[1] = -1; c[2] = Y; break;
    case 5: c[0] = Y; c[1] = X; c[2] = -1; break;
    default: c[0] = -Y; c[1] = X; c[2] = 1; break;
  }
  double r = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  c[0] /= r; c[1] /= r; c[2] /= r;
}

// cube_mod.F90:596-704 (vmap): D maps contravariant cube vectors to (u,v) on the sphere.
void vmap(double D[2][2], double x1, double x2, int face) {
  const double t1 = std::tan(x1), t2 = std::tan(x2), c1 = std::cos(x1), c2 = std::cos(x2);
  const double r = std::sqrt(1.0 + t1 * t1 + t2 * t2);
  if (face <= 4) {
    D[0][0] = 1.0 / (r * c1);
    D[0][1] = 0.0;
    D[1][0] = -t1 * t2 / (c1 * r * r);
    D[1][1] = 1.0 / (r * r * c1 * c2 * c2);
    return;
  }
  const double poledist = std::sqrt(t1 * t1 + t2 * t2);
  if (poledist <= DIST_THRESHOLD) {
    D[0][0] = 1; D[0][1] = 0; D[1][0] = 0; D[1][1] = 1;
    return;
  }
  const double s = (face == 6) ? 1.0 : -1.0;
  D[0][0] = -s * t2 / (poledist * c1 * c1 * r);
  D[0][1] = s * t1 / (poledist * c2 * c2 * r);
  D[1][0] = -s * t1 / (poledist * c1 * c1 * r * r);
  D[1][1] = -s * t2 / (poledist * c2 * c2 * r * r);
}

using fn_v = void (*)();

}  // namespace

struct HommeDriver {
  HommeParams p;
  Gll gll;
  int nelem = 0;   // global
  int nelemd = 0;  // local
  std::vector<double> hyai, hybi, hyam, hybm;

  // global mesh
  std::vector<int> sfc_order;   // position along the curve -> gid
  std::vector<int> owner;       // gid -> rank
  std::vector<int> gid2lid;     // gid -> local id on its owner
  std::vector<std::array<int, 8>> nbr, nbr_pos;  // gid -> neighbour gid / its position, -1 missing
  std::vector<int> local_gids;  // lid -> gid

  // Fortran-layout arrays for the local elements
  std::vector<double> D, Dinv, metinv, tensorvisc, vec_sph2cart;
  std::vector<double> fcor, mp, spheremp, rspheremp, metdet, phis, lat, lon, gidf;
  std::vector<double> v, T, dp3d, Qdp, Q, ps_v, omega_p;
  std::vector<double> accum[7];  // Qvar, Qmass, Q1mass, IEner, IEner_wet, KEner, PEner (elem%accum)
  std::vector<double> FM, FT, FQ;  // elem%derived%FM/FT/FQ (CAM forcing), Fortran layout
  int last_step = 1 << 30;         // nEndStep
  std::vector<int> conn;  // add_connection tuples

  int nstep = 0, nm1 = 1, n0 = 2, np1 = 3;  // Fortran 1-based time levels

  // bound dycore
  void* lib = nullptr;
  std::map<std::string, void*> sym;
};

namespace {

void elem_geometry(const HommeDriver& h, int gid, double alpha, double* Dm /*[2][2][16] math*/,
                   double* lat, double* lon) {
  const int ne = h.p.ne;
  const int face = gid / (ne * ne) + 1;
  const int ei = gid % ne, ej = (gid / ne) % ne;
  const double dx = DD_PI / (2.0 * ne);
  const double sa = std::sqrt(alpha);
  for (int igp = 0; igp < NP; ++igp)
    for (int jgp = 0; jgp < NP; ++jgp) {
      const double a = (double)h.gll.x[jgp], b = (double)h.gll.x[igp];
      const double x1 = -DD_PI / 4 + dx * (ei + 0.5 * (1.0 + a));
      const double x2 = -DD_PI / 4 + dx * (ej + 0.5 * (1.0 + b));
      double tD[2][2];
      vmap(tD, x1, x2, face);
      const int pt = igp * NP + jgp;
      // D = vmap * Jp with Jp = diag(dx/2, dx/2) on the uniform grid (cube_mod.F90:570-593)
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) Dm[(r * 2 + c) * NPSQ + pt] = tD[r][c] * (dx / 2) * sa;
      double cart[3];
      face_to_cart(face, std::tan(x1), std::tan(x2), cart);
      lat[pt] = std::asin(cart[2]);
      double lo = 0.0;
      if (std::fabs(std::fabs(lat[pt]) - DD_PI / 2) >= DIST_THRESHOLD) lo = std::atan2(cart[1], cart[0]);
      if (lo < 0) lo += 2 * DD_PI;
      lon[pt] = lo;
    }
}

// Ordered contributions of the DSS for one scalar defined at every (gid, point): the same
// S,N,W,E-then-corners accumulation the dycore uses (BoundaryExchange.cpp:505-537).
struct EdgePts { int p[4]; };
const EdgePts EDGE_FWD[4] = {{{0, 1, 2, 3}}, {{12, 13, 14, 15}}, {{0, 4, 8, 12}}, {{3, 7, 11, 15}}};
const int CORNER_PT[4] = {0, 3, 12, 15};
// ConnectivityHelpers.hpp:133-142 — true = BACKWARD
const bool DIR_BWD[4][4] = {{true, false, false, true},
                            {false, true, true, false},
                            {false, true, true, false},
                            {true, false, false, true}};

void global_dss(const HommeDriver& h, std::vector<double>& f /*[gid][16][nk]*/, int nk) {
  std::vector<double> out(f);
  for (int g = 0; g < h.nelem; ++g) {
    double* o = &out[(size_t)g * NPSQ * nk];
    for (int k = 0; k < NP; ++k)
      for (int e = 0; e < 4; ++e) {
        const int ng = h.nbr[g][e], npos = h.nbr_pos[g][e];
        const bool bwd = DIR_BWD[npos][e];
        const int rp = EDGE_FWD[npos].p[bwd ? 3 - k : k];
        const double* src = &f[((size_t)ng * NPSQ + rp) * nk];
        double* dst = &o[(size_t)EDGE_FWD[e].p[k] * nk];
        for (int l = 0; l < nk; ++l) dst[l] += src[l];
      }
    for (int c = 0; c < 4; ++c) {
      const int ng = h.nbr[g][4 + c];
      if (ng < 0) continue;
      const int rp = CORNER_PT[h.nbr_pos[g][4 + c] - 4];
      const double* src = &f[((size_t)ng * NPSQ + rp) * nk];
      double* dst = &o[(size_t)CORNER_PT[c] * nk];
      for (int l = 0; l < nk; ++l) dst[l] += src[l];
    }
  }
  f.swap(out);
}

void build_topology(HommeDriver& h) {
  const int ne = h.p.ne;
  h.nelem = 6 * ne * ne;
  std::map<Vec3i, std::vector<std::pair<int, int>>> vert2elem;  // vertex -> (gid, local corner 0..3)
  std::vector<std::array<Vec3i, 4>> verts(h.nelem);
  for (int g = 0; g < h.nelem; ++g) {
    const int face = g / (ne * ne) + 1, ei = g % ne, ej = (g / ne) % ne;
    // local corners in C corner order SW, SE, NW, NE  (x <-> jgp, y <-> igp)
    const int ci[4] = {ei, ei + 1, ei, ei + 1}, cj[4] = {ej, ej, ej + 1, ej + 1};
    for (int c = 0; c < 4; ++c) {
      verts[g][c] = face_to_cube(face, 2 * ci[c] - ne, 2 * cj[c] - ne, ne);
      vert2elem[verts[g][c]].emplace_back(g, c);
    }
  }
  // edges as corner pairs in FORWARD order: S: SW->SE, N: NW->NE, W: SW->NW, E: SE->NE
  static const int edge_c[4][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}};
  h.nbr.assign(h.nelem, {{-1, -1, -1, -1, -1, -1, -1, -1}});
  h.nbr_pos.assign(h.nelem, {{-1, -1, -1, -1, -1, -1, -1, -1}});
  for (int g = 0; g < h.nelem; ++g) {
    for (int e = 0; e < 4; ++e) {
      const Vec3i A = verts[g][edge_c[e][0]], B = verts[g][edge_c[e][1]];
      int found = -1, fpos = -1;
      bool bwd = false;
      for (auto& ge : vert2elem[A]) {
        const int g2 = ge.first;
        if (g2 == g) continue;
        for (int e2 = 0; e2 < 4; ++e2) {
          const Vec3i A2 = verts[g2][edge_c[e2][0]], B2 = verts[g2][edge_c[e2][1]];
          if ((A2 == A && B2 == B) || (A2 == B && B2 == A)) { found = g2; fpos = e2; bwd = (A2 == B); }
        }
      }
      assert(found >= 0);
      // The reference derives the point ordering from a table; it must agree with geometry.
      assert(bwd == DIR_BWD[e][fpos]);
      (void)bwd;
      h.nbr[g][e] = found; h.nbr_pos[g][e] = fpos;
    }
    for (int c = 0; c < 4; ++c) {
      const Vec3i V = verts[g][c];
      for (auto& ge : vert2elem[V]) {
        const int g2 = ge.first;
        if (g2 == g) continue;
        bool is_edge_nbr = false;
        for (int e = 0; e < 4; ++e) is_edge_nbr |= (h.nbr[g][e] == g2);
        if (is_edge_nbr) continue;
        assert(h.nbr[g][4 + c] < 0);
        h.nbr[g][4 + c] = g2; h.nbr_pos[g][4 + c] = 4 + ge.second;
      }
    }
  }
  // The face curve laid on the six faces so that it runs on continuously from one face to the next:
  // faces in the order 1, 2, 6, 4, 5, 3, each with its own reflection / rotation (cube_mod.F90:1527-1587).
  // (i, j) = 1-based element indices of GridElem(i, j, face); gid = (face-1) ne^2 + (j-1) ne + (i-1).
  {
    const std::vector<int> mesh = face_curve(ne);
    auto M = [&](int i, int j) { return mesh[(i - 1) + (size_t)ne * (j - 1)]; };
    h.sfc_order.assign(h.nelem, -1);
    static const int face_order[6] = {1, 2, 6, 4, 5, 3};
    for (int fo = 0; fo < 6; ++fo) {
      const int face = face_order[fo], offset = fo * ne * ne;
      for (int j = 1; j <= ne; ++j)
        for (int i = 1; i <= ne; ++i) {
          int sc;
          switch (face) {
            case 1: case 2: sc = M(i, ne - j + 1); break;
            case 6: sc = M(ne - i + 1, ne - j + 1); break;
            case 4: sc = M(ne - j + 1, i); break;
            default: sc = M(i, j); break;   // faces 5 and 3
          }
          h.sfc_order[offset + sc] = (face - 1) * ne * ne + (j - 1) * ne + (i - 1);
        }
    }
  }
  // genspacepart (spacecurve_mod.F90:1232-1264): contiguous runs, first nelem%npart parts get +1
  const int npart = h.p.npart;
  h.owner.assign(h.nelem, 0);
  h.gid2lid.assign(h.nelem, -1);
  const int base = h.nelem / npart, extra = h.nelem % npart;
  int pos = 0;
  for (int r = 0; r < npart; ++r) {
    const int cnt = base + (r < extra ? 1 : 0);
    for (int k = 0; k < cnt; ++k, ++pos) {
      const int g = h.sfc_order[pos];
      h.owner[g] = r; h.gid2lid[g] = k;
      if (r == h.p.part_id) h.local_gids.push_back(g);
    }
  }
  h.nelemd = (int)h.local_gids.size();
}

void build_geometry(HommeDriver& h) {
  const int n = h.nelemd;
  // global area correction (prim_driver_mod.F90:401-414): alpha = 4*pi / sum(mp*metdet)
  std::vector<double> Dm(4 * NPSQ), la(NPSQ), lo(NPSQ);
  double area0 = 0.0;
  for (int g = 0; g < h.nelem; ++g) {
    elem_geometry(h, g, 1.0, Dm.data(), la.data(), lo.data());
    for (int pt = 0; pt < NPSQ; ++pt) {
      const double det = Dm[0 * NPSQ + pt] * Dm[3 * NPSQ + pt] - Dm[1 * NPSQ + pt] * Dm[2 * NPSQ + pt];
      area0 += (double)(h.gll.w[pt / NP] * h.gll.w[pt % NP]) * std::fabs(det);
    }
  }
  const double alpha = 4 * DD_PI / area0;
  // rspheremp = 1 / DSS(spheremp) (mass_matrix_mod.F90:85-110)
  std::vector<double> rsph((size_t)h.nelem * NPSQ);
  for (int g = 0; g < h.nelem; ++g) {
    elem_geometry(h, g, alpha, Dm.data(), la.data(), lo.data());
    for (int pt = 0; pt < NPSQ; ++pt) {
      const double det = Dm[0 * NPSQ + pt] * Dm[3 * NPSQ + pt] - Dm[1 * NPSQ + pt] * Dm[2 * NPSQ + pt];
      rsph[(size_t)g * NPSQ + pt] = (double)(h.gll.w[pt / NP] * h.gll.w[pt % NP]) * std::fabs(det);
    }
  }
  global_dss(h, rsph, 1);
  for (auto& x : rsph) x = 1.0 / x;

  h.D.assign((size_t)n * 4 * NPSQ, 0); h.Dinv = h.D; h.metinv = h.D; h.tensorvisc = h.D;
  h.vec_sph2cart.assign((size_t)n * 6 * NPSQ, 0);
  h.fcor.assign((size_t)n * NPSQ, 0);
  h.mp = h.spheremp = h.rspheremp = h.metdet = h.phis = h.lat = h.lon = h.gidf = h.fcor;
  for (int l = 0; l < n; ++l) {
    const int g = h.local_gids[l];
    elem_geometry(h, g, alpha, Dm.data(), &h.lat[(size_t)l * NPSQ], &h.lon[(size_t)l * NPSQ]);
    for (int pt = 0; pt < NPSQ; ++pt) {
      const double d11 = Dm[0 * NPSQ + pt], d12 = Dm[1 * NPSQ + pt], d21 = Dm[2 * NPSQ + pt], d22 = Dm[3 * NPSQ + pt];
      const double det = d11 * d22 - d12 * d21;
      const double Dmath[2][2] = {{d11, d12}, {d21, d22}};
      const double Dimath[2][2] = {{d22 / det, -d12 / det}, {-d21 / det, d11 / det}};
      // met = D^T D ; metinv = adj(met)/det^2   (cube_mod.F90:258-315)
      const double m11 = d11 * d11 + d21 * d21, m12 = d11 * d12 + d21 * d22, m22 = d12 * d12 + d22 * d22;
      const double Mimath[2][2] = {{m22 / (det * det), -m12 / (det * det)}, {-m12 / (det * det), m11 / (det * det)}};
      // F90 X(i,j,r,c) read by C as [a][b][igp][jgp] with (a,b) = (c,r): memory [c][r][pt]
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) {
          const size_t o = ((size_t)l * 4 + c * 2 + r) * NPSQ + pt;
          h.D[o] = Dmath[r][c]; h.Dinv[o] = Dimath[r][c]; h.metinv[o] = Mimath[r][c];
        }
      if (h.p.hypervis_scaling != 0.0) {
        // tensor hyperviscosity V = (DE) (Lam*)^2 Lam (DE)^T from the eigen-decomposition of metinv
        // (cube_mod.F90:315-428); maxloc scans column-major, first maximum wins
        const double (&M)[2][2] = Mimath;
        const double disc = std::sqrt(4.0 * M[0][1] * M[1][0] + (M[0][0] - M[1][1]) * (M[0][0] - M[1][1]));
        const double eig[2] = {(M[0][0] + M[1][1] + disc) / 2.0, (M[0][0] + M[1][1] - disc) / 2.0};
        double DE[2][2] = {{M[0][0] - eig[0], M[0][1]}, {M[1][0], M[1][1] - eig[0]}};
        int ir = 0, ic = 0;
        double mx = -1.0;
        for (int c = 0; c < 2; ++c)
          for (int r = 0; r < 2; ++r)
            if (std::fabs(DE[r][c]) > mx) { mx = std::fabs(DE[r][c]); ir = r; ic = c; }
        double E[2][2];
        if (mx == 0.0) { E[0][0] = 1; E[1][0] = 0; }
        else if (ir == 0 && ic == 0) { E[1][0] = 1; E[0][0] = -DE[1][0] / DE[0][0]; }
        else if (ir == 0 && ic == 1) { E[1][0] = 1; E[0][0] = -DE[1][1] / DE[0][1]; }
        else if (ir == 1 && ic == 0) { E[0][0] = 1; E[1][0] = -DE[0][0] / DE[1][0]; }
        else { E[0][0] = 1; E[1][0] = -DE[0][1] / DE[1][1]; }
        E[0][1] = -E[1][0];
        E[1][1] = E[0][0];
        for (int c = 0; c < 2; ++c) {
          const double nrm = std::sqrt(E[0][c] * E[0][c] + E[1][c] * E[1][c]);
          E[0][c] /= nrm; E[1][c] /= nrm;
        }
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 2; ++c) DE[r][c] = Dmath[r][0] * E[0][c] + Dmath[r][1] * E[1][c];
        const double rearth = 6.376e6;
        double DEL[2][2];
        for (int c = 0; c < 2; ++c) {
          const double lamStar = 1.0 / std::pow(eig[c], h.p.hypervis_scaling / 4.0) * (rearth * rearth);
          for (int r = 0; r < 2; ++r) DEL[r][c] = (lamStar * lamStar) * eig[c] * DE[r][c];
        }
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 2; ++c)
            h.tensorvisc[((size_t)l * 4 + c * 2 + r) * NPSQ + pt] = DEL[r][0] * DE[c][0] + DEL[r][1] * DE[c][1];
        // vec_sphere2cart (cube_mod.F90:172-178), F90 (np,np,3,2): memory [dir][comp][pt]
        const double la = h.lat[(size_t)l * NPSQ + pt], lo = h.lon[(size_t)l * NPSQ + pt];
        const double vs[2][3] = {{-std::sin(lo), std::cos(lo), 0.0},
                                 {-std::sin(la) * std::cos(lo), -std::sin(la) * std::sin(lo), std::cos(la)}};
        for (int d = 0; d < 2; ++d)
          for (int c3 = 0; c3 < 3; ++c3) h.vec_sph2cart[((size_t)l * 6 + d * 3 + c3) * NPSQ + pt] = vs[d][c3];
      }
      const size_t s = (size_t)l * NPSQ + pt;
      const double w = (double)(h.gll.w[pt / NP] * h.gll.w[pt % NP]);
      h.mp[s] = w;
      h.metdet[s] = std::fabs(det);
      h.spheremp[s] = w * std::fabs(det);
      h.rspheremp[s] = rsph[(size_t)g * NPSQ + pt];
      h.fcor[s] = 2.0 * OMEGA * std::sin(h.lat[s]);
      h.gidf[s] = g;
    }
  }
  // add_connection tuples, one per directed GridEdge whose head is local
  // (prim_cxx_driver_mod.F90:120-140): lid,gid,pos,pid of both ends, all 1-based.
  h.conn.clear();
  for (int l = 0; l < n; ++l) {
    const int g = h.local_gids[l];
    for (int c = 0; c < 8; ++c) {
      const int g2 = h.nbr[g][c];
      if (g2 < 0) continue;
      const int own2 = h.owner[g2];
      const int lid2 = own2 == h.p.part_id ? h.gid2lid[g2] : h.gid2lid[g2];
      const int t[8] = {l + 1, g + 1, c_pos_to_f90(c), h.p.part_id + 1,
                        lid2 + 1, g2 + 1, c_pos_to_f90(h.nbr_pos[g][c]), own2 + 1};
      h.conn.insert(h.conn.end(), t, t + 8);
    }
  }
}

// element-local vorticity of (u,0)-type fields for the q2 tracer (compute_zeta_C0): strong
// form with the covariant transform (derivative_mod_base.F90 vorticity_sphere), global DSS.
void jw_init(HommeDriver& h) {
  const HommeParams& p = h.p;
  const int nlev = p.nlev, n = h.nelemd;
  const double u0 = 35.0, t0 = 288.0, gama = 0.005, ddt = 4.8e5, eta_t = 0.2, eta_s = 1.0, eta_0 = 0.252;
  const double latc = DD_PI * (2.0 / 9.0), lonc = DD_PI * (1.0 / 9.0);
  std::vector<double> eta(nlev), etv(nlev), tbar(nlev);
  for (int k = 0; k < nlev; ++k) {
    eta[k] = h.hyam[k] + h.hybm[k];
    etv[k] = (eta[k] - eta_0) * DD_PI * 0.5;
    tbar[k] = t0 * std::pow(eta[k], RGAS * gama / GRAV);
    if (eta[k] <= eta_t) tbar[k] += ddt * std::pow(eta_t - eta[k], 5);
  }
  auto u_at = [&](double lat, double lon, int k) {
    const double aa = std::sin(latc) * std::sin(lat) + std::cos(latc) * std::cos(lat) * std::cos(lon - lonc);
    const double rc = 10.0 * std::acos(std::min(1.0, std::max(-1.0, aa)));
    const double s2 = std::sin(2.0 * lat);
    return u0 * std::pow(std::cos(etv[k]), 1.5) * s2 * s2 + p.u_perturb * std::exp(-rc * rc);
  };
  auto t_at = [&](double lat, int k) {
    const double sn = std::sin(lat), cs = std::cos(lat);
    const double trm1 = 0.75 * (eta[k] * DD_PI * u0 / RGAS) * std::sin(etv[k]) * std::sqrt(std::cos(etv[k]));
    const double trm2 = -2.0 * std::pow(sn, 6) * (cs * cs + 1.0 / 3.0) + 10.0 / 63.0;
    const double trm3 = 2.0 * u0 * std::pow(std::cos(etv[k]), 1.5);
    const double trm4 = (1.60 * cs * cs * cs * (sn * sn + 2.0 / 3.0) - DD_PI * 0.25) * REARTH * OMEGA;
    return tbar[k] + trm1 * (trm2 * trm3 + trm4);
  };
  h.v.assign((size_t)n * 3 * nlev * 2 * NPSQ, 0.0);
  h.T.assign((size_t)n * 3 * nlev * NPSQ, 0.0);
  h.dp3d.assign((size_t)n * 3 * nlev * NPSQ, 0.0);
  h.ps_v.assign((size_t)n * 3 * NPSQ, P0);
  h.Qdp.assign((size_t)n * 2 * p.qsize_d * nlev * NPSQ, 0.0);
  h.Q.assign((size_t)n * p.qsize_d * nlev * NPSQ, 0.0);
  h.omega_p.assign((size_t)n * nlev * NPSQ, 0.0);
  for (auto& a : h.accum) a.assign((size_t)n * 4 * std::max(1, p.qsize_d) * NPSQ, 0.0);
  h.FM.assign((size_t)n * nlev * 2 * NPSQ, 0.0);
  h.FT.assign((size_t)n * nlev * NPSQ, 0.0);
  h.FQ.clear();  // 40 tiles per element: allocated on first use (need_FQ), a standalone run never touches it
  for (int l = 0; l < n; ++l)
    for (int pt = 0; pt < NPSQ; ++pt) {
      const double lat = h.lat[(size_t)l * NPSQ + pt], lon = h.lon[(size_t)l * NPSQ + pt];
      const double sn = std::sin(lat), cs = std::cos(lat);
      const double trm1 = u0 * std::pow(std::cos((eta_s - eta_0) * DD_PI * 0.5), 1.5);
      const double trm2 = -2.0 * std::pow(sn, 6) * (cs * cs + 1.0 / 3.0) + 10.0 / 63.0;
      const double trm3 = (1.60 * cs * cs * cs * (sn * sn + 2.0 / 3.0) - DD_PI * 0.25) * REARTH * OMEGA;
      h.phis[(size_t)l * NPSQ + pt] = trm1 * (trm2 * trm1 + trm3);
      for (int tl = 0; tl < 3; ++tl)
        for (int k = 0; k < nlev; ++k) {
          h.v[((((size_t)l * 3 + tl) * nlev + k) * 2 + 0) * NPSQ + pt] = u_at(lat, lon, k);
          h.T[(((size_t)l * 3 + tl) * nlev + k) * NPSQ + pt] = t_at(lat, k);
          // dp3d = delta(hyai)*ps0 + delta(hybi)*ps_v  (prim_driver_mod.F90:948-958)
          h.dp3d[(((size_t)l * 3 + tl) * nlev + k) * NPSQ + pt] =
              (h.hyai[k + 1] - h.hyai[k]) * P0 + (h.hybi[k + 1] - h.hybi[k]) * P0;
        }
    }
  if (p.qsize <= 0) return;
  // tracers: q_i = T/400 for all i; q2 = clipped DSS'd vorticity / 2e-5; q3 = 1
  // (baroclinic_inst_mod.F90:195-236). Vorticity needs the global velocity field.
  std::vector<double> zeta;  // [gid][16][nlev]
  if (p.qsize >= 2) {
    zeta.assign((size_t)h.nelem * NPSQ * nlev, 0.0);
    std::vector<double> Dm(4 * NPSQ), la(NPSQ), lo(NPSQ), ucov(2 * NPSQ);
    // alpha: recover from local metdet if available, else recompute globally
    double area0 = 0.0;
    for (int g = 0; g < h.nelem; ++g) {
      elem_geometry(h, g, 1.0, Dm.data(), la.data(), lo.data());
      for (int pt = 0; pt < NPSQ; ++pt)
        area0 += (double)(h.gll.w[pt / NP] * h.gll.w[pt % NP]) *
                 std::fabs(Dm[0 * NPSQ + pt] * Dm[3 * NPSQ + pt] - Dm[1 * NPSQ + pt] * Dm[2 * NPSQ + pt]);
    }
    const double alpha = 4 * DD_PI / area0;
    std::vector<double> sph((size_t)h.nelem * NPSQ);
    for (int g = 0; g < h.nelem; ++g) {
      elem_geometry(h, g, alpha, Dm.data(), la.data(), lo.data());
      for (int k = 0; k < nlev; ++k) {
        for (int pt = 0; pt < NPSQ; ++pt) {
          const double u = u_at(la[pt], lo[pt], k), vv = 0.0;
          // covariant components: D^T (u,v)
          ucov[0 * NPSQ + pt] = Dm[0 * NPSQ + pt] * u + Dm[2 * NPSQ + pt] * vv;
          ucov[1 * NPSQ + pt] = Dm[1 * NPSQ + pt] * u + Dm[3 * NPSQ + pt] * vv;
        }
        for (int igp = 0; igp < NP; ++igp)
          for (int jgp = 0; jgp < NP; ++jgp) {
            double dvdx = 0, dudy = 0;
            for (int m = 0; m < NP; ++m) {
              dvdx += h.gll.dvv[jgp][m] * ucov[1 * NPSQ + igp * NP + m];
              dudy += h.gll.dvv[igp][m] * ucov[0 * NPSQ + m * NP + jgp];
            }
            const int pt = igp * NP + jgp;
            const double det = std::fabs(Dm[0 * NPSQ + pt] * Dm[3 * NPSQ + pt] - Dm[1 * NPSQ + pt] * Dm[2 * NPSQ + pt]);
            const double w = (double)(h.gll.w[igp] * h.gll.w[jgp]);
            if (k == 0) sph[(size_t)g * NPSQ + pt] = w * det;
            zeta[((size_t)g * NPSQ + pt) * nlev + k] = (dvdx - dudy) * (1.0 / det) * (1.0 / REARTH) * (w * det);
          }
      }
    }
    global_dss(h, zeta, nlev);
    global_dss(h, sph, 1);
    for (int g = 0; g < h.nelem; ++g)
      for (int pt = 0; pt < NPSQ; ++pt)
        for (int k = 0; k < nlev; ++k) zeta[((size_t)g * NPSQ + pt) * nlev + k] /= sph[(size_t)g * NPSQ + pt];
  }
  for (int l = 0; l < n; ++l) {
    const int g = h.local_gids[l];
    for (int q = 0; q < p.qsize; ++q)
      for (int k = 0; k < nlev; ++k)
        for (int pt = 0; pt < NPSQ; ++pt) {
          double val = h.T[(((size_t)l * 3 + 0) * nlev + k) * NPSQ + pt] / 400.0;
          if (q == 1) {
            const double z = zeta[((size_t)g * NPSQ + pt) * nlev + k];
            val = z < 0 ? 0.0 : z / 2e-5;
          }
          if (q == 2) val = 1.0;
          h.Q[(((size_t)l * p.qsize_d + q) * nlev + k) * NPSQ + pt] = val;
          const double dp = h.dp3d[(((size_t)l * 3 + 0) * nlev + k) * NPSQ + pt];
          for (int tq = 0; tq < 2; ++tq)
            h.Qdp[((((size_t)l * 2 + tq) * p.qsize_d + q) * nlev + k) * NPSQ + pt] = val * dp;
        }
  }
}

template <typename F>
F get_sym(HommeDriver* h, const char* name) {
  auto it = h->sym.find(name);
  if (it == h->sym.end()) { std::fprintf(stderr, "homme_driver: symbol %s not bound\n", name); std::abort(); }
  return reinterpret_cast<F>(it->second);
}

const char* const ABI_SYMBOLS[] = {
    "reset_cxx_comm", "initialize_hommexx_session", "finalize_hommexx_session", "init_connectivity",
    "add_connection", "finalize_connectivity", "init_derivative_c", "init_simulation_params_c",
    "init_elements_2d_c", "init_elements_states_c", "init_diagnostics_c", "init_hvcoord_c",
    "init_boundary_exchanges_c", "init_time_level_c", "prim_run_subcycle_c", "cxx_push_results_to_f90",
    "f90_push_forcing_to_cxx", "cxx_push_forcing_to_f90", "hommexx_b200_nlev", "hommexx_b200_qsize_d"};

}  // namespace

extern "C" {

HommeDriver* hd_create(const HommeParams* p, const double* hyai, const double* hybi, const double* hyam,
                       const double* hybm) {
  auto* h = new HommeDriver;
  h->p = *p;
  h->gll = make_gll();
  h->hyai.assign(hyai, hyai + p->nlev + 1);
  h->hybi.assign(hybi, hybi + p->nlev + 1);
  h->hyam.assign(hyam, hyam + p->nlev);
  h->hybm.assign(hybm, hybm + p->nlev);
  build_topology(*h);
  build_geometry(*h);
  return h;
}

void hd_destroy(HommeDriver* h) {
  if (!h) return;
  if (h->lib) dlclose(h->lib);
  delete h;
}

void hd_init_jw(HommeDriver* h) { jw_init(*h); }

const char* hd_last_error() { return g_last_error.c_str(); }

int hd_bind(HommeDriver* h, const char* libpath) {
  h->lib = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
  if (!h->lib) { g_last_error = std::string("dlopen failed: ") + dlerror(); return 1; }
  for (const char* s : ABI_SYMBOLS) {
    void* f = dlsym(h->lib, s);
    if (!f) { g_last_error = std::string("missing symbol: ") + s; return 2; }
    h->sym[s] = f;
  }
  const int nlev = get_sym<int (*)()>(h, "hommexx_b200_nlev")();
  const int qd = get_sym<int (*)()>(h, "hommexx_b200_qsize_d")();
  if (nlev != h->p.nlev || qd != h->p.qsize_d) {
    g_last_error = "library built for nlev=" + std::to_string(nlev) + " qsize_d=" + std::to_string(qd) +
                   " but driver configured for nlev=" + std::to_string(h->p.nlev) +
                   " qsize_d=" + std::to_string(h->p.qsize_d);
    return 3;
  }
  return 0;
}

void hd_upload_state(HommeDriver* h) {
  using P = const double*;
  P v = h->v.data(), T = h->T.data(), dp = h->dp3d.data(), q = h->Qdp.data(), ps = h->ps_v.data();
  get_sym<void (*)(P const*, P const*, P const*, P const*, P const*)>(h, "init_elements_states_c")(&v, &T, &dp, &q, &ps);
  const int nstep0 = 2;
  get_sym<void (*)(const int*, const int*, const int*, const int*, const int*)>(h, "init_time_level_c")(
      &h->nm1, &h->n0, &h->np1, &h->nstep, &nstep0);
}

void hd_init_dycore(HommeDriver* h) {
  const HommeParams& p = h->p;
  using P = const double*;
  const int fcomm = 0;
  get_sym<void (*)(const int*)>(h, "reset_cxx_comm")(&fcomm);
  get_sym<void (*)()>(h, "initialize_hommexx_session")();
  // init_cxx_connectivity (prim_cxx_driver_mod.F90:43)
  get_sym<void (*)(const int*)>(h, "init_connectivity")(&h->nelemd);
  auto addc = get_sym<void (*)(const int*, const int*, const int*, const int*, const int*, const int*,
                               const int*, const int*)>(h, "add_connection");
  for (size_t i = 0; i < h->conn.size(); i += 8) {
    const int* t = &h->conn[i];
    addc(t, t + 1, t + 2, t + 3, t + 4, t + 5, t + 6, t + 7);
  }
  get_sym<void (*)()>(h, "finalize_connectivity")();
  // prim_init2 tail (prim_driver_mod.F90:1043-1104)
  P dvv = &h->gll.dvv[0][0];
  get_sym<void (*)(P const*)>(h, "init_derivative_c")(&dvv);
  // namelist post-processing (namelist_mod.F90:970-972): negative nu_* default to nu
  double nu_s = p.nu_s < 0 ? p.nu : p.nu_s, nu_q = p.nu_q < 0 ? p.nu : p.nu_q, nu_div = p.nu_div < 0 ? p.nu : p.nu_div;
  const bool pw = p.prescribed_wind, moist = p.moisture, dd = p.disable_diagnostics, cps = p.use_cpstar,
             sl = p.use_semi_lagrangian_transport;
  get_sym<void (*)(const int*, const int*, const int*, const int*, const int*, const int*, const int*, const int*,
                   const double*, const double*, const double*, const double*, const double*, const double*,
                   const int*, const int*, const double*, const int*, const bool*, const bool*, const bool*,
                   const bool*, const bool*)>(h, "init_simulation_params_c")(
      &p.remap_alg, &p.limiter_option, &p.rsplit, &p.qsplit, &p.time_step_type, &p.energy_fixer, &p.qsize,
      &p.state_frequency, &p.nu, &p.nu_p, &nu_q, &nu_s, &nu_div, &p.nu_top, &p.hypervis_order,
      &p.hypervis_subcycle, &p.hypervis_scaling, &p.ftype, &pw, &moist, &dd, &cps, &sl);
  P D = h->D.data(), Di = h->Dinv.data(), fc = h->fcor.data(), mp = h->mp.data(), sp = h->spheremp.data(),
    rsp = h->rspheremp.data(), md = h->metdet.data(), mi = h->metinv.data(), ph = h->phis.data(),
    tv = h->tensorvisc.data(), vs = h->vec_sph2cart.data();
  const bool consthv = (p.hypervis_scaling == 0.0);
  get_sym<void (*)(const int*, P const*, P const*, P const*, P const*, P const*, P const*, P const*, P const*,
                   P const*, P const*, P const*, const bool*)>(h, "init_elements_2d_c")(
      &h->nelemd, &D, &Di, &fc, &mp, &sp, &rsp, &md, &mi, &ph, &tv, &vs, &consthv);
  {
    P v = h->v.data(), T = h->T.data(), dp = h->dp3d.data(), q = h->Qdp.data(), ps = h->ps_v.data();
    get_sym<void (*)(P const*, P const*, P const*, P const*, P const*)>(h, "init_elements_states_c")(&v, &T, &dp, &q, &ps);
  }
  {
    double* a[8] = {h->Q.data(), h->accum[0].data(), h->accum[1].data(), h->accum[2].data(),
                    h->accum[3].data(), h->accum[4].data(), h->accum[5].data(), h->accum[6].data()};
    get_sym<void (*)(double* const*, double* const*, double* const*, double* const*, double* const*,
                     double* const*, double* const*, double* const*)>(h, "init_diagnostics_c")(
        &a[0], &a[1], &a[2], &a[3], &a[4], &a[5], &a[6], &a[7]);
  }
  {
    const double ps0 = P0;
    P am = h->hyam.data(), ai = h->hyai.data(), bm = h->hybm.data(), bi = h->hybi.data();
    get_sym<void (*)(const double*, P const*, P const*, P const*, P const*)>(h, "init_hvcoord_c")(&ps0, &am, &ai, &bm, &bi);
  }
  get_sym<void (*)()>(h, "init_boundary_exchanges_c")();
  const int nstep0 = 2;
  get_sym<void (*)(const int*, const int*, const int*, const int*, const int*)>(h, "init_time_level_c")(
      &h->nm1, &h->n0, &h->np1, &h->nstep, &nstep0);
}

int hd_run_subcycle(HommeDriver* h) {
  // prim_main.F90:300-308
  const int last = h->last_step;
  int nstep_c, nm1_c, n0_c, np1_c;
  get_sym<void (*)(const double*, int*, int*, int*, int*, const int*)>(h, "prim_run_subcycle_c")(
      &h->p.tstep, &nstep_c, &nm1_c, &n0_c, &np1_c, &last);
  h->nstep = nstep_c; h->nm1 = nm1_c + 1; h->n0 = n0_c + 1; h->np1 = np1_c + 1;
  return h->nstep;
}

void hd_push_results(HommeDriver* h) {
  double *v = h->v.data(), *T = h->T.data(), *dp = h->dp3d.data(), *q = h->Qdp.data(), *Q = h->Q.data(),
         *ps = h->ps_v.data(), *om = h->omega_p.data();
  get_sym<void (*)(double* const*, double* const*, double* const*, double* const*, double* const*,
                   double* const*, double* const*)>(h, "cxx_push_results_to_f90")(&v, &T, &dp, &q, &Q, &ps, &om);
}

// prim_driver_mod.F90:1380 / :1402 (the CAM-coupled prim_run_subcycle wrapper)
static void need_FQ(HommeDriver* h) {
  const size_t n = (size_t)h->nelemd * h->p.qsize_d * h->p.nlev * NPSQ;
  if (h->FQ.size() != n) h->FQ.assign(n, 0.0);
}
void hd_push_forcing(HommeDriver* h) {
  need_FQ(h);
  get_sym<void (*)(double*, double*, double*, double*)>(h, "f90_push_forcing_to_cxx")(h->FM.data(), h->FT.data(),
                                                                                     h->FQ.data(), h->Qdp.data());
}
void hd_pull_forcing(HommeDriver* h) {
  need_FQ(h);
  get_sym<void (*)(double*, double*, double*)>(h, "cxx_push_forcing_to_f90")(h->FM.data(), h->FT.data(), h->FQ.data());
}
void hd_set_last_step(HommeDriver* h, int nEndStep) { h->last_step = nEndStep; }

void hd_finalize_dycore(HommeDriver* h) { get_sym<void (*)()>(h, "finalize_hommexx_session")(); }

int hd_nelemd(const HommeDriver* h) { return h->nelemd; }
int hd_nelem_global(const HommeDriver* h) { return h->nelem; }

double* hd_array(HommeDriver* h, const char* name, int64_t* n) {
  std::vector<double>* a = nullptr;
  const std::string s(name);
  if (s == "D") a = &h->D; else if (s == "Dinv") a = &h->Dinv; else if (s == "fcor") a = &h->fcor;
  else if (s == "mp") a = &h->mp; else if (s == "spheremp") a = &h->spheremp;
  else if (s == "rspheremp") a = &h->rspheremp; else if (s == "metdet") a = &h->metdet;
  else if (s == "metinv") a = &h->metinv; else if (s == "phis") a = &h->phis; else if (s == "v") a = &h->v;
  else if (s == "T") a = &h->T; else if (s == "dp3d") a = &h->dp3d; else if (s == "Qdp") a = &h->Qdp;
  else if (s == "Q") a = &h->Q; else if (s == "ps_v") a = &h->ps_v; else if (s == "omega_p") a = &h->omega_p;
  else if (s == "lat") a = &h->lat; else if (s == "lon") a = &h->lon; else if (s == "gid") a = &h->gidf;
  else if (s == "tensorvisc") a = &h->tensorvisc; else if (s == "vec_sph2cart") a = &h->vec_sph2cart;
  else if (s == "FM") a = &h->FM; else if (s == "FT") a = &h->FT; else if (s == "FQ") { need_FQ(h); a = &h->FQ; }
  else if (s == "Qvar") a = &h->accum[0]; else if (s == "Qmass") a = &h->accum[1]; else if (s == "Q1mass") a = &h->accum[2];
  else if (s == "IEner") a = &h->accum[3]; else if (s == "IEner_wet") a = &h->accum[4];
  else if (s == "KEner") a = &h->accum[5]; else if (s == "PEner") a = &h->accum[6];
  else if (s == "dvv") { if (n) *n = 16; return &h->gll.dvv[0][0]; }
  if (!a) { if (n) *n = 0; return nullptr; }
  if (n) *n = (int64_t)a->size();
  return a->data();
}

int hd_connections(const HommeDriver* h, const int** tuples) {
  *tuples = h->conn.data();
  return (int)(h->conn.size() / 8);
}

void hd_time_levels(const HommeDriver* h, int* nstep, int* nm1, int* n0, int* np1) {
  *nstep = h->nstep; *nm1 = h->nm1; *n0 = h->n0; *np1 = h->np1;
}

// restart runs (runtype = 1): the time levels ReadRestart (restart_io_mod.F90:669-698) restores with elem%state;
// call before hd_init_dycore / hd_upload_state, which hand them to init_time_level_c. 1-based like hd_time_levels.
void hd_set_time_levels(HommeDriver* h, int nstep, int nm1, int n0, int np1) {
  h->nstep = nstep; h->nm1 = nm1; h->n0 = n0; h->np1 = np1;
}

const int* hd_local_gids(const HommeDriver* h) { return h->local_gids.data(); }
const int* hd_owner(const HommeDriver* h) { return h->owner.data(); }

}  // extern "C"
