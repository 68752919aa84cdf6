// Boundary exchange (DSS) — replaces mpi/BoundaryExchange.{hpp,cpp}, BuffersManager and
// Connectivity of the reference.
//
// Design (not the reference's pack -> buffer -> unpack): the exchange is NODE-CENTRIC. At
// init_boundary_exchanges_c the 8-neighbour connection list is turned into one record per
// unique GLL boundary node (2 sharers on an edge, 4 at an element corner, 3 at a cube vertex).
// One thread handles (node, level): it loads each member's value once, forms for every local
// member the sum in the reference's unpack order (own value, then edges S,N,W,E for k=0..3,
// then corners — BoundaryExchange.cpp:512-524, so results are bit-identical), applies
// rspheremp and stores. Traffic: 12/16 of a field read + 12/16 written, no staging buffers for
// on-rank neighbours (the reference's `local_buffer`, BoundaryExchange.cpp:327, disappears).
// Interior points never take part; when rspheremp is requested their scaling is folded into
// the producing kernel. Members owned by another rank arrive through a halo buffer filled by a
// pack kernel + grouped ncclSend/ncclRecv over NVLink (one message per neighbour rank, slot
// order agreed by sorting connections on (owner gid, owner pos), cf. BoundaryExchange.cpp:1050).
#include <algorithm>
#include <array>
#include <cstring>
#include <map>
#include <tuple>
#include <unordered_map>

#include "hxx.cuh"
#ifdef HXX_WITH_NCCL
#include <nccl.h>
#endif

namespace hxx {

// ConnectivityHelpers.hpp:145-176
static const int EDGE_PTS[4][4] = {{0, 1, 2, 3}, {12, 13, 14, 15}, {0, 4, 8, 12}, {3, 7, 11, 15}};
static const int CORNER_PTS[4] = {0, 3, 12, 15};


#ifndef HXX_DSS_FPB
#define HXX_DSS_FPB 8
#endif
constexpr int DSS_FPB = HXX_DSS_FPB;  // fields per thread (grid.y chunks)

#ifndef HXX_DSS_PAIR_UB
#define HXX_DSS_PAIR_UB 4
#endif
#ifndef HXX_DSS_QUAD_UB
#define HXX_DSS_QUAD_UB 2
#endif
#ifndef HXX_DSS_MINB
#define HXX_DSS_MINB 5
#endif
constexpr int DSS_TPB = 128;

// A thread owns DSS_V consecutive levels of its node (16-byte loads and stores: a column is NLEV * 8 bytes, so
// every even level is 16-byte aligned): the same number of memory instructions and address registers keeps
// twice the bytes in flight, which is what a pass with no arithmetic to hide its latency behind needs.
constexpr int DSS_V = (NLEV % 2 == 0) ? 2 : 1;
constexpr int DSS_NLV = NLEV / DSS_V;  // level groups per node
struct DV { double v[DSS_V]; };
__device__ __forceinline__ DV ldv(const double* p) {
  DV r;
  if constexpr (DSS_V == 2) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}
__device__ __forceinline__ void stv(double* p, const DV& x) {
  if constexpr (DSS_V == 2) *reinterpret_cast<double2*>(p) = make_double2(x.v[0], x.v[DSS_V - 1]);
  else *p = x.v[0];
}

// Generic nodes: any number of sharers up to four, any of them in the halo, explicit summation orders.
template <bool RSP, bool AVG>
__device__ __forceinline__ void dss_nodes_body(const DssNode* __restrict__ nodes, int nnodes, const FieldList& fl,
                                               const double* __restrict__ geo, const double* __restrict__ halo,
                                               long long g, int ychunk) {
  const int node = (int)(g / DSS_NLV), k = (int)(g % DSS_NLV) * DSS_V;
  if (node >= nnodes) return;
  const DssNode nd = nodes[node];
  double rs[4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
    rs[m] = (RSP && m < nd.nmem && nd.src[m] >= 0) ? __ldg(geo + (size_t)nd.src[m] * GEO_N + G_RSPHEREMP) : 1.0;
  const int f0 = ychunk * DSS_FPB, f1 = min(fl.nf, f0 + DSS_FPB);
  for (int f = f0; f < f1; ++f) {
    // every load is issued before the first store (the fields may alias as far as the compiler knows,
    // which would otherwise serialise load -> store -> load)
    DV val[4], avg[4];
    double* ptr[4];
    double* base = fl.base[f];
    const long long es = fl.estride[f];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ptr[m] = nullptr;
#pragma unroll
      for (int i = 0; i < DSS_V; ++i) { val[m].v[i] = 0.0; avg[m].v[i] = 0.0; }
      if (m < nd.nmem) {
        const int s = nd.src[m];
        if (s >= 0) {
          ptr[m] = base + (size_t)(s >> 4) * es + (s & 15) * NLEV + k;
          val[m] = ldv(ptr[m]);
          if (AVG && f < fl.navg) avg[m] = ldv(ptr[m] + fl.avg_delta);
        } else {
          val[m] = ldv(halo + ((size_t)(~s) * fl.nf + f) * NLEV + k);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (ptr[m]) {
        DV acc = val[m];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int o = nd.ord[m][t];
          if (o < 4) {
#pragma unroll
            for (int i = 0; i < DSS_V; ++i)
              acc.v[i] += o == 0 ? val[0].v[i] : o == 1 ? val[1].v[i] : o == 2 ? val[2].v[i] : val[3].v[i];
          }
        }
#pragma unroll
        for (int i = 0; i < DSS_V; ++i) {
          if (RSP) acc.v[i] *= rs[m];
          if (AVG && f < fl.navg) acc.v[i] = (avg[m].v[i] + 2.0 * acc.v[i]) / 3.0;  // qdp_time_avg, EulerStepFunctorImpl.hpp:379-403
        }
        stv(ptr[m], acc);
      }
    }
  }
}

// ---- lean paths -------------------------------------------------------------------------------
// 80 % of the boundary nodes are edge nodes with two on-rank sharers and nearly all the others
// are regular element corners with four: they get branch-free code with a handful of
// instructions per (node, level, field). What is left (cube vertices with three sharers, nodes
// with a sharer on another rank) goes through the generic body above.
//
// Pair: both results are a + b (addition commutes, so the reference's "own value first" order
// gives the same bits for both members).
// Quad: members are stored in a canonical order (0; 1 = its W/E-edge neighbour; 2 = its
// S/N-edge neighbour; 3 = the diagonal one), so member 0 always sums as ((v0 + v2) + v1) + v3
// (unpack order BoundaryExchange.cpp:512-524: S/N edge, W/E edge, corner) and each other member
// needs one bit saying which of its two edge neighbours comes first (it differs where the
// element-local axes rotate across cube edges).
struct DssPair { int a, b; };
struct DssQuad { int m[4]; int swaps; int pad[3]; };
static_assert(sizeof(DssQuad) == 32, "DssQuad layout");

template <bool RSP, bool AVG>
__device__ __forceinline__ void dss_pair_apply(const DssPair pr, const FieldList& fl, const double* __restrict__ geo, int k,
                                               int ychunk);
template <bool RSP, bool AVG>
__device__ __forceinline__ void dss_quad_apply(const DssQuad qd, const FieldList& fl, const double* __restrict__ geo, int k,
                                               int ychunk);
template <bool RSP, bool AVG>
__device__ __forceinline__ void dss_pair_body(const DssPair* __restrict__ pairs, int npairs, const FieldList& fl,
                                              const double* __restrict__ geo, long long g, int ychunk) {
  const int ip = (int)(g / DSS_NLV), k = (int)(g % DSS_NLV) * DSS_V;
  if (ip >= npairs) return;
  dss_pair_apply<RSP, AVG>(pairs[ip], fl, geo, k, ychunk);
}
template <bool RSP, bool AVG>
__device__ __forceinline__ void dss_pair_apply(const DssPair pr, const FieldList& fl, const double* __restrict__ geo, int k,
                                               int ychunk) {
  const long long ea = pr.a >> 4, eb = pr.b >> 4;
  const int ca = (pr.a & 15) * NLEV + k, cb = (pr.b & 15) * NLEV + k;
  double ra = 1.0, rb = 1.0;
  if (RSP) {
    ra = __ldg(geo + (size_t)pr.a * GEO_N + G_RSPHEREMP);
    rb = __ldg(geo + (size_t)pr.b * GEO_N + G_RSPHEREMP);
  }
  const int f0 = ychunk * DSS_FPB, f1 = min(fl.nf, f0 + DSS_FPB);
  constexpr int UB = HXX_DSS_PAIR_UB;  // fields whose loads are issued before the first store
  for (int fb = f0; fb < f1; fb += UB) {
    double *pa[UB], *pb[UB];
    DV va[UB], vb[UB], qa[UB], qb[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int f = min(fb + j, f1 - 1);
      double* base = fl.base[f];
      const long long es = fl.estride[f];
      pa[j] = base + ea * es + ca;
      pb[j] = base + eb * es + cb;
      va[j] = ldv(pa[j]);
      vb[j] = ldv(pb[j]);
      if (AVG) {
        if (f < fl.navg) { qa[j] = ldv(pa[j] + fl.avg_delta); qb[j] = ldv(pb[j] + fl.avg_delta); }
        else { qa[j] = va[j]; qb[j] = vb[j]; }
      }
    }
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      if (fb + j < f1) {
        DV xa, xb;
#pragma unroll
        for (int i = 0; i < DSS_V; ++i) {
          const double s = va[j].v[i] + vb[j].v[i];
          xa.v[i] = s; xb.v[i] = s;
          if (RSP) { xa.v[i] = s * ra; xb.v[i] = s * rb; }
          if (AVG && fb + j < fl.navg) {  // qdp_time_avg, EulerStepFunctorImpl.hpp:379-403
            xa.v[i] = (qa[j].v[i] + 2.0 * xa.v[i]) / 3.0;
            xb.v[i] = (qb[j].v[i] + 2.0 * xb.v[i]) / 3.0;
          }
        }
        stv(pa[j], xa);
        stv(pb[j], xb);
      }
    }
  }
}

template <bool RSP, bool AVG>
__device__ __forceinline__ void dss_quad_body(const DssQuad* __restrict__ quads, int nquads, const FieldList& fl,
                                              const double* __restrict__ geo, long long g, int ychunk) {
  const int iq = (int)(g / DSS_NLV), k = (int)(g % DSS_NLV) * DSS_V;
  if (iq >= nquads) return;
  dss_quad_apply<RSP, AVG>(quads[iq], fl, geo, k, ychunk);
}
template <bool RSP, bool AVG>
__device__ __forceinline__ void dss_quad_apply(const DssQuad qd, const FieldList& fl, const double* __restrict__ geo, int k,
                                               int ychunk) {
  long long e[4];
  int c[4];
  double rs[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    e[m] = qd.m[m] >> 4;
    c[m] = (qd.m[m] & 15) * NLEV + k;
    rs[m] = RSP ? __ldg(geo + (size_t)qd.m[m] * GEO_N + G_RSPHEREMP) : 1.0;
  }
  const bool s1 = qd.swaps & 1, s2 = qd.swaps & 2, s3 = qd.swaps & 4;
  const int f0 = ychunk * DSS_FPB, f1 = min(fl.nf, f0 + DSS_FPB);
  constexpr int UB = HXX_DSS_QUAD_UB;
  for (int fb = f0; fb < f1; fb += UB) {
    double* ptr[UB][4];
    DV v[UB][4], qa[UB][4];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int f = min(fb + j, f1 - 1);
      double* base = fl.base[f];
      const long long es = fl.estride[f];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        ptr[j][m] = base + e[m] * es + c[m];
        v[j][m] = ldv(ptr[j][m]);
        if (AVG) qa[j][m] = f < fl.navg ? ldv(ptr[j][m] + fl.avg_delta) : v[j][m];
      }
    }
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      if (fb + j < f1) {
        DV x[4];
#pragma unroll
        for (int i = 0; i < DSS_V; ++i) {
          const double v0 = v[j][0].v[i], v1 = v[j][1].v[i], v2 = v[j][2].v[i], v3 = v[j][3].v[i];
          x[0].v[i] = ((v0 + v2) + v1) + v3;
          x[1].v[i] = s1 ? ((v1 + v0) + v3) + v2 : ((v1 + v3) + v0) + v2;
          x[2].v[i] = s2 ? ((v2 + v3) + v0) + v1 : ((v2 + v0) + v3) + v1;
          x[3].v[i] = s3 ? ((v3 + v2) + v1) + v0 : ((v3 + v1) + v2) + v0;
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          DV r = x[m];
#pragma unroll
          for (int i = 0; i < DSS_V; ++i) {
            if (RSP) r.v[i] *= rs[m];
            if (AVG && fb + j < fl.navg) r.v[i] = (qa[j][m].v[i] + 2.0 * r.v[i]) / 3.0;
          }
          stv(ptr[j][m], r);
        }
      }
    }
  }
}

// PARTS selects what a launch covers: 1 = the pair list, 2 = the quad list, 4 = the generic nodes; the lists
// present are consecutive block ranges (the split is block-uniform, so nothing diverges). Pairs are a launch of
// their own — their body needs far fewer registers; the few generic nodes of a single-rank run ride with the
// quads, on several ranks they wait for the halo. Measured at ne30 / 41 fields (profiles/r2_dss_experiments.txt):
// the pass moves its 0.75 R + 0.75 W of every field at 4.2 TB/s whatever the unroll (2 / 4 fields in flight),
// the fields per thread (4 / 8 / 16), the vector width, or the list order (pairs and quads interleaved in element
// order was 5 % slower): 576-byte columns scattered over gigabytes are what bounds it, not this kernel's shape.
struct DssLists {
  const DssPair* pairs; int npairs, nb_pair;
  const DssQuad* quads; int nquads, nb_quad;
  const DssNode* nodes; int nnodes;
  const double* halo;
};
template <int PARTS, bool RSP, bool AVG>
__global__ void __launch_bounds__(DSS_TPB, PARTS == 1 ? (AVG ? 4 : 8) : (AVG ? 4 : HXX_DSS_MINB)) dss_kernel(DssLists L, FieldList fl,
                                                                                    const double* __restrict__ geo) {
  int b = blockIdx.x;
  if constexpr (PARTS & 1) {
    if (b < L.nb_pair) {
      dss_pair_body<RSP, AVG>(L.pairs, L.npairs, fl, geo, (long long)b * DSS_TPB + threadIdx.x, blockIdx.y);
      return;
    }
    b -= L.nb_pair;
  }
  if constexpr (PARTS & 2) {
    if (b < L.nb_quad) {
      dss_quad_body<RSP, AVG>(L.quads, L.nquads, fl, geo, (long long)b * DSS_TPB + threadIdx.x, blockIdx.y);
      return;
    }
    b -= L.nb_quad;
  }
  if constexpr (PARTS & 4)
    dss_nodes_body<RSP, AVG>(L.nodes, L.nnodes, fl, geo, L.halo, (long long)b * DSS_TPB + threadIdx.x, blockIdx.y);
}

__global__ void scale_interior_kernel(FieldList fl, const double* __restrict__ geo, int nelem) {
  const int ie = blockIdx.x, f = blockIdx.y;
  if (ie >= nelem) return;
  double* fld = fl.base[f] + (size_t)ie * fl.estride[f];
  for (int i = threadIdx.x; i < 4 * NLEV; i += blockDim.x) {
    const int ip = i / NLEV, k = i % NLEV;
    const int p = ip == 0 ? 5 : ip == 1 ? 6 : ip == 2 ? 9 : 10;
    fld[p * NLEV + k] *= geo[((size_t)ie * NPSQ + p) * GEO_N + G_RSPHEREMP];
  }
}

// send buffer [pt][field][lev]
__global__ void halo_pack_kernel(const int* __restrict__ send_src, int npts, FieldList fl, double* __restrict__ buf) {
  const int i = blockIdx.x, f = blockIdx.y;
  const int s = send_src[i];
  const double* src = fl.base[f] + (size_t)(s >> 4) * fl.estride[f] + (s & 15) * NLEV;
  double* dst = buf + ((size_t)i * fl.nf + f) * NLEV;
  for (int k = threadIdx.x; k < NLEV; k += blockDim.x) dst[k] = src[k];
}

// ---- P2P halo over NVLink -----------------------------------------------------------------------
// Every rank maps its neighbours' receive buffers (CUDA IPC) and the pack kernels store the boundary
// points straight into them: no send buffer, no NCCL launch. An exchange is numbered (epoch, the same
// sequence on every rank) and uses receive buffer epoch & 1; after the pack a one-warp kernel publishes
// the epoch in the neighbours' flag words (fence + release at system scope), and the consumer stream
// runs a one-warp kernel that spins on this rank's own flags (acquire at system scope) before the kernel
// that reads the halo. Two buffers suffice: a rank can pack exchange n only after it consumed exchange
// n - 1 from each neighbour, and the neighbour packed n - 1 after it had consumed n - 2.
constexpr int MAX_PEERS = 16;
struct HaloPeers {
  double* recv[2][MAX_PEERS];  // the peers' receive buffers (mapped)
  int* flags[MAX_PEERS];       // the peers' flag arrays (mapped), indexed by sender rank
  int rank_of[MAX_PEERS];
  int npeers, my_rank;
};
static HaloPeers g_peers;

__global__ void halo_pack_p2p_kernel(const int* __restrict__ send_src, const int* __restrict__ dst_pt,
                                     const int* __restrict__ dst_peer, FieldList fl, HaloPeers P, int sel) {
  const int i = blockIdx.x, f = blockIdx.y;
  const int s = send_src[i];
  const double* src = fl.base[f] + (size_t)(s >> 4) * fl.estride[f] + (s & 15) * NLEV;
  double* dst = P.recv[sel][dst_peer[i]] + ((size_t)dst_pt[i] * fl.nf + f) * NLEV;
  for (int k = threadIdx.x; k < NLEV; k += blockDim.x) dst[k] = src[k];
}
__global__ void minmax_pack_p2p_kernel(const int* __restrict__ send_elem, const int* __restrict__ dst_conn,
                                       const int* __restrict__ dst_peer, const double* __restrict__ qlim, int qsize,
                                       HaloPeers P, int sel) {
  const int i = blockIdx.x, q = blockIdx.y;
  const double* src = qlim + ((size_t)send_elem[i] * QSIZE_D + q) * 2 * NLEV;
  double* dst = P.recv[sel][dst_peer[i]] + ((size_t)dst_conn[i] * qsize + q) * 2 * NLEV;
  for (int k = threadIdx.x; k < 2 * NLEV; k += blockDim.x) dst[k] = src[k];
}
// launched behind the pack kernel on the same stream: its stores are complete; make them visible, then publish
__global__ void halo_signal_kernel(HaloPeers P, int epoch) {
  __threadfence_system();
  if (threadIdx.x < P.npeers) {
    int* f = P.flags[threadIdx.x] + P.my_rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
}
__global__ void halo_wait_kernel(const int* flags, HaloPeers P, int epoch) {
  if (threadIdx.x < P.npeers) {
    const int* f = flags + P.rank_of[threadIdx.x];
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while (v - epoch < 0);
  }
  __threadfence_system();
}

// min/max: qlim [ie][QSIZE_D][2][NLEV]; halo slot [conn][qsize][2][NLEV]
__global__ void minmax_pack_kernel(const int* __restrict__ send_elem, const double* __restrict__ qlim, int qsize,
                                   double* __restrict__ buf) {
  const int i = blockIdx.x, q = blockIdx.y;
  const double* src = qlim + ((size_t)send_elem[i] * QSIZE_D + q) * 2 * NLEV;
  double* dst = buf + ((size_t)i * qsize + q) * 2 * NLEV;
  for (int k = threadIdx.x; k < 2 * NLEV; k += blockDim.x) dst[k] = src[k];
}

// elist = the elements to process (null: all, in order): elements without an off-rank neighbour
// go first, overlapping the halo exchange the others wait for.
// A block takes MM_E consecutive elements of the list — neighbours on the space-filling curve, so most of the
// eight neighbours' rows a thread reads are rows other threads of the block read too — and one tracer: the
// pass reads every qlim row nine times, and it is the L1 that absorbs the repeats instead of the L2.
#ifndef HXX_MM_E
#define HXX_MM_E 4
#endif
constexpr int MM_E = HXX_MM_E;
__global__ void __launch_bounds__(MM_E* NLEV)
    minmax_kernel(const int* __restrict__ nbr8, const double* qin, double* __restrict__ qout, const double* halo,
                  int qsize, const int* __restrict__ elist, int nlist) {
  const int el = blockIdx.x * MM_E + threadIdx.x / NLEV, k = threadIdx.x % NLEV, q = blockIdx.y;
  if (el >= nlist) return;
  const int ie = elist ? elist[el] : el;
  const double* mine = qin + ((size_t)ie * QSIZE_D + q) * 2 * NLEV;
  double mn = mine[k], mx = mine[NLEV + k];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int nb = nbr8[ie * 8 + c];
    if (nb == DSS_NONE) continue;
    const double* th = nb >= 0 ? qin + ((size_t)nb * QSIZE_D + q) * 2 * NLEV
                               : halo + ((size_t)(~nb) * qsize + q) * 2 * NLEV;
    mn = fmin(mn, th[k]);
    mx = fmax(mx, th[NLEV + k]);
  }
  double* o = qout + ((size_t)ie * QSIZE_D + q) * 2 * NLEV;
  o[k] = mn;
  o[NLEV + k] = mx;
}

// ---- plan ---------------------------------------------------------------------------------
namespace {
struct UF {
  std::vector<int> parent;
  int find(int x) { while (parent[x] != x) x = parent[x] = parent[parent[x]]; return x; }
  void unite(int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
};
}  // namespace

void free_exchange_plan() {
  if (S.nodes) { cudaFree(S.nodes); S.nodes = nullptr; }
  if (S.dss_pairs) { cudaFree(S.dss_pairs); S.dss_pairs = nullptr; }
  if (S.dss_quads) { cudaFree(S.dss_quads); S.dss_quads = nullptr; }

  S.npairs = S.nquads = 0;
  if (S.nbr8) { cudaFree(S.nbr8); S.nbr8 = nullptr; }
  if (S.elem_order) { cudaFree(S.elem_order); S.elem_order = nullptr; }
  if (S.send_src) { cudaFree(S.send_src); S.send_src = nullptr; }
  if (S.send_conn_elem) { cudaFree(S.send_conn_elem); S.send_conn_elem = nullptr; }
  if (S.sendbuf) { cudaFree(S.sendbuf); S.sendbuf = nullptr; }
  if (S.recvbuf) { cudaFree(S.recvbuf); S.recvbuf = nullptr; }
  for (void* m : S.peer_alloc)
    if (m) cudaIpcCloseMemHandle(m);
  S.peer_alloc.clear();
  if (S.halo_alloc) { cudaFree(S.halo_alloc); S.halo_alloc = nullptr; }
  S.halo_recv[0] = S.halo_recv[1] = nullptr; S.halo_flags = nullptr; S.halo_buf_doubles = 0;
  if (S.send_pt_dst) { cudaFree(S.send_pt_dst); S.send_pt_dst = nullptr; }
  if (S.send_pt_peer) { cudaFree(S.send_pt_peer); S.send_pt_peer = nullptr; }
  if (S.send_conn_dst) { cudaFree(S.send_conn_dst); S.send_conn_dst = nullptr; }
  if (S.send_conn_peer) { cudaFree(S.send_conn_peer); S.send_conn_peer = nullptr; }
  S.p2p = false; S.halo_epoch = 0;
  S.nnodes = 0; S.n_halo_pts = S.n_send_pts = S.n_halo_conn = S.n_send_conn = 0;
  // every per-peer table: build_exchange_plan() appends, and halo_sendrecv() indexes them by peer slot
  S.peer.clear();
  S.peer_send_off.clear(); S.peer_send_cnt.clear(); S.peer_recv_off.clear(); S.peer_recv_cnt.clear();
  S.peer_csend_off.clear(); S.peer_csend_cnt.clear(); S.peer_crecv_off.clear(); S.peer_crecv_cnt.clear();
}

// Receive buffers of a multi-rank session. Collective over the session's ranks (every rank calls it from
// init_boundary_exchanges_c): the IPC handle of this rank's [recv 0 | recv 1 | flags] allocation and the
// offsets at which it expects each sender's points are all-gathered, and every neighbour's allocation is
// mapped. HXX_HALO=nccl (or a failed mapping on any rank) keeps the staged path: pack -> grouped
// ncclSend/ncclRecv -> receive buffer 0.
static void setup_halo_buffers(size_t n_send_pts, size_t n_send_conn) {
#ifdef HXX_WITH_NCCL
  ncclComm_t comm = (ncclComm_t)S.nccl;
  if (!comm) runtime_abort("halo exchange: NCCL communicator not initialised", 13);
  const int R = S.nranks;
  if ((int)S.peer.size() > MAX_PEERS) runtime_abort("halo exchange: more neighbour ranks than MAX_PEERS", 13);
  // this rank's needs, and the largest over the ranks (buffers are symmetric per pair, sized per rank)
  const size_t nd = std::max<size_t>(
      1, std::max((size_t)std::max(S.n_send_pts, S.n_halo_pts) * MAX_DSS_FIELDS * NLEV,
                  (size_t)std::max(S.n_send_conn, S.n_halo_conn) * QSIZE_D * 2 * NLEV));
  S.halo_buf_doubles = nd;
  const size_t flag_bytes = ((size_t)R * sizeof(int) + 255) / 256 * 256;
  CUDA_OK(cudaMalloc(&S.halo_alloc, 2 * nd * sizeof(double) + flag_bytes));
  CUDA_OK(cudaMemset(S.halo_alloc, 0, 2 * nd * sizeof(double) + flag_bytes));
  S.halo_recv[0] = (double*)S.halo_alloc;
  S.halo_recv[1] = S.halo_recv[0] + nd;
  S.halo_flags = (int*)(S.halo_recv[1] + nd);
  S.recvbuf = nullptr;
  CUDA_OK(cudaMalloc(&S.sendbuf, nd * sizeof(double)));  // staged (NCCL) path only
  // record = [IPC handle (64 B) | nd (8 B) | want_p2p (8 B) | recv_off[R] | crecv_off[R]] (ints, -1 = not a neighbour)
  const size_t rec = 64 + 16 + 2 * (size_t)R * sizeof(int);
  std::vector<unsigned char> mine(rec, 0), all(rec * R, 0);
  cudaIpcMemHandle_t hnd;
  const char* env = std::getenv("HXX_HALO");
  long long want = !(env && !std::strcmp(env, "nccl"));
  if (want && cudaIpcGetMemHandle(&hnd, S.halo_alloc) != cudaSuccess) { want = 0; cudaGetLastError(); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (want) std::memcpy(mine.data(), &hnd, 64);
  const long long nd_ll = (long long)nd;
  std::memcpy(mine.data() + 64, &nd_ll, 8);
  std::memcpy(mine.data() + 72, &want, 8);
  int* offs = reinterpret_cast<int*>(mine.data() + 80);
  for (int r = 0; r < 2 * R; ++r) offs[r] = -1;
  for (size_t j = 0; j < S.peer.size(); ++j) {
    offs[S.peer[j]] = S.peer_recv_off[j];
    offs[R + S.peer[j]] = S.peer_crecv_off[j];
  }
  unsigned char* d_all = nullptr;
  CUDA_OK(cudaMalloc(&d_all, rec * R));
  CUDA_OK(cudaMemcpy(d_all + rec * S.rank, mine.data(), rec, cudaMemcpyHostToDevice));
  if (ncclAllGather(d_all + rec * S.rank, d_all, rec, ncclChar, comm, S.stream) != ncclSuccess)
    runtime_abort("halo exchange: ncclAllGather of the IPC handles failed", 1);
  CUDA_OK(cudaStreamSynchronize(S.stream));
  CUDA_OK(cudaMemcpy(all.data(), d_all, rec * R, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaFree(d_all));
  bool p2p = true;
  for (int r = 0; r < R; ++r) {
    long long w;
    std::memcpy(&w, all.data() + rec * r + 72, 8);
    p2p = p2p && w;
  }
  HaloPeers& P = g_peers;
  P = HaloPeers{};
  P.npeers = (int)S.peer.size();
  P.my_rank = S.rank;
  std::vector<int> dst_off(S.peer.size(), 0), cdst_off(S.peer.size(), 0);
  S.peer_alloc.assign(S.peer.size(), nullptr);
  for (size_t j = 0; j < S.peer.size() && p2p; ++j) {
    const unsigned char* pr = all.data() + rec * S.peer[j];
    cudaIpcMemHandle_t ph;
    std::memcpy(&ph, pr, 64);
    long long pnd;
    std::memcpy(&pnd, pr + 64, 8);
    const int* poffs = reinterpret_cast<const int*>(pr + 80);
    void* base = nullptr;
    if (cudaIpcOpenMemHandle(&base, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      p2p = false;
      break;
    }
    S.peer_alloc[j] = base;
    P.recv[0][j] = (double*)base;
    P.recv[1][j] = (double*)base + pnd;
    P.flags[j] = (int*)((double*)base + 2 * pnd);
    P.rank_of[j] = S.peer[j];
    dst_off[j] = poffs[S.rank];
    cdst_off[j] = poffs[R + S.rank];
    if (dst_off[j] < 0 || cdst_off[j] < 0) runtime_abort("halo exchange: a neighbour rank does not list this rank", 13);
  }
  // a mapping that failed anywhere sends every rank down the staged path (one more tiny collective)
  {
    int* d_ok = nullptr;
    CUDA_OK(cudaMalloc(&d_ok, sizeof(int)));
    const int ok = p2p ? 1 : 0;
    CUDA_OK(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    if (ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, S.stream) != ncclSuccess)
      runtime_abort("halo exchange: ncclAllReduce failed", 1);
    CUDA_OK(cudaStreamSynchronize(S.stream));
    int okall = 0;
    CUDA_OK(cudaMemcpy(&okall, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaFree(d_ok));
    p2p = okall != 0;
  }
  S.p2p = p2p;
  if (S.rank == 0 && std::getenv("HXX_BANNER"))
    std::printf("HOMMEXX-B200 halo: %s\n", p2p ? "P2P stores over NVLink (CUDA IPC)" : "pack + ncclSend/ncclRecv");
  if (!p2p) {
    for (void*& m : S.peer_alloc)
      if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
    return;
  }
  // destination tables of the pack kernels
  std::vector<int> pt_dst(n_send_pts), pt_peer(n_send_pts), c_dst(n_send_conn), c_peer(n_send_conn);
  for (size_t j = 0; j < S.peer.size(); ++j) {
    for (int i = 0; i < S.peer_send_cnt[j]; ++i) {
      pt_dst[S.peer_send_off[j] + i] = dst_off[j] + i;
      pt_peer[S.peer_send_off[j] + i] = (int)j;
    }
    for (int i = 0; i < S.peer_csend_cnt[j]; ++i) {
      c_dst[S.peer_csend_off[j] + i] = cdst_off[j] + i;
      c_peer[S.peer_csend_off[j] + i] = (int)j;
    }
  }
  auto up = [](int*& d, const std::vector<int>& h) {
    CUDA_OK(cudaMalloc(&d, std::max<size_t>(1, h.size()) * sizeof(int)));
    if (!h.empty()) CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
  };
  up(S.send_pt_dst, pt_dst); up(S.send_pt_peer, pt_peer); up(S.send_conn_dst, c_dst); up(S.send_conn_peer, c_peer);
#else
  (void)n_send_pts; (void)n_send_conn;
  runtime_abort("halo exchange: built without NCCL", 12);
#endif
}

void build_exchange_plan() {
  free_exchange_plan();
  const int n = S.nelemd;
  // -- remote connections: send order (l_gid, l_pos), receive order (r_gid, r_pos), per peer
  struct RC { int peer, gid, pos, ie, c; };
  std::vector<RC> snd, rcv;
  for (int ie = 0; ie < n; ++ie)
    for (int c = 0; c < 8; ++c) {
      const ConnInfo& i = S.conn[(size_t)ie * 8 + c];
      if (i.kind == 2 || i.sharing != 1) continue;
      snd.push_back({i.remote_pid, i.l_gid, i.l_pos, ie, c});
      rcv.push_back({i.remote_pid, i.r_gid, i.r_pos, ie, c});
    }
  auto cmp = [](const RC& a, const RC& b) { return std::tie(a.peer, a.gid, a.pos) < std::tie(b.peer, b.gid, b.pos); };
  std::sort(snd.begin(), snd.end(), cmp);
  std::sort(rcv.begin(), rcv.end(), cmp);
  std::vector<int> send_src, send_conn_elem;
  std::map<int, std::array<int, 8>> per;  // peer -> {send_off,send_cnt,recv_off,recv_cnt, csend_off,csend_cnt,crecv_off,crecv_cnt}
  for (size_t j = 0; j < snd.size(); ++j) {
    const RC& r = snd[j];
    auto& a = per.try_emplace(r.peer, std::array<int, 8>{-1, 0, -1, 0, -1, 0, -1, 0}).first->second;
    if (a[0] < 0) { a[0] = (int)send_src.size(); a[4] = (int)j; }
    if (r.c < 4) for (int k = 0; k < 4; ++k) send_src.push_back(r.ie * 16 + EDGE_PTS[r.c][k]);
    else send_src.push_back(r.ie * 16 + CORNER_PTS[r.c - 4]);
    a[1] = (int)send_src.size() - a[0];
    a[5] = (int)j + 1 - a[4];
    send_conn_elem.push_back(r.ie);
  }
  std::unordered_map<long long, int> halo_of;  // remote (gid*16+pt) -> halo point index
  std::vector<int> halo_conn_of((size_t)n * 8, -1);
  int nh = 0;
  for (size_t j = 0; j < rcv.size(); ++j) {
    const RC& r = rcv[j];
    auto& a = per.try_emplace(r.peer, std::array<int, 8>{-1, 0, -1, 0, -1, 0, -1, 0}).first->second;
    if (a[2] < 0) { a[2] = nh; a[6] = (int)j; }
    if (r.pos < 4) for (int k = 0; k < 4; ++k) halo_of.emplace((long long)r.gid * 16 + EDGE_PTS[r.pos][k], nh++);
    else halo_of.emplace((long long)r.gid * 16 + CORNER_PTS[r.pos - 4], nh++);
    a[3] = nh - a[2];
    a[7] = (int)j + 1 - a[6];
    halo_conn_of[(size_t)r.ie * 8 + r.c] = (int)j;
  }
  // note: emplace keeps the first index of a repeated remote point, but nh advanced for every
  // slot point, so the buffer layout stays the plain concatenation of the slots.
  S.n_send_pts = (int)send_src.size(); S.n_halo_pts = nh;
  S.n_send_conn = (int)snd.size(); S.n_halo_conn = (int)rcv.size();
  for (auto& kv : per) {
    S.peer.push_back(kv.first);
    S.peer_send_off.push_back(kv.second[0]); S.peer_send_cnt.push_back(kv.second[1]);
    S.peer_recv_off.push_back(kv.second[2]); S.peer_recv_cnt.push_back(kv.second[3]);
    S.peer_csend_off.push_back(kv.second[4]); S.peer_csend_cnt.push_back(kv.second[5]);
    S.peer_crecv_off.push_back(kv.second[6]); S.peer_crecv_cnt.push_back(kv.second[7]);
    if (kv.second[1] != kv.second[3] || kv.second[5] != kv.second[7])
      runtime_abort("build_exchange_plan: asymmetric connections with a neighbour rank", 13);
  }

  // -- ordered contributions per local boundary point, and node grouping
  struct Src { long long key; int local; };  // local: lid*16+pt or -1
  std::vector<std::vector<Src>> contrib((size_t)n * 16);
  std::unordered_map<long long, int> id_of;  // member key -> dense id
  std::vector<long long> key_of;
  std::vector<int> local_of;
  auto member = [&](long long key, int local) {
    auto it = id_of.find(key);
    if (it != id_of.end()) { if (local >= 0) local_of[it->second] = local; return it->second; }
    const int id = (int)key_of.size();
    id_of.emplace(key, id); key_of.push_back(key); local_of.push_back(local);
    return id;
  };
  auto gid_of_local = [&](int ie) {
    for (int c = 0; c < 8; ++c) if (S.conn[(size_t)ie * 8 + c].kind != 2) return S.conn[(size_t)ie * 8 + c].l_gid;
    return -1 - ie;  // isolated element (unit-test sessions): private key space
  };
  UF uf;
  std::vector<int> self_id((size_t)n * 16, -1);
  auto key = [](int gid, int pt) { return (long long)gid * 16 + pt; };
  for (int ie = 0; ie < n; ++ie) {
    const int gid = gid_of_local(ie);
    auto add = [&](int pt, const ConnInfo& i, int rpt) {
      const bool loc = i.sharing == 0;
      contrib[(size_t)ie * 16 + pt].push_back({key(i.r_gid, rpt), loc ? i.r_lid * 16 + rpt : -1});
    };
    for (int k = 0; k < 4; ++k)
      for (int ed = 0; ed < 4; ++ed) {
        const ConnInfo& i = S.conn[(size_t)ie * 8 + ed];
        if (i.kind == 2) continue;
        add(EDGE_PTS[ed][k], i, EDGE_PTS[i.r_pos][i.direction ? 3 - k : k]);
      }
    for (int c = 0; c < 4; ++c) {
      const ConnInfo& i = S.conn[(size_t)ie * 8 + 4 + c];
      if (i.kind == 2) continue;
      add(CORNER_PTS[c], i, CORNER_PTS[i.r_pos - 4]);
    }
    for (int pt = 0; pt < 16; ++pt) {
      if (contrib[(size_t)ie * 16 + pt].empty()) continue;
      self_id[(size_t)ie * 16 + pt] = member(key(gid, pt), ie * 16 + pt);
    }
  }
  for (size_t lp = 0; lp < contrib.size(); ++lp)
    for (const Src& s : contrib[lp]) member(s.key, s.local);
  uf.parent.resize(key_of.size());
  for (size_t i = 0; i < uf.parent.size(); ++i) uf.parent[i] = (int)i;
  for (size_t lp = 0; lp < contrib.size(); ++lp)
    for (const Src& s : contrib[lp]) uf.unite(self_id[lp], id_of[s.key]);
  std::unordered_map<int, int> node_of_root;
  std::vector<DssNode> nodes;
  std::vector<std::vector<int>> node_members;
  for (size_t i = 0; i < key_of.size(); ++i) {
    const int r = uf.find((int)i);
    auto it = node_of_root.find(r);
    int nid;
    if (it == node_of_root.end()) {
      nid = (int)nodes.size();
      node_of_root.emplace(r, nid);
      DssNode nd;
      for (int m = 0; m < 4; ++m) { nd.src[m] = DSS_NONE; for (int t = 0; t < 3; ++t) nd.ord[m][t] = 255; }
      nd.nmem = 0; nd.pad[0] = nd.pad[1] = nd.pad[2] = 0;
      nodes.push_back(nd);
      node_members.emplace_back();
    } else nid = it->second;
    if (nodes[nid].nmem >= 4) runtime_abort("build_exchange_plan: a GLL node has more than 4 sharers", 13);
    const int m = nodes[nid].nmem++;
    node_members[nid].push_back((int)i);
    if (local_of[i] >= 0) nodes[nid].src[m] = local_of[i];
    else {
      auto h = halo_of.find(key_of[i]);
      if (h == halo_of.end()) runtime_abort("build_exchange_plan: remote GLL point without a halo slot", 13);
      nodes[nid].src[m] = ~h->second;
    }
  }
  for (size_t lp = 0; lp < contrib.size(); ++lp) {
    if (contrib[lp].empty()) continue;
    const int nid = node_of_root[uf.find(self_id[lp])];
    const auto& mem = node_members[nid];
    const int me = (int)(std::find(mem.begin(), mem.end(), self_id[lp]) - mem.begin());
    if (contrib[lp].size() > 3) runtime_abort("build_exchange_plan: more than 3 contributions to a point", 13);
    for (size_t t = 0; t < contrib[lp].size(); ++t) {
      const int id = id_of[contrib[lp][t].key];
      nodes[nid].ord[me][t] = (uint8_t)(std::find(mem.begin(), mem.end(), id) - mem.begin());
    }
  }
  // -- split into the lean pair / quad lists and the generic remainder
  std::vector<DssPair> pairs;
  std::vector<DssQuad> quads;
  std::vector<DssNode> rest;
  for (const DssNode& nd : nodes) {
    bool local = true;
    for (int m = 0; m < nd.nmem; ++m) local = local && nd.src[m] >= 0;
    if (local && nd.nmem == 2 && nd.ord[0][0] == 1 && nd.ord[0][1] == 255 && nd.ord[1][0] == 0 && nd.ord[1][1] == 255) {
      pairs.push_back({nd.src[0], nd.src[1]});
      continue;
    }
    if (local && nd.nmem == 4 && nd.ord[0][2] < 4) {
      // canonical order: C0 = member 0, C2 = its first (S/N) neighbour, C1 = its second (W/E), C3 = diagonal
      const int C[4] = {0, nd.ord[0][1], nd.ord[0][0], nd.ord[0][2]};
      int inv[4] = {-1, -1, -1, -1};
      bool okp = true;
      for (int i = 0; i < 4; ++i) { if (C[i] > 3 || inv[C[i]] >= 0) okp = false; else inv[C[i]] = i; }
      int swaps = 0;
      if (okp) {
        // expected (first, second, diag) of canonical members 1..3 and the swapped alternative
        static const int expect[4][3] = {{2, 1, 3}, {3, 0, 2}, {0, 3, 1}, {1, 2, 0}};
        for (int i = 1; i < 4 && okp; ++i) {
          const uint8_t* o = nd.ord[C[i]];
          if (o[0] > 3 || o[1] > 3 || o[2] > 3) { okp = false; break; }
          const int a0 = inv[o[0]], a1 = inv[o[1]], a2 = inv[o[2]];
          if (a2 != expect[i][2]) okp = false;
          else if (a0 == expect[i][0] && a1 == expect[i][1]) {}
          else if (a0 == expect[i][1] && a1 == expect[i][0]) swaps |= 1 << (i - 1);
          else okp = false;
        }
      }
      if (okp) {
        DssQuad q{};
        for (int i = 0; i < 4; ++i) q.m[i] = nd.src[C[i]];
        q.swaps = swaps;
        quads.push_back(q);
        continue;
      }
    }
    rest.push_back(nd);
  }
  S.nnodes = (int)rest.size();
  S.npairs = (int)pairs.size();
  S.nquads = (int)quads.size();
  if (S.nnodes) {
    CUDA_OK(cudaMalloc(&S.nodes, rest.size() * sizeof(DssNode)));
    CUDA_OK(cudaMemcpy(S.nodes, rest.data(), rest.size() * sizeof(DssNode), cudaMemcpyHostToDevice));
  }
  if (S.npairs) {
    CUDA_OK(cudaMalloc(&S.dss_pairs, pairs.size() * sizeof(DssPair)));
    CUDA_OK(cudaMemcpy(S.dss_pairs, pairs.data(), pairs.size() * sizeof(DssPair), cudaMemcpyHostToDevice));
  }
  if (S.nquads) {
    CUDA_OK(cudaMalloc(&S.dss_quads, quads.size() * sizeof(DssQuad)));
    CUDA_OK(cudaMemcpy(S.dss_quads, quads.data(), quads.size() * sizeof(DssQuad), cudaMemcpyHostToDevice));
  }

  // -- neighbour table for the min/max exchange
  std::vector<int> nbr8((size_t)n * 8, DSS_NONE);
  for (int ie = 0; ie < n; ++ie)
    for (int c = 0; c < 8; ++c) {
      const ConnInfo& i = S.conn[(size_t)ie * 8 + c];
      if (i.kind == 2) continue;
      nbr8[(size_t)ie * 8 + c] = i.sharing == 0 ? i.r_lid : ~halo_conn_of[(size_t)ie * 8 + c];
    }
  {
    // interior elements (no off-rank neighbour) first, then the ones that need the halo
    std::vector<int> order;
    order.reserve(n);
    for (int pass = 0; pass < 2; ++pass)
      for (int ie = 0; ie < n; ++ie) {
        bool needs = false;
        for (int c = 0; c < 8; ++c) needs = needs || (nbr8[(size_t)ie * 8 + c] < 0 && nbr8[(size_t)ie * 8 + c] != DSS_NONE);
        if (needs == (pass == 1)) order.push_back(ie);
      }
    S.n_interior = 0;
    for (int ie = 0; ie < n; ++ie) {
      bool needs = false;
      for (int c = 0; c < 8; ++c) needs = needs || (nbr8[(size_t)ie * 8 + c] < 0 && nbr8[(size_t)ie * 8 + c] != DSS_NONE);
      if (!needs) ++S.n_interior;
    }
    CUDA_OK(cudaMalloc(&S.elem_order, std::max<size_t>(1, order.size()) * sizeof(int)));
    if (n) CUDA_OK(cudaMemcpy(S.elem_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaMalloc(&S.nbr8, std::max<size_t>(1, nbr8.size()) * sizeof(int)));
  if (n) CUDA_OK(cudaMemcpy(S.nbr8, nbr8.data(), nbr8.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (S.n_send_pts) {
    CUDA_OK(cudaMalloc(&S.send_src, send_src.size() * sizeof(int)));
    CUDA_OK(cudaMemcpy(S.send_src, send_src.data(), send_src.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&S.send_conn_elem, send_conn_elem.size() * sizeof(int)));
    CUDA_OK(cudaMemcpy(S.send_conn_elem, send_conn_elem.data(), send_conn_elem.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (S.nranks > 1) setup_halo_buffers(send_src.size(), snd.size());
}

// grouped send/recv with every neighbour rank; offsets/counts in units of `unit` doubles
static void halo_sendrecv(const std::vector<int>& soff, const std::vector<int>& scnt, const std::vector<int>& roff,
                          const std::vector<int>& rcnt, size_t unit) {
#ifdef HXX_WITH_NCCL
  ncclComm_t comm = (ncclComm_t)S.nccl;
  if (!comm) runtime_abort("halo exchange: NCCL communicator not initialised", 13);
  bool ok = ncclGroupStart() == ncclSuccess;
  for (size_t i = 0; i < S.peer.size(); ++i) {
    ok = ok && ncclSend(S.sendbuf + (size_t)soff[i] * unit, (size_t)scnt[i] * unit, ncclDouble, S.peer[i], comm,
                        S.comm_stream) == ncclSuccess;
    ok = ok && ncclRecv(S.halo_recv[0] + (size_t)roff[i] * unit, (size_t)rcnt[i] * unit, ncclDouble, S.peer[i], comm,
                        S.comm_stream) == ncclSuccess;
  }
  if (ncclGroupEnd() != ncclSuccess || !ok) runtime_abort("halo exchange: NCCL send/recv failed", 1);
#else
  (void)soff; (void)scnt; (void)roff; (void)rcnt; (void)unit;
  runtime_abort("halo exchange: built without NCCL", 12);
#endif
}

// Overlap: the pack kernel and the NCCL transfers run on the session's communication stream as
// soon as the producer has finished, while the compute stream does the DSS of every node whose
// sharers are all on this rank (pair and quad lists — they never touch a packed point); only the
// generic kernel, which reads the receive buffer, waits for the halo.
// One halo exchange on the communication stream, behind everything issued so far on the compute stream.
// Returns the receive buffer the consumer must read once halo_arrived() has been issued on its stream.
template <class PackP2P, class PackStaged>
static const double* halo_start(PackP2P&& pack_p2p, PackStaged&& pack_staged) {
  CUDA_OK(cudaEventRecord(S.ev_produced, S.stream));
  CUDA_OK(cudaStreamWaitEvent(S.comm_stream, S.ev_produced, 0));
  ++S.halo_epoch;
  if (S.p2p) {
    const int sel = (int)(S.halo_epoch & 1u);
    pack_p2p(sel);
    KERNEL_LAUNCHED(K_HALO_PACK);
    halo_signal_kernel<<<1, 32, 0, S.comm_stream>>>(g_peers, (int)S.halo_epoch);
    KERNEL_LAUNCHED(K_HALO_PACK);
    return S.halo_recv[sel];
  }
  pack_staged();
  KERNEL_LAUNCHED(K_HALO_PACK);
  return S.halo_recv[0];
}
static void halo_arrived() {  // on the compute stream, before the first kernel that reads the halo
  if (S.p2p) {
    halo_wait_kernel<<<1, 32, 0, S.stream>>>(S.halo_flags, g_peers, (int)S.halo_epoch);
    KERNEL_LAUNCHED(K_HALO_PACK);
  } else {
    CUDA_OK(cudaStreamWaitEvent(S.stream, S.ev_halo, 0));
  }
}

void dss_exchange(const FieldList& fl, bool rspheremp) {
  const bool halo = S.n_send_pts > 0;
  const double* recv = nullptr;
  if (halo) {
    recv = halo_start(
        [&](int sel) {
          halo_pack_p2p_kernel<<<dim3(S.n_send_pts, fl.nf), 96, 0, S.comm_stream>>>(S.send_src, S.send_pt_dst,
                                                                                    S.send_pt_peer, fl, g_peers, sel);
        },
        [&] {
          halo_pack_kernel<<<dim3(S.n_send_pts, fl.nf), 96, 0, S.comm_stream>>>(S.send_src, S.n_send_pts, fl, S.sendbuf);
          halo_sendrecv(S.peer_send_off, S.peer_send_cnt, S.peer_recv_off, S.peer_recv_cnt, (size_t)fl.nf * NLEV);
          CUDA_OK(cudaEventRecord(S.ev_halo, S.comm_stream));
        });
  }
  const int ny = (fl.nf + DSS_FPB - 1) / DSS_FPB;
  const bool avg = fl.navg > 0;
  auto nblk = [](int n) { return (int)(((long long)n * DSS_NLV + DSS_TPB - 1) / DSS_TPB); };
  DssLists L{(const DssPair*)S.dss_pairs, S.npairs, nblk(S.npairs), (const DssQuad*)S.dss_quads, S.nquads, nblk(S.nquads),
             S.nodes, S.nnodes, recv};
#define HXX_DSS_LAUNCH(PARTS, NB)                                                                        \
  do {                                                                                                   \
    const int nb_ = (NB);                                                                                \
    if (nb_) {                                                                                           \
      const dim3 grid(nb_, ny);                                                                          \
      PROBE(K_DSS);                                                                                      \
      if (avg && rspheremp) dss_kernel<PARTS, true, true><<<grid, DSS_TPB, 0, S.stream>>>(L, fl, S.geo); \
      else if (avg) dss_kernel<PARTS, false, true><<<grid, DSS_TPB, 0, S.stream>>>(L, fl, S.geo);        \
      else if (rspheremp) dss_kernel<PARTS, true, false><<<grid, DSS_TPB, 0, S.stream>>>(L, fl, S.geo);  \
      else dss_kernel<PARTS, false, false><<<grid, DSS_TPB, 0, S.stream>>>(L, fl, S.geo);                \
      KERNEL_LAUNCHED(K_DSS);                                                                            \
    }                                                                                                    \
  } while (0)
  HXX_DSS_LAUNCH(1, L.nb_pair);
  if (!halo) {
    HXX_DSS_LAUNCH(6, L.nb_quad + nblk(L.nnodes));
  } else {
    HXX_DSS_LAUNCH(2, L.nb_quad);
    halo_arrived();
    HXX_DSS_LAUNCH(4, nblk(L.nnodes));
  }
#undef HXX_DSS_LAUNCH
}

// Asynchronous NCCL errors (a peer that died) surface here instead of as a hang; polled once per
// prim_run_subcycle_c, next to the remap's abort flag.
void check_comm_errors() {
#ifdef HXX_WITH_NCCL
  if (!S.nccl) return;
  ncclResult_t st = ncclSuccess;
  if (ncclCommGetAsyncError((ncclComm_t)S.nccl, &st) != ncclSuccess || (st != ncclSuccess && st != ncclInProgress))
    runtime_abort("halo exchange: asynchronous NCCL error (a peer rank failed?)", 1);
#endif
}

void scale_interior_rspheremp(const FieldList& fl) {
  if (!S.nelemd) return;
  PROBE(K_DSS);
  scale_interior_kernel<<<dim3(S.nelemd, fl.nf), 96, 0, S.stream>>>(fl, S.geo, S.nelemd);
  KERNEL_LAUNCHED(K_DSS);
}

void minmax_exchange() {
  const int nq = S.p.qsize;
  if (!S.nelemd || !nq) return;
  const bool halo = S.n_send_conn > 0;
  const double* recv = nullptr;
  if (halo) {
    recv = halo_start(
        [&](int sel) {
          minmax_pack_p2p_kernel<<<dim3(S.n_send_conn, nq), 96, 0, S.comm_stream>>>(
              S.send_conn_elem, S.send_conn_dst, S.send_conn_peer, S.qlim, nq, g_peers, sel);
        },
        [&] {
          minmax_pack_kernel<<<dim3(S.n_send_conn, nq), 96, 0, S.comm_stream>>>(S.send_conn_elem, S.qlim, nq, S.sendbuf);
          halo_sendrecv(S.peer_csend_off, S.peer_csend_cnt, S.peer_crecv_off, S.peer_crecv_cnt, (size_t)nq * 2 * NLEV);
          CUDA_OK(cudaEventRecord(S.ev_halo, S.comm_stream));
        });
  }
  PROBE(K_MINMAX);
  if (!halo) {
    minmax_kernel<<<dim3((S.nelemd + MM_E - 1) / MM_E, nq), MM_E * NLEV, 0, S.stream>>>(S.nbr8, S.qlim, S.qlim_x, nullptr, nq,
                                                                                       nullptr, S.nelemd);
  } else {
    // elements whose eight neighbours are on this rank first, the others once the halo has landed
    if (S.n_interior)
      minmax_kernel<<<dim3((S.n_interior + MM_E - 1) / MM_E, nq), MM_E * NLEV, 0, S.stream>>>(
          S.nbr8, S.qlim, S.qlim_x, recv, nq, S.elem_order, S.n_interior);
    halo_arrived();
    if (S.nelemd > S.n_interior)
      minmax_kernel<<<dim3((S.nelemd - S.n_interior + MM_E - 1) / MM_E, nq), MM_E * NLEV, 0, S.stream>>>(
          S.nbr8, S.qlim, S.qlim_x, recv, nq, S.elem_order + S.n_interior, S.nelemd - S.n_interior);
  }
  KERNEL_LAUNCHED(K_MINMAX);
  std::swap(S.qlim, S.qlim_x);
}

// ---- registered field sets (the reference's register_field calls) -------------------------
FieldList fields_caar(int tl) {  // CaarFunctorImpl.hpp:68-79
  FieldList f;
  f.nf = 4;
  f.base[0] = S.v + off_v(0, tl, 0); f.estride[0] = (long long)NTL * 2 * NLF;
  f.base[1] = S.v + off_v(0, tl, 1); f.estride[1] = (long long)NTL * 2 * NLF;
  f.base[2] = S.t + off_s(0, tl); f.estride[2] = (long long)NTL * NLF;
  f.base[3] = S.dp3d + off_s(0, tl); f.estride[3] = (long long)NTL * NLF;
  return f;
}
FieldList fields_hv() {  // HyperviscosityFunctorImpl.cpp:44-54
  FieldList f;
  f.nf = 4;
  f.base[0] = S.vtens; f.estride[0] = 2 * NLF;
  f.base[1] = S.vtens + NLF; f.estride[1] = 2 * NLF;
  f.base[2] = S.ttens; f.estride[2] = NLF;
  f.base[3] = S.dptens; f.estride[3] = NLF;
  return f;
}
double* dss_var(int dss_opt) {
  return dss_opt == DSS_ETA ? S.eta_dot_dpdn : dss_opt == DSS_OMEGA ? S.omega_p : S.divdp_proj;
}
FieldList fields_euler(int tq, int dss_opt, int tavg_n0_qdp) {  // EulerStepFunctorImpl.hpp:137-153
  FieldList f;
  if (tavg_n0_qdp >= 0) {
    f.navg = S.p.qsize;
    f.avg_delta = (long long)off_q(0, tavg_n0_qdp, 0) - (long long)off_q(0, tq, 0);
  }
  const int nq = S.p.qsize;
  f.nf = nq + 1;
  for (int q = 0; q < nq; ++q) { f.base[q] = S.qdp + off_q(0, tq, q); f.estride[q] = (long long)QNTL * QSIZE_D * NLF; }
  f.base[nq] = dss_var(dss_opt); f.estride[nq] = NLF;
  return f;
}
FieldList fields_qtens() {  // EulerStepFunctorImpl.hpp:155-161
  FieldList f;
  const int nq = S.p.qsize;
  f.nf = nq;
  for (int q = 0; q < nq; ++q) { f.base[q] = S.qtens_biharmonic + (size_t)q * NLF; f.estride[q] = (long long)QSIZE_D * NLF; }
  return f;
}

}  // namespace hxx
