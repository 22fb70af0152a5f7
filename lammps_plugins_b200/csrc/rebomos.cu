// b200md -- REBOMoS force path on sm_100a.
//
// Reference semantics: lammps-plugins USER-REBOMOS/pair_rebomos.cpp
//   REBO_neigh :281-352, FREBO :358-447, bondorder :571-847, FLJ :453-558,
//   gSpline/PijSpline/Sp pair_rebomos.h:68-211.
//
// Formulation (DESIGN.md "REBOMoS kernels"): the reference walks half-bonds (i,j) and
// scatters 3 forces per (bond,k) triplet.  Here the energy is regrouped by CENTER atom,
//   E = sum_i E_i,  E_i = sum_{j in R(i)} 1/2 [ VR_ij + p_ij VA_ij ],
// p_ij depends only on i and its REBO neighbors, so all forces of E_i land on i and R(i):
//   K3+K4 rebo_center  lane group / owned atom: ordered filter of the short row (REBO sub-list), N_i,
//                      p_im and prefactors, forces F_m of all terms in which m takes part, F_i = -sum F_m
//   K5  lj             8 lanes / owned atom over the directed LJ-window row, no atomics
//   K8  fdotr          sum_{nall} x (x) f for the many-body part
// FP64 throughout; neighbor-position gathers are one 32-byte sector (double4 {x,y,z,elem}).

#include "common.cuh"

#include <cmath>

#define TOL 1.0e-9
#define BLOCK 256

// ================================================================== device math
__device__ __forceinline__ int elem_of(const double4 &q) { return __double2int_rn(q.w); }

// keep a value loaded from the constant bank in a register: without this the compiler re-materialises
// `cond ? par.a[k] : par.a[k+1]` as an indexed constant load at every use
__device__ __forceinline__ double pin(double v)
{
  asm volatile("" : "+d"(v));
  return v;
}

// Sp cutoff (pair_rebomos.h:195-211): value and derivative
__device__ __forceinline__ double sp_switch(double r, double rmin, double rw, double &dS)
{
  if (r <= rmin) {    // t <= 0 (rw > 0): most bonds of a crystal; no division spent
    dS = 0.0;
    return 1.0;
  }
  const double t = (r - rmin) / rw;    // rw = rcmax - rcmin
  if (t <= 0.0) {
    dS = 0.0;
    return 1.0;
  }
  if (t >= 1.0) {
    dS = 0.0;
    return 0.0;
  }
  double s, c;
  sincospi(t, &s, &c);
  dS = (-0.5 * 3.14159265358979323846 * s) / rw;
  return 0.5 * (1.0 + c);
}

// G(cos) value only (pair_rebomos.h:68-167)
__device__ __forceinline__ double gspline_val(const RebomosDev &par, double c, int t)
{
  const double *b = par.b[t];
  double g = b[6];
  g = fma(g, c, b[5]);
  g = fma(g, c, b[4]);
  g = fma(g, c, b[3]);
  g = fma(g, c, b[2]);
  g = fma(g, c, b[1]);
  g = fma(g, c, b[0]);
  if (c >= 0.5) {
    const double *bg = par.bg[t];
    double gam = bg[6];
    gam = fma(gam, c, bg[5]);
    gam = fma(gam, c, bg[4]);
    gam = fma(gam, c, bg[3]);
    gam = fma(gam, c, bg[2]);
    gam = fma(gam, c, bg[1]);
    gam = fma(gam, c, bg[0]);
    const double psi = 0.5 * (1.0 - cospi(2.0 * (c - 0.5)));
    g = g + psi * (gam - g);
  }
  return g;
}

// G(cos) and dG/dcos
__device__ __forceinline__ double gspline(const RebomosDev &par, double c, int t, double &dgdc)
{
  const double *b = par.b[t];
  double g = b[6];
  double dg = 6.0 * b[6];
  g = fma(g, c, b[5]);
  dg = fma(dg, c, 5.0 * b[5]);
  g = fma(g, c, b[4]);
  dg = fma(dg, c, 4.0 * b[4]);
  g = fma(g, c, b[3]);
  dg = fma(dg, c, 3.0 * b[3]);
  g = fma(g, c, b[2]);
  dg = fma(dg, c, 2.0 * b[2]);
  g = fma(g, c, b[1]);
  dg = fma(dg, c, b[1]);
  g = fma(g, c, b[0]);
  if (c >= 0.5) {
    const double *bg = par.bg[t];
    double gam = bg[6];
    double dgam = 6.0 * bg[6];
    gam = fma(gam, c, bg[5]);
    dgam = fma(dgam, c, 5.0 * bg[5]);
    gam = fma(gam, c, bg[4]);
    dgam = fma(dgam, c, 4.0 * bg[4]);
    gam = fma(gam, c, bg[3]);
    dgam = fma(dgam, c, 3.0 * bg[3]);
    gam = fma(gam, c, bg[2]);
    dgam = fma(dgam, c, 2.0 * bg[2]);
    gam = fma(gam, c, bg[1]);
    dgam = fma(dgam, c, bg[1]);
    gam = fma(gam, c, bg[0]);
    double sn, cs;
    sincospi(2.0 * (c - 0.5), &sn, &cs);
    const double psi = 0.5 * (1.0 - cs);
    const double dpsi = 3.14159265358979323846 * sn;
    dgdc = dg + dpsi * (gam - g) + psi * (dgam - dg);
    return g + psi * (gam - g);
  }
  dgdc = dg;
  return g;
}

// cen_scan[t] packs the list positions of atom-index threshold t for both center lists: Mo-like count in the low
// 30 bits, S-like count above
#define CEN_SHIFT 30
#define CEN_MASK ((1LL << CEN_SHIFT) - 1)

// ================================================================== staging kernels
__global__ void __launch_bounds__(BLOCK) pack_xq_kernel(const double *__restrict__ x,
                                                        const int *__restrict__ type,
                                                        const int *__restrict__ map, int ntypes, int lo, int hi,
                                                        double4 *__restrict__ xq, int *__restrict__ flags)
{
  int i = lo + blockIdx.x * BLOCK + threadIdx.x;
  if (i >= hi) return;
  int t = type[i];
  int e = -1;
  if (t >= 1 && t <= ntypes) e = map[t];
  else flags[3] = 1;    // invalid atom type
  xq[i] = make_double4(x[3 * (size_t) i], x[3 * (size_t) i + 1], x[3 * (size_t) i + 2], (double) e);
}

// flags[1] = 1: an atom moved more than margin/2 since the inner (wide) rows were derived.  Plugin mode with tight rows
// (xhold_t != NULL): flags[2] = 2 when an atom moved more than margin_t/2 since the tight rows were derived (they may
// miss a pair now), 1 when more than soon_sq (they are re-derived after this call, before they can miss one)
__global__ void __launch_bounds__(BLOCK) check_disp_kernel(const double4 *__restrict__ xq,
                                                           const double4 *__restrict__ xhold, int nall,
                                                           double thresh_sq, int *__restrict__ flags,
                                                           const double4 *__restrict__ xhold_t = nullptr,
                                                           double tight_sq = 0.0, double soon_sq = 0.0)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nall) return;
  double4 a = xq[i], b = xhold[i];
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  if (dx * dx + dy * dy + dz * dz > thresh_sq) flags[1] = 1;
  if (xhold_t) {
    b = xhold_t[i];
    dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    const double d = dx * dx + dy * dy + dz * dz;
    if (d > tight_sq) atomicMax(&flags[2], 2);
    else if (d > soon_sq) atomicMax(&flags[2], 1);
  }
}

// inner lists from the master rows: one warp per row, order-preserving ballot compaction
__global__ void __launch_bounds__(BLOCK) build_inner_kernel(
    const __grid_constant__ RebomosDev par, const double4 *__restrict__ xq,
    const int64_t *__restrict__ list_off, const int *__restrict__ list_num,
    const int *__restrict__ list_val, int rows, int inum, int *__restrict__ short_idx,
    int *__restrict__ short_num, const int64_t *__restrict__ lj_off, int *__restrict__ lj_num,
    int *__restrict__ lj_val, int *__restrict__ flags)
{
  const int lane = threadIdx.x & 31;
  const int i = (int) (((size_t) blockIdx.x * BLOCK + threadIdx.x) >> 5);
  if (i >= rows) return;
  const double4 xi = xq[i];
  const int ti = elem_of(xi);
  const int n = list_num[i];
  const int64_t base = list_off[i];
  const int64_t ljbase = (i < inum) ? lj_off[i] : 0;
  const int ljcap = (i < inum) ? (int) (lj_off[i + 1] - lj_off[i]) : 0;
  const unsigned lt = (1u << lane) - 1u;
  int ns = 0, nlA = 0, nlB = 0;
  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    bool ps = false, pl = false;
    int j = 0, tj = 0;
    if (e < n) {
      j = ld_stream_int(list_val + base + e) & B200MD_NEIGHMASK;
      const double4 xj = xq[j];
      tj = elem_of(xj);
      if (ti >= 0 && tj >= 0) {
        const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        const double rsq = dx * dx + dy * dy + dz * dz;
        const int pt = ti * 2 + tj;
        ps = rsq <= par.shortsq[pt];
        pl = (i < inum) && rsq <= par.ljsq[pt];
      }
    }
    const unsigned ms = __ballot_sync(0xffffffffu, ps);
    // LJ rows are segmented by the partner's element so that lj_kernel runs each segment with compile-time
    // pair constants: Mo partners fill the row's slots from the front, S partners from the back
    const unsigned mA = __ballot_sync(0xffffffffu, pl && tj == 0);
    const unsigned mB = __ballot_sync(0xffffffffu, pl && tj == 1);
    if (ps) {
      const int pos = ns + __popc(ms & lt);
      if (pos < B200MD_SHORT_WIDTH) short_idx[(size_t) i * B200MD_SHORT_WIDTH + pos] = j;
      else flags[0] = 1;
    }
    if (pl) {
      if (tj == 0) lj_val[ljbase + nlA + __popc(mA & lt)] = j;
      else lj_val[ljbase + ljcap - 1 - (nlB + __popc(mB & lt))] = j;
    }
    ns += __popc(ms);
    nlA += __popc(mA);
    nlB += __popc(mB);
  }
  if (lane == 0) {
    short_num[i] = min(ns, B200MD_SHORT_WIDTH);
    if (i < inum) {
      lj_num[2 * i] = nlA;
      lj_num[2 * i + 1] = nlB;
    }
    atomicAdd(&flags[4], ns);    // statistics (low contention: one per row, only at rebuilds)
    atomicAdd(&flags[5], nlA + nlB);
  }
}

// ================================================================== REBO rows (parity API)
// parity/debug variant: REBO rows for owned AND ghost atoms written to caller-visible arrays
__global__ void __launch_bounds__(BLOCK) rebo_rows_kernel(
    const __grid_constant__ RebomosDev par, const double4 *__restrict__ xq,
    const int *__restrict__ short_idx, const int *__restrict__ short_num, int nrows,
    int stride, int *__restrict__ out_num, int *__restrict__ out_rows, double *__restrict__ nM_out,
    double *__restrict__ nS_out, int *__restrict__ flags)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nrows) return;
  const double4 xi = xq[i];
  const int ti = elem_of(xi);
  const int n = (ti >= 0) ? short_num[i] : 0;
  int nb = 0;
  double nM = 0.0, nS = 0.0;
  for (int e = 0; e < n; e++) {
    const int j = short_idx[(size_t) i * B200MD_SHORT_WIDTH + e];
    const double4 xj = xq[j];
    const int tj = elem_of(xj);
    const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const int pt = ti * 2 + tj;
    if (rsq < par.rcmaxsq[pt]) {
      if (nb < stride) out_rows[(size_t) i * stride + nb] = j;
      else flags[0] = 1;
      nb++;
      double dS;
      const double w = sp_switch(sqrt(rsq), par.rcmin[pt], par.rcw[pt], dS);
      if (tj == 0) nM += w;
      else nS += w;
    }
  }
  out_num[i] = nb;
  nM_out[i] = nM;
  nS_out[i] = nS;
}

// ================================================================== K3+K4 fused: one lane GROUP per center
// REBO_neigh + FREBO + bondorder for one owned center i entirely inside a group of G lanes:
//   A  the group scans i's short row (G candidates per trip), keeps rsq < rcmax^2 with the reference's
//      operation order, and stages the bonds in row order in shared memory {d, 1/r, w, w', j, elem};
//      N_i = nM + nS by a group reduction, P(N), P'(N)
//   P  the nb(nb-1)/2 UNORDERED bond pairs are dealt to the lanes: cos_mq, G(cos), G'(cos) once per pair into
//      a triangular shared table.  The reference evaluates gSpline four times per pair (twice per ordered
//      pair, pair_rebomos.cpp:611-622 and :639-667); cos and G are symmetric in (m, q).
//   B  lane m: S_m = sum_{q != m} w_q G_mq, p_m, VR/VA, radial coefficient, prefactor dE/dS_m, energy
//   C  lane m: force on neighbor j_m from every term of E_i it appears in; one FP64 atomic triple per bond,
//      -sum_m F_m to the center by a group reduction
// Nothing but the forces goes through global memory (v1 wrote and re-read a 500 B/atom bond table between
// three kernels).  Centers are launched by element (Mo: G = 16, S: G = 4), which makes every group of a warp
// run the same trip counts AND turns every spline / P(N) coefficient into an immediate constant-bank
// operand (ELEM is a template argument; indexing the parameter bank with a runtime element costs one LDC
// per Horner step: ncu r01 pipe_adu 39 %).  A center with more bonds than its class' staging capacity is
// deferred to an overflow list handled by a G = 16, CAP = 16 launch of the same element.
// DET: per-bond forces go to a (center, slot) table instead of atomics; rebo_gather_kernel sums them by
// destination in a fixed order (deterministic mode).
struct DetTables {
  double *fb;    // [ncen * MAX_REBO * 3] force of bond slot m on its neighbor
  double *fi;    // [ncen * 3]            force on the center
  int *j;        // [ncen * MAX_REBO]     neighbor of bond slot m
  int *nb;       // [ncen]
};

// G(cos) and dG/dcos for a compile-time element: the base polynomial (pair_rebomos.h:104-130) ...
template <int ELEM>
__device__ __forceinline__ double gspline_base(const RebomosDev &par, double c, double &dgdc)
{
  const double *b = par.b[ELEM];
  double g = b[6];
  double dg = 6.0 * b[6];
  g = fma(g, c, b[5]);
  dg = fma(dg, c, 5.0 * b[5]);
  g = fma(g, c, b[4]);
  dg = fma(dg, c, 4.0 * b[4]);
  g = fma(g, c, b[3]);
  dg = fma(dg, c, 3.0 * b[3]);
  g = fma(g, c, b[2]);
  dg = fma(dg, c, 2.0 * b[2]);
  g = fma(g, c, b[1]);
  dg = fma(dg, c, b[1]);
  g = fma(g, c, b[0]);
  dgdc = dg;
  return g;
}
// ... and the blend towards the gamma polynomial for cos >= 1/2 (pair_rebomos.h:131-167), applied to (g, dg) in place
template <int ELEM>
__device__ __forceinline__ void gspline_blend(const RebomosDev &par, double c, double &g, double &dg)
{
  const double *bg = par.bg[ELEM];
  double gam = bg[6];
  double dgam = 6.0 * bg[6];
  gam = fma(gam, c, bg[5]);
  dgam = fma(dgam, c, 5.0 * bg[5]);
  gam = fma(gam, c, bg[4]);
  dgam = fma(dgam, c, 4.0 * bg[4]);
  gam = fma(gam, c, bg[3]);
  dgam = fma(dgam, c, 3.0 * bg[3]);
  gam = fma(gam, c, bg[2]);
  dgam = fma(dgam, c, 2.0 * bg[2]);
  gam = fma(gam, c, bg[1]);
  dgam = fma(dgam, c, bg[1]);
  gam = fma(gam, c, bg[0]);
  double sn, cs;
  sincospi(2.0 * (c - 0.5), &sn, &cs);
  const double psi = 0.5 * (1.0 - cs);
  const double dpsi = 3.14159265358979323846 * sn;
  dg = dg + dpsi * (gam - g) + psi * (dgam - dg);
  g = g + psi * (gam - g);
}

template <int NT, int G, int CAP, int ELEM, bool EV, bool DET, bool ATOM, int MINB>
__global__ void __launch_bounds__(NT, MINB) rebo_center_kernel(
    const __grid_constant__ RebomosDev par, const double4 *__restrict__ xq, const int *__restrict__ short_idx,
    const int *__restrict__ short_num, const int *__restrict__ cen_list, const int *__restrict__ cen_count_ptr,
    const long long *__restrict__ cen_scan, int t_lo, int t_hi, int *__restrict__ ovf_list, int *__restrict__ ovf_count, double *__restrict__ f, const DetTables det,
    double *__restrict__ scal, int *__restrict__ flags, double *__restrict__ pa_e, double *__restrict__ pa_v)
{
  constexpr int NG = NT / G;    // groups per block
  // pair table in ROUND-ROBIN layout: the unordered bond pair {m, (m + k) mod nb}, 1 <= k <= nb/2, lives at [k-1][m].
  // Indexing is an add (v2 decoded a triangular index with sqrtf per pair and min/max/mul per use: 13 % of all
  // instructions, ncu r01 source view), lanes of a group touch consecutive words, and a lane that walks its partners
  // q = m+1 .. m+nb-1 (mod nb) never meets q == m.  For even nb the round k = nb/2 holds every pair twice.
  constexpr int NTRI = (CAP / 2) * CAP;
  constexpr unsigned GBITS = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
  constexpr int tb = ELEM * 2;
  // staged bonds: UNIT vector u = d / r (the angular terms need nothing else: cos = u_m . u_q, and every force of the
  // angular part is a combination of unit vectors), r, 1/r, switch w and w'
  __shared__ double s_dx[NG * CAP], s_dy[NG * CAP], s_dz[NG * CAP], s_ri[NG * CAP], s_r[NG * CAP], s_w[NG * CAP],
      s_dw[NG * CAP], s_pref[NG * CAP], s_frad[NG * CAP];
  __shared__ double s_c[NG * NTRI], s_g[NG * NTRI], s_dg[NG * NTRI];
  __shared__ int s_j[NG * CAP], s_tj[NG * CAP];
  __shared__ unsigned char s_bl[NG * NTRI];    // table slots whose cos >= 1/2 (blend towards the gamma polynomial)
  static_assert(NTRI <= 256, "blend list stores table slots in bytes");
  const int lane = threadIdx.x & 31;
  const int sub = threadIdx.x & (G - 1);
  const int gl = threadIdx.x / G;    // group within the block
  const int gshift = lane & ~(G - 1);
  // Every vote, shuffle and warp barrier below runs on the FULL warp: the loops that contain them take the warp's
  // largest trip count (groups with less work are predicated off).  With per-group masks each of them was a WARPSYNC +
  // vote pair on a run-time mask: 8 % of the instructions and 17 % of the stall samples of the S-center launch sat on
  // the ballot of phase A (ncu source view, r02).
  constexpr unsigned FULL = 0xffffffffu;
  const int sb = gl * CAP;     // this group's staging base
  const int st = gl * NTRI;    // ... and pair-table base
  // centers [first, count) of the list: all of it (count read from the device), or -- plugin-mode pipelining -- the
  // piece of the ascending list whose atom indices lie in [t_lo, t_hi)
  int first = 0, count;
  if (cen_scan) {
    const long long c0 = cen_scan[t_lo], c1 = cen_scan[t_hi];
    first = (ELEM == 0) ? (int) (c0 & CEN_MASK) : (int) (c0 >> CEN_SHIFT);
    count = (ELEM == 0) ? (int) (c1 & CEN_MASK) : (int) (c1 >> CEN_SHIFT);
  } else
    count = *cen_count_ptr;
  double eacc[1] = {0.0};
  for (int g0 = first + blockIdx.x * NG + (gl & ~(32 / G - 1)); g0 < count; g0 += gridDim.x * NG) {    // warp-uniform
    const int g = g0 + (gl & (32 / G - 1));
    const bool valid = g < count;
    const int i = valid ? cen_list[g] : 0;
    const double4 xi = xq[i];
    const int n = valid ? short_num[i] : 0;
    const int nmax = __reduce_max_sync(FULL, n);
    // ---- A: ordered REBO sub-list (pair_rebomos.cpp:328-343)
    int nb = 0;
    double nM = 0.0, nS = 0.0;
    const int *row = short_idx + (size_t) i * B200MD_SHORT_WIDTH;
    // UB trips' worth of candidates are gathered before the first one is looked at (UB gathers in flight per lane:
    // ncu r01 v6 showed the S-center launch waiting on one gather per trip, long-scoreboard 3.6 per issue); they are then
    // consumed in row order, so membership and bond order stay those of the reference
    constexpr int UB = (G <= 4) ? 4 : 2;
    for (int e0 = 0; e0 < nmax; e0 += G * UB) {
      int jb[UB];
      double4 xb[UB];
#pragma unroll
      for (int u = 0; u < UB; u++) {
        const int e = e0 + u * G + sub;
        jb[u] = (e < n) ? row[e] : -1;
      }
#pragma unroll
      for (int u = 0; u < UB; u++) {
        xb[u] = make_double4(0.0, 0.0, 0.0, 0.0);
        if (jb[u] >= 0) xb[u] = ld_sector(xq + jb[u]);
      }
#pragma unroll
      for (int u = 0; u < UB; u++) {
        if (e0 + u * G >= nmax) break;    // warp-uniform
        bool in = false;
        const int j = jb[u];
        int tj = 0;
        double dx = 0, dy = 0, dz = 0, rsq = 1.0;
        if (j >= 0) {
          const double4 xj = xb[u];
          tj = elem_of(xj);
          dx = xi.x - xj.x;
          dy = xi.y - xj.y;
          dz = xi.z - xj.z;
          // same operation order as the reference, no FMA contraction: membership must be bit-exact
          rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
          in = rsq < (tj ? par.rcmaxsq[tb + 1] : par.rcmaxsq[tb]);
        }
        const unsigned bits = (__ballot_sync(FULL, in) >> gshift) & GBITS;
        if (in) {
          const int pos = nb + __popc(bits & ((1u << sub) - 1u));
          const double rinv = rsqrt_nr(rsq);
          const double r = rsq * rinv;
          double dw;
          const double w = sp_switch(r, tj ? par.rcmin[tb + 1] : par.rcmin[tb], tj ? par.rcw[tb + 1] : par.rcw[tb], dw);
          if (tj == 0) nM += w;
          else nS += w;
          if (pos < CAP) {
            s_dx[sb + pos] = dx * rinv;
            s_dy[sb + pos] = dy * rinv;
            s_dz[sb + pos] = dz * rinv;
            s_ri[sb + pos] = rinv;
            s_r[sb + pos] = r;
            s_w[sb + pos] = w;
            s_dw[sb + pos] = dw;
            s_j[sb + pos] = j;
            s_tj[sb + pos] = tj;
          }
        }
        nb += __popc(bits);
      }
    }
    __syncwarp();
    bool deferred = false;
    if (nb > CAP) {
      // more bonds than this launch class stages: hand the center to the wide launch, or fail at the cap
      if (sub == 0) {
        if (CAP < B200MD_MAX_REBO) ovf_list[atomicAdd(ovf_count, 1)] = i;
        else flags[0] = 1;
      }
      nb = 0;    // nothing more happens for this center in this launch (no `continue`: the warp stays together)
      deferred = CAP < B200MD_MAX_REBO;
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      nM += __shfl_xor_sync(FULL, nM, o);
      nS += __shfl_xor_sync(FULL, nS, o);
    }
    // PijSpline (pair_rebomos.h:173-179); N_i includes j (pair_rebomos.cpp:596-599)
    const double N = nM + nS;
    const double ex = exp(-par.a[ELEM][2] * N);
    const double dP = -par.a[ELEM][0] + par.a[ELEM][1] * par.a[ELEM][2] * ex;
    const double P = -par.a[ELEM][0] * (N - 1.0) - par.a[ELEM][1] * ex + par.a[ELEM][3];
    // ---- P: cos, G, G' of every unordered bond pair (round k: lane m takes the pair {m, m+k mod nb})
    const int nr = nb >> 1;
    int nbl = 0;
    // The nb (nb - 1) / 2 unordered pairs are dealt to ALL G lanes as one flat sequence p = (k - 1) nb + m: full rounds
    // k < nb/2 hold nb pairs each, the last round of an even nb only its first nb/2 (the other half are the same pairs
    // again: their table slots are filled by the lane that has the pair).  12 bonds on 16 lanes: 5 trips of 16 instead
    // of 6 trips of 12 with 6 pairs evaluated twice.
    const bool even = (nb & 1) == 0;
    const int npairs = (nb * (nb - 1)) >> 1;
    const unsigned inv_nb = nb > 0 ? (65536u + (unsigned) nb - 1u) / (unsigned) nb : 0u;    // p / nb for p < 256
    const int npmax = __reduce_max_sync(FULL, npairs);
    for (int p0 = 0; p0 < npmax; p0 += G) {
      const int p = p0 + sub;
      const bool act = p < npairs;
      const int k = act ? (int) (((unsigned) p * inv_nb) >> 16) + 1 : 1;
      const int m = act ? p - (k - 1) * nb : 0;
      int q = m + k;
      if (q >= nb) q -= nb;
      double c = 0.0;
      if (act) {
        c = s_dx[sb + m] * s_dx[sb + q] + s_dy[sb + m] * s_dy[sb + q] + s_dz[sb + m] * s_dz[sb + q];
        c = fmin(c, 1.0);
        c = fmax(c, -1.0);
      }
      double dg;
      const double gg = gspline_base<ELEM>(par, c, dg);
      const bool bl = act && c >= 0.5;
      const unsigned bits = (__ballot_sync(FULL, bl) >> gshift) & GBITS;
      const int t = (k - 1) * CAP + m;
      const bool twin = act && even && k == nr;    // the pair's second slot, [nr-1][q]
      if (act) {
        s_c[st + t] = c;
        s_g[st + t] = gg;
        s_dg[st + t] = dg;
        if (twin) {
          const int t2 = (k - 1) * CAP + q;
          s_c[st + t2] = c;
          s_g[st + t2] = gg;
          s_dg[st + t2] = dg;
        }
      }
      // the blend needs a second polynomial and a sincospi; only a few pairs of a center are in it (the 60-degree
      // Mo-Mo-Mo angles), so they are collected and fixed up in one extra trip instead of making every trip pay
      // (v2 took the blend path on a group vote: sincospi alone was 11 % of all instructions)
      if (bl) s_bl[st + nbl + __popc(bits & ((1u << sub) - 1u))] = (unsigned char) t;
      nbl += __popc(bits);
    }
    __syncwarp();
    for (int b0 = 0; b0 < nbl; b0 += G) {
      const int b = b0 + sub;
      if (b < nbl) {
        const int tl = s_bl[st + b];
        const int t = st + tl;
        double gg = s_g[t], dg = s_dg[t];
        gspline_blend<ELEM>(par, s_c[t], gg, dg);
        s_g[t] = gg;
        s_dg[t] = dg;
        const int kr = tl / CAP;    // round index k - 1 of the slot
        if (even && kr == nr - 1) {    // last round of an even nb: the twin slot [nr-1][q], q = m + nr
          const int t2 = st + kr * CAP + (tl - kr * CAP) + nr;
          s_g[t2] = gg;
          s_dg[t2] = dg;
        }
      }
    }
    __syncwarp();
    // ---- B: bond order and pair terms of bond m
    double ei_atom = 0.0;
    for (int m = sub; m < nb; m += G) {
      const double rinv = s_ri[sb + m];
      const double wm = s_w[sb + m], dwm = s_dw[sb + m];
      double pref = 0.0, frad = 0.0;
      if (wm > TOL) {
        double S = 0.0;
        // partners q = m+k (mod nb): rounds 1..nr are this lane's own table column, the rest are the partner's
        for (int k = 1; k <= nr; k++) {
          int q = m + k;
          if (q >= nb) q -= nb;
          S += s_w[sb + q] * s_g[st + (k - 1) * CAP + m];
        }
        for (int k = nr + 1; k < nb; k++) {
          int q = m + k;
          if (q >= nb) q -= nb;
          S += s_w[sb + q] * s_g[st + (nb - k - 1) * CAP + q];
        }
        const double p = rsqrt_nr(1.0 + S + P);
        const int pt = tb + s_tj[sb + m];
        const double r = s_r[sb + m];
        const double Q = par.Q[pt], al = par.alpha[pt];
        // VR = wm * VR0, VA = wm * VA0: the reference's VR / wm * dwm is VR0 * dwm without the division
        const double pre0 = par.A[pt] * exp(-al * r);
        const double VR0 = pre0 * (1.0 + Q * rinv);
        const double VR = wm * VR0;
        const double dVR = wm * pre0 * (-al - Q * rinv * rinv - Q * al * rinv) + VR0 * dwm;
        const double be = par.Beta[pt];
        const double VA0 = -par.BIJc[pt] * exp(-be * r);
        const double VA = wm * VA0;
        const double dVA = -be * VA + VA0 * dwm;
        pref = VA * 0.5 * (-0.5 * p * p * p);
        frad = 0.5 * (dVR + p * dVA) + pref * dP * dwm;    // multiplies the unit vector u_m
        if (EV) eacc[0] += 0.5 * (VR + p * VA);
        if (ATOM) {
          // ev_tally (pair_rebomos.cpp:443-444): the half-bond energy VR + bbar*VA goes half to i, half to j;
          // bbar = (p_ij + p_ji)/2, so each DIRECTED piece (VR + p_ij VA)/2 is split the same way
          const double q4 = 0.25 * (VR + p * VA);
          ei_atom += q4;
          atomicAdd(&pa_e[s_j[sb + m]], q4);
        }
      }
      s_pref[sb + m] = pref;
      s_frad[sb + m] = frad;
    }
    __syncwarp();
    // ---- C: forces.  With unit vectors the force bond m receives from its pairing with q is
    //   A (c u_m - u_q) / r_m + pref_q w'_m (G + P') u_m,   A = -(pref_m w_q + pref_q w_m) G'   (symmetric in m, q)
    // so the partner loop only accumulates sum A c, sum A u_q and sum pref_q (G + P'): 9 FP64 operations and 8 shared
    // loads per (m, q) instead of 19 and 9; 1/r_m, w'_m and u_m are applied once per bond.
    double fix = 0.0, fiy = 0.0, fiz = 0.0;
    double vi[6] = {0, 0, 0, 0, 0, 0};    // ATOM: this lane's share of the center's per-atom virial
    for (int m = sub; m < nb; m += G) {
      const double ux = s_dx[sb + m], uy = s_dy[sb + m], uz = s_dz[sb + m], rinvm = s_ri[sb + m];
      const double wm = s_w[sb + m], dwm = s_dw[sb + m], prefm = s_pref[sb + m];
      double sAc = 0.0, sB = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
      double fx = 0.0, fy = 0.0, fz = 0.0;
      double vm[6] = {0, 0, 0, 0, 0, 0};
      const double rm = ATOM ? s_r[sb + m] : 0.0;
      const double mx = ux * rm, my = uy * rm, mz = uz * rm;    // ATOM: d_m
      for (int k = 1; k < nb; k++) {
        int q = m + k;
        if (q >= nb) q -= nb;
        const int t = st + ((k <= nr) ? (k - 1) * CAP + m : (nb - k - 1) * CAP + q);
        const double prefn = s_pref[sb + q];
        const double ca = -(prefm * s_w[sb + q] + prefn * wm);
        const double A = ca * s_dg[t];
        const double gp = s_g[t] + dP;
        if (!ATOM) {
          sAc = fma(A, s_c[t], sAc);
          sB = fma(prefn, gp, sB);
          vx = fma(A, s_dx[sb + q], vx);
          vy = fma(A, s_dy[sb + q], vy);
          vz = fma(A, s_dz[sb + q], vz);
        } else {
          const double cmu = A * s_c[t] * rinvm + prefn * dwm * gp;    // multiplies u_m
          const double cnu = -A * rinvm;                                // multiplies u_q
          const double qx = cmu * ux + cnu * s_dx[sb + q], qy = cmu * uy + cnu * s_dy[sb + q], qz = cmu * uz + cnu * s_dz[sb + q];
          fx += qx;
          fy += qy;
          fz += qz;
          // v_tally3 (pair_rebomos.cpp:707-711): each (bond, k) triplet's r_ji (x) f_j + r_ki (x) f_k goes in thirds
          // to i, j, k.  (qx,qy,qz) is what bond m receives from its pairing with q -- once as the triplet's j, once
          // as the other triplet's k -- so (-d_m) (x) q is this lane's part of both; thirds to i, m, q.
          const double t0 = -mx * qx * (1.0 / 3.0), t1 = -my * qy * (1.0 / 3.0), t2 = -mz * qz * (1.0 / 3.0);
          const double t3 = -mx * qy * (1.0 / 3.0), t4 = -mx * qz * (1.0 / 3.0), t5 = -my * qz * (1.0 / 3.0);
          vi[0] += t0; vi[1] += t1; vi[2] += t2; vi[3] += t3; vi[4] += t4; vi[5] += t5;
          vm[0] += t0; vm[1] += t1; vm[2] += t2; vm[3] += t3; vm[4] += t4; vm[5] += t5;
          double *vq = pa_v + 6 * (size_t) s_j[sb + q];
          atomicAdd(vq, t0); atomicAdd(vq + 1, t1); atomicAdd(vq + 2, t2);
          atomicAdd(vq + 3, t3); atomicAdd(vq + 4, t4); atomicAdd(vq + 5, t5);
        }
      }
      const double fru = s_frad[sb + m];    // radial part, multiplies u_m
      if (!ATOM) {
        const double cu = rinvm * sAc + dwm * sB + fru;
        fx = cu * ux - rinvm * vx;
        fy = cu * uy - rinvm * vy;
        fz = cu * uz - rinvm * vz;
      } else {
        fx += fru * ux;
        fy += fru * uy;
        fz += fru * uz;
        // radial part: ev_tally (:444) and v_tally2 (:725): -d (x) (fr d), half to i, half to j
        const double h = -0.5 * fru * rinvm;
        const double r0 = h * mx * mx, r1 = h * my * my, r2 = h * mz * mz, r3 = h * mx * my, r4 = h * mx * mz, r5 = h * my * mz;
        vi[0] += r0; vi[1] += r1; vi[2] += r2; vi[3] += r3; vi[4] += r4; vi[5] += r5;
        double *vj = pa_v + 6 * (size_t) s_j[sb + m];
        atomicAdd(vj, vm[0] + r0); atomicAdd(vj + 1, vm[1] + r1); atomicAdd(vj + 2, vm[2] + r2);
        atomicAdd(vj + 3, vm[3] + r3); atomicAdd(vj + 4, vm[4] + r4); atomicAdd(vj + 5, vm[5] + r5);
      }
      fix -= fx;
      fiy -= fy;
      fiz -= fz;
      if (DET) {
        const size_t k = (size_t) i * B200MD_MAX_REBO + m;
        det.fb[3 * k] = fx;
        det.fb[3 * k + 1] = fy;
        det.fb[3 * k + 2] = fz;
        det.j[k] = s_j[sb + m];
      } else {
        const size_t j = s_j[sb + m];
        atomicAdd(&f[3 * j], fx);
        atomicAdd(&f[3 * j + 1], fy);
        atomicAdd(&f[3 * j + 2], fz);
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      fix += __shfl_xor_sync(FULL, fix, o);
      fiy += __shfl_xor_sync(FULL, fiy, o);
      fiz += __shfl_xor_sync(FULL, fiz, o);
    }
    if (sub == 0 && valid && !deferred) {
      if (DET) {
        det.nb[i] = nb;
        det.fi[3 * (size_t) i] = fix;
        det.fi[3 * (size_t) i + 1] = fiy;
        det.fi[3 * (size_t) i + 2] = fiz;
      } else if (nb > 0) {
        atomicAdd(&f[3 * (size_t) i], fix);
        atomicAdd(&f[3 * (size_t) i + 1], fiy);
        atomicAdd(&f[3 * (size_t) i + 2], fiz);
      }
    }
    if (ATOM && valid && !deferred) {
      atomicAdd(&pa_e[i], ei_atom);
#pragma unroll
      for (int k = 0; k < 6; k++) atomicAdd(&pa_v[6 * (size_t) i + k], vi[k]);
    }
    __syncwarp();
  }
  if (EV) block_accumulate<1, NT>(eacc, scal);
}

// deterministic mode: f[a] = F_center(a) + sum over the REBO neighbors k of a (in short-row order) of the
// force that center k's bond to a exerts on a.  Plain stores, fixed order -> bitwise reproducible.
__global__ void __launch_bounds__(BLOCK) rebo_gather_kernel(const int *__restrict__ short_idx,
                                                            const int *__restrict__ short_num, int rows, int ncen,
                                                            const DetTables det, double *__restrict__ f)
{
  const int a = blockIdx.x * BLOCK + threadIdx.x;
  if (a >= rows) return;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  if (a < ncen) {
    fx = det.fi[3 * (size_t) a];
    fy = det.fi[3 * (size_t) a + 1];
    fz = det.fi[3 * (size_t) a + 2];
  }
  const int n = short_num[a];
  const int *row = short_idx + (size_t) a * B200MD_SHORT_WIDTH;
  for (int e = 0; e < n; e++) {
    const int k = row[e];
    if (k >= ncen) continue;    // bonds centered on ghosts are evaluated by the ghost's owner
    const int nbk = det.nb[k];
    const int *jl = det.j + (size_t) k * B200MD_MAX_REBO;
    for (int m = 0; m < nbk; m++)
      if (jl[m] == a) {
        const size_t q = (size_t) k * B200MD_MAX_REBO + m;
        fx += det.fb[3 * q];
        fy += det.fb[3 * q + 1];
        fz += det.fb[3 * q + 2];
        break;
      }
  }
  f[3 * (size_t) a] += fx;
  f[3 * (size_t) a + 1] += fy;
  f[3 * (size_t) a + 2] += fz;
}

// centers by element (static between list builds): warp-aggregated append, so list order follows atom order
// at warp granularity
// Center lists in ASCENDING index order (stable compaction through a scan, not atomics): consecutive groups of a
// launch work on neighboring atoms, and an index range [t_lo, t_hi) of owned atoms maps to a contiguous piece of
// each list -- cen_scan[t] holds the list positions of threshold t for both elements (Mo count in the low 30
// bits, S count above), so ranged LJ launches need no host round trip.
__global__ void __launch_bounds__(BLOCK) center_key_kernel(const double4 *__restrict__ xq, int ncen,
                                                           int *__restrict__ key)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= ncen) return;
  const int t = elem_of(xq[i]);
  key[i] = (t == 0) ? 1 : ((t == 1) ? (1 << CEN_SHIFT) : 0);
}
__global__ void __launch_bounds__(BLOCK) center_lists_kernel(const int *__restrict__ key,
                                                             const long long *__restrict__ scan, int ncen,
                                                             int *__restrict__ listA, int *__restrict__ listB,
                                                             int *__restrict__ counts)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= ncen) return;
  const int k = key[i];
  const long long s = scan[i];
  if (k == 1) listA[(int) (s & CEN_MASK)] = i;
  else if (k != 0) listB[(int) (s >> CEN_SHIFT)] = i;
  if (i == ncen - 1) {
    const long long tot = s + k;
    counts[0] = (int) (tot & CEN_MASK);
    counts[1] = (int) (tot >> CEN_SHIFT);
  }
}

// ================================================================== K8: fdotr virial over owned + ghost
__global__ void __launch_bounds__(BLOCK) fdotr_kernel(const double4 *__restrict__ xq,
                                                      const double *__restrict__ f, int nall,
                                                      double *__restrict__ scal)
{
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * BLOCK + threadIdx.x; i < nall; i += gridDim.x * BLOCK) {
    const double4 x = xq[i];
    const double fx = f[3 * (size_t) i], fy = f[3 * (size_t) i + 1], fz = f[3 * (size_t) i + 2];
    v[0] += fx * x.x;
    v[1] += fy * x.y;
    v[2] += fz * x.z;
    v[3] += fy * x.x;
    v[4] += fz * x.x;
    v[5] += fz * x.y;
  }
  block_accumulate<6, BLOCK>(v, scal + 1);
}

// ================================================================== K5: tapered LJ, directed rows
// 8 lanes per owned atom, centers launched by element, rows segmented by partner element (all pair constants are
// immediate operands): each group streams its row segments as full sectors (U row loads and U position
// gathers in flight per lane), gathers one double4 sector per neighbor, reduces with 3 shuffles.
// Every directed pair is evaluated from both ends, so nothing is scattered: f_i is complete, energy and
// virial carry a factor 1/2.  Only ~40 % of the candidates are inside the LJ window, so both sides of that
// branch are kept minimal: the window and regime tests of pair_rebomos.cpp:518-543 (on rij = sqrt(rsq)) are
// replaced by their exact rsq equivalents precomputed on the host (no sqrt, no division outside the window).
// 1/a for a normal, positive a in a harmless range (here 1 <= rsq <= 200 A^2): MUFU.RCP64H seed + two Newton
// steps, no slow path.  Relative error <= 2 ulp, far inside the 1e-10 force tolerance.
__device__ __forceinline__ double rcp_nr(double a)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

// seed + ONE Newton step: relative error <= 2^-46 = 1.4e-14 (the seed is good to 2^-23), four orders inside the
// 1e-10 force tolerance; used by the pair kernel, whose FP64 pipe is the limiter
__device__ __forceinline__ double rcp_nr1(double a)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

// one row segment [lo, hi) whose partners all have element TJ: every pair constant is an immediate operand
template <bool EV, int PT, int U>
__device__ __forceinline__ void lj_segment(const RebomosDev &par, const double4 *__restrict__ xq,
                                           const int *__restrict__ row, int lo, int hi, int sub, const double4 &xi,
                                           double &fx, double &fy, double &fz, double (&ev)[7])
{
  int e = (lo & ~7) + sub;    // sector-aligned start; entries before lo belong to nobody
  int jn[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int k = e + 8 * u;
    jn[u] = (k >= lo && k < hi) ? ld_stream_int(row + k) : -1;
  }
  for (; e - sub < hi; e += 8 * U) {
    int jj[U];
    double4 xj[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      jj[u] = jn[u];
      if (jj[u] >= 0) xj[u] = ld_sector(xq + jj[u]);
    }
    // indices of the NEXT trip are requested before this trip's arithmetic: the row stream comes from HBM
    // (ncu r01: long-scoreboard stalls 8 per issue, all on the index load -> gather chain)
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int k = e + 8 * (U + u);
      jn[u] = (k < hi) ? ld_stream_int(row + k) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (jj[u] < 0) continue;
      const double dx = xi.x - xj[u].x, dy = xi.y - xj[u].y, dz = xi.z - xj[u].z;
      const double rsq = dx * dx + dy * dy + dz * dz;
      // rij > rcLJmax  <=>  rsq >= lj_out_hi ;  rij < rcLJmin  <=>  rsq < lj_in_lo   (exact, rsq_smallest_with_sqrt)
      if (rsq >= par.lj_out_hi[PT] || rsq < par.lj_in_lo[PT]) continue;
      double VLJ, fpair;
      if (rsq >= par.lj_s95[PT]) {               // rij >= 0.95 sigma: 12-6 LJ
        const double r2inv = rcp_nr(rsq);
        const double r6inv = r2inv * r2inv * r2inv;
        fpair = r6inv * (par.lj1[PT] * r6inv - par.lj2[PT]) * r2inv;
        if (EV) VLJ = r6inv * (par.lj3[PT] * r6inv - par.lj4[PT]);
      } else {                                   // cubic taper down to rcLJmin (rare)
        const double rij = sqrt(rsq);
        const double drp = rij - par.rcLJmin[PT];
        VLJ = drp * drp * (drp * par.c3[PT] + par.c2[PT]);
        const double dVLJ = drp * (3.0 * drp * par.c3[PT] + 2.0 * par.c2[PT]);
        fpair = -dVLJ / rij;
      }
      fx += dx * fpair;
      fy += dy * fpair;
      fz += dz * fpair;
      if (EV) {
        ev[0] += 0.5 * VLJ;
        const double hf = 0.5 * fpair;
        ev[1] += dx * dx * hf;
        ev[2] += dy * dy * hf;
        ev[3] += dz * dz * hf;
        ev[4] += dx * dy * hf;
        ev[5] += dx * dz * hf;
        ev[6] += dy * dz * hf;
      }
    }
  }
}

template <bool EV, int ELEM, int U, int MINB, bool ATOM>
__global__ void __launch_bounds__(BLOCK, MINB) lj_kernel(const __grid_constant__ RebomosDev par,
                                                         const double4 *__restrict__ xq,
                                                         const int64_t *__restrict__ lj_off,
                                                         const int *__restrict__ lj_num,
                                                         const int *__restrict__ lj_val,
                                                         const int *__restrict__ cen_list,
                                                         const long long *__restrict__ cen_scan, int t_lo, int t_hi,
                                                         double *__restrict__ f, double *__restrict__ scal,
                                                         double *__restrict__ pa_e, double *__restrict__ pa_v)
{
  // centers of this element whose atom index lies in [t_lo, t_hi): a contiguous piece of the ascending list
  const long long s0 = cen_scan[t_lo], s1 = cen_scan[t_hi];
  const int first = (ELEM == 0) ? (int) (s0 & CEN_MASK) : (int) (s0 >> CEN_SHIFT);
  const int count = (ELEM == 0) ? (int) (s1 & CEN_MASK) : (int) (s1 >> CEN_SHIFT);
  const int sub = threadIdx.x & 7;
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int g = first + ((blockIdx.x * BLOCK + threadIdx.x) >> 3); g < count; g += (gridDim.x * BLOCK) >> 3) {
    const int i = cen_list[g];
    const double4 xi = xq[i];
    const int64_t off = lj_off[i];
    const int cap = (int) (lj_off[i + 1] - off);
    const int nA = lj_num[2 * i], nB = lj_num[2 * i + 1];
    const int *row = lj_val + off;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    double ca[7] = {0, 0, 0, 0, 0, 0, 0};    // this center's energy/virial share (every pair visited from both ends)
    lj_segment<EV, ELEM * 2, U>(par, xq, row, 0, nA, sub, xi, fx, fy, fz, ca);
    lj_segment<EV, ELEM * 2 + 1, U>(par, xq, row, cap - nB, cap, sub, xi, fx, fy, fz, ca);
    fx = group_sum<8>(fx);
    fy = group_sum<8>(fy);
    fz = group_sum<8>(fz);
    if (sub == 0) {
      // one add per owned atom; atomic because the bond-order launches (which scatter into f) may run beside this
      // kernel on another stream.  A single adder per address: the result does not depend on the schedule.
      atomicAdd(f + 3 * (size_t) i, fx);
      atomicAdd(f + 3 * (size_t) i + 1, fy);
      atomicAdd(f + 3 * (size_t) i + 2, fz);
    }
    if (EV) {
#pragma unroll
      for (int k = 0; k < 7; k++) ev[k] += ca[k];
    }
    if (ATOM) {
      // ev_tally (pair_rebomos.cpp:554): half of the pair's energy and virial to each end = exactly what this
      // center accumulated over its directed row; rebo_center_kernel added its part with atomics before this kernel
#pragma unroll
      for (int k = 0; k < 7; k++) ca[k] = group_sum<8>(ca[k]);
      if (sub == 0) {
        pa_e[i] += ca[0];
#pragma unroll
        for (int k = 0; k < 6; k++) pa_v[6 * (size_t) i + k] += ca[1 + k];
      }
    }
  }
  if (EV) block_accumulate<7, BLOCK>(ev, scal);
}

// ================================================================== K5b: tapered LJ over PAIRS of centers
// ncu (r01 v3/v6): lj_kernel is bound by the L1 data pipe, which delivers one gathered 32-byte position sector per
// cycle and SM -- one per (center, candidate).  Two neighboring centers of the same element share most of their
// candidates (spheres of radius rcLJmax + margin whose centers are ~3-5 A apart: union = 1.2-1.3 x one sphere), so a
// lane group works on a PAIR (a, b) of consecutive centers of the ascending center list and streams the UNION row:
// one gathered sector serves two interactions, 30-40 % fewer sectors per atom for one extra distance test per
// candidate.  Row layout, element segmentation, exact rsq window thresholds and arithmetic are those of lj_kernel.
// A center without partner (odd count) has b = -1 and a far-away dummy position (every candidate fails the window).
// One (center, candidate) term, BRANCH-FREE in the 12-6 regime: ncu r01 v6 showed the branchy form bound by
// fixed-latency dependency stalls ("wait" 4.2 per issue) -- every term was its own divergent region, so the FP64 chains
// of the 2 x U terms of a trip could not overlap.  Here a term outside the window runs the same arithmetic on a
// harmless rsq (1.0) and contributes fpair = 0, and the compiler interleaves all chains of a trip.  The cubic-taper
// regime (rcLJmin <= r < 0.95 sigma; no pair of a bulk crystal is in it) stays a branch, taken only by warps that have one.
template <bool EV, int PT>
__device__ __forceinline__ bool lj_term(const RebomosDev &par, const double xi, const double yi, const double zi,
                                        const double4 &xj, double &fx, double &fy, double &fz, double (&ev)[7])
{
  const double dx = xi - xj.x, dy = yi - xj.y, dz = zi - xj.z;
  const double rsq = dx * dx + dy * dy + dz * dz;
  // rsq >= 0: comparing the bit patterns as integers is the same (exact) comparison and runs on the integer pipe
  // instead of the FP64 pipe, which is the busiest one in this kernel
  const long long rb = __double_as_longlong(rsq);
  const bool in126 = rb < __double_as_longlong(par.lj_out_hi[PT]) && rb >= __double_as_longlong(par.lj_s95[PT]);
  const double r2inv = rcp_nr1(in126 ? rsq : 1.0);
  const double r6inv = r2inv * r2inv * r2inv;
  double fpair = r6inv * (par.lj1[PT] * r6inv - par.lj2[PT]) * r2inv;
  fpair = in126 ? fpair : 0.0;
  fx += dx * fpair;
  fy += dy * fpair;
  fz += dz * fpair;
  if (EV) {
    double VLJ = r6inv * (par.lj3[PT] * r6inv - par.lj4[PT]);
    VLJ = in126 ? VLJ : 0.0;
    ev[0] += 0.5 * VLJ;
    const double hf = 0.5 * fpair;
    ev[1] += dx * dx * hf;
    ev[2] += dy * dy * hf;
    ev[3] += dz * dz * hf;
    ev[4] += dx * dy * hf;
    ev[5] += dx * dz * hf;
    ev[6] += dy * dz * hf;
  }
  // the caller takes ONE rarely-taken branch per trip for all terms that report the taper regime
  return rb < __double_as_longlong(par.lj_s95[PT]) && rb >= __double_as_longlong(par.lj_in_lo[PT]);
}

// cubic taper down to rcLJmin (pair_rebomos.cpp:533-543) for a term lj_term() reported: {fpair, VLJ} by value, so
// that the out-of-line call does not force the caller's accumulators into local memory
template <int PT>
__device__ __noinline__ double2 lj_taper(const RebomosDev &par, double rsq)
{
  if (!(rsq < par.lj_s95[PT] && rsq >= par.lj_in_lo[PT])) return make_double2(0.0, 0.0);
  const double rij = sqrt(rsq);
  const double drp = rij - par.rcLJmin[PT];
  const double VLJ = drp * drp * (drp * par.c3[PT] + par.c2[PT]);
  const double dVLJ = drp * (3.0 * drp * par.c3[PT] + 2.0 * par.c2[PT]);
  return make_double2(-dVLJ / rij, VLJ);
}
template <bool EV, int PT>
__device__ __forceinline__ void lj_term_taper(const RebomosDev &par, const double xi, const double yi, const double zi,
                                              const double4 &xj, double &fx, double &fy, double &fz, double (&ev)[7])
{
  const double dx = xi - xj.x, dy = yi - xj.y, dz = zi - xj.z;
  const double2 t = lj_taper<PT>(par, dx * dx + dy * dy + dz * dz);
  fx += dx * t.x;
  fy += dy * t.x;
  fz += dz * t.x;
  if (EV) {
    ev[0] += 0.5 * t.y;
    const double hf = 0.5 * t.x;
    ev[1] += dx * dx * hf;
    ev[2] += dy * dy * hf;
    ev[3] += dz * dz * hf;
    ev[4] += dx * dy * hf;
    ev[5] += dx * dz * hf;
    ev[6] += dy * dz * hf;
  }
}

// One row segment [lo, hi) of a union row for the two centers of a pair, software-pipelined: D position buffers per
// lane; while the two terms of one buffer are computed the gathers of the other D-1 are in flight, and the row indices
// run two rounds ahead of the gathers (ncu r01 v6: with load-then-use in the same trip the kernel sat on long-scoreboard
// stalls, 6.3 per issue at 34 % occupancy).  An empty slot holds a far-away dummy position that fails every window test.
template <bool EV, int PT, int D>
__device__ __forceinline__ void ljp_segment(const RebomosDev &par, const double4 *__restrict__ xq,
                                            const int *__restrict__ row, int lo, int hi, int sub, const double4 &xa,
                                            const double4 &xb, double (&fa)[3], double (&fb)[3], double (&eva)[7],
                                            double (&evb)[7])
{
  int e = (lo & ~7) + sub;
  double4 xj[D];
  int jn[D];
#pragma unroll
  for (int d = 0; d < D; d++) {
    const int k = e + 8 * d;
    const int j = (k >= lo && k < hi) ? ld_stream_int(row + k) : -1;
    xj[d] = make_double4(1.0e30, 1.0e30, 1.0e30, 0.0);
    if (j >= 0) xj[d] = ld_sector(xq + j);
  }
#pragma unroll
  for (int d = 0; d < D; d++) {
    const int k = e + 8 * (D + d);
    jn[d] = (k < hi) ? ld_stream_int(row + k) : -1;
  }
  for (; e - sub < hi; e += 8 * D) {
#pragma unroll
    for (int d = 0; d < D; d++) {
      const double4 x = xj[d];
      bool taper = lj_term<EV, PT>(par, xa.x, xa.y, xa.z, x, fa[0], fa[1], fa[2], eva);
      taper |= lj_term<EV, PT>(par, xb.x, xb.y, xb.z, x, fb[0], fb[1], fb[2], evb);
      if (taper) {
        lj_term_taper<EV, PT>(par, xa.x, xa.y, xa.z, x, fa[0], fa[1], fa[2], eva);
        lj_term_taper<EV, PT>(par, xb.x, xb.y, xb.z, x, fb[0], fb[1], fb[2], evb);
      }
      const int j = jn[d];
      xj[d] = make_double4(1.0e30, 1.0e30, 1.0e30, 0.0);
      if (j >= 0) xj[d] = ld_sector(xq + j);
      const int k = e + 8 * (2 * D + d);
      jn[d] = (k < hi) ? ld_stream_int(row + k) : -1;
    }
  }
}

template <bool EV, int ELEM, int D, int MINB, bool ATOM, int NT>
__global__ void __launch_bounds__(NT, MINB) lj_pair_kernel(const __grid_constant__ RebomosDev par,
                                                              const double4 *__restrict__ xq,
                                                              const int64_t *__restrict__ ljp_off,
                                                              const int *__restrict__ ljp_num,
                                                              const int2 *__restrict__ ljp_ab,
                                                              const int *__restrict__ lj_val, int P,
                                                              const long long *__restrict__ cen_scan, int t_lo, int t_hi,
                                                              double *__restrict__ f, double *__restrict__ scal,
                                                              double *__restrict__ pa_e, double *__restrict__ pa_v)
{
  // pairs of this element with a center whose atom index lies in [t_lo, t_hi): list positions [s0, s1) -> pairs
  // [(s0+1)/2, (s1+1)/2); the pair straddling t_lo was done by the range below, the one straddling t_hi is done here
  const long long c0 = cen_scan[t_lo], c1 = cen_scan[t_hi];
  const int s0 = (ELEM == 0) ? (int) (c0 & CEN_MASK) : (int) (c0 >> CEN_SHIFT);
  const int s1 = (ELEM == 0) ? (int) (c1 & CEN_MASK) : (int) (c1 >> CEN_SHIFT);
  const int first = (s0 + 1) >> 1, last = (s1 + 1) >> 1;
  const int sub = threadIdx.x & 7;
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int g = first + ((blockIdx.x * NT + threadIdx.x) >> 3); g < last; g += (gridDim.x * NT) >> 3) {
    const int q = ELEM * P + g;
    const int2 ab = ljp_ab[q];
    const double4 xa = xq[ab.x];
    double4 xb = make_double4(1.0e30, 1.0e30, 1.0e30, 0.0);    // no partner: nothing is in its window
    if (ab.y >= 0) xb = xq[ab.y];
    const int64_t off = ljp_off[q];
    const int cap = (int) (ljp_off[q + 1] - off);
    const int nA = ljp_num[2 * q], nB = ljp_num[2 * q + 1];
    const int *row = lj_val + off;
    double fa[3] = {0, 0, 0}, fb[3] = {0, 0, 0};
    double ca[7] = {0, 0, 0, 0, 0, 0, 0}, cb[7] = {0, 0, 0, 0, 0, 0, 0};
    if (ATOM) {
      ljp_segment<EV, ELEM * 2, D>(par, xq, row, 0, nA, sub, xa, xb, fa, fb, ca, cb);
      ljp_segment<EV, ELEM * 2 + 1, D>(par, xq, row, cap - nB, cap, sub, xa, xb, fa, fb, ca, cb);
    } else {
      ljp_segment<EV, ELEM * 2, D>(par, xq, row, 0, nA, sub, xa, xb, fa, fb, ev, ev);
      ljp_segment<EV, ELEM * 2 + 1, D>(par, xq, row, cap - nB, cap, sub, xa, xb, fa, fb, ev, ev);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      fa[k] = group_sum<8>(fa[k]);
      fb[k] = group_sum<8>(fb[k]);
    }
    // one add per owned atom (a single adder per address: schedule-independent)
    if (sub < 3) atomicAdd(f + 3 * (size_t) ab.x + sub, sub == 0 ? fa[0] : (sub == 1 ? fa[1] : fa[2]));
    else if (sub < 6 && ab.y >= 0) atomicAdd(f + 3 * (size_t) ab.y + (sub - 3), sub == 3 ? fb[0] : (sub == 4 ? fb[1] : fb[2]));
    if (ATOM) {
      // ev_tally (pair_rebomos.cpp:554): half of each pair's energy and virial to each end = what a center accumulated
#pragma unroll
      for (int k = 0; k < 7; k++) {
        ca[k] = group_sum<8>(ca[k]);
        cb[k] = group_sum<8>(cb[k]);
        ev[k] += (sub == 0) ? ca[k] + cb[k] : 0.0;
      }
      if (sub == 0) {
        pa_e[ab.x] += ca[0];
#pragma unroll
        for (int k = 0; k < 6; k++) pa_v[6 * (size_t) ab.x + k] += ca[1 + k];
      } else if (sub == 1 && ab.y >= 0) {
        pa_e[ab.y] += cb[0];
#pragma unroll
        for (int k = 0; k < 6; k++) pa_v[6 * (size_t) ab.y + k] += cb[1 + k];
      }
    }
  }
  if (EV) block_accumulate<7, NT>(ev, scal);
}

// capacity of every pair row = master rows of both centers (8-aligned by the scan that follows)
__global__ void __launch_bounds__(BLOCK) ljpair_cap_kernel(const int *__restrict__ cen_list, int inum, int P,
                                                           const int *__restrict__ counts,
                                                           const int *__restrict__ list_num, int *__restrict__ cap,
                                                           int2 *__restrict__ ab_out)
{
  const int q = blockIdx.x * BLOCK + threadIdx.x;
  if (q >= 2 * P) return;
  const int E = q / P, p = q - E * P;
  const int cnt = counts[E];
  const int *list = cen_list + (size_t) E * (inum + 32);
  int a = -1, b = -1, n = 0;
  if (2 * p < cnt) {
    a = list[2 * p];
    n = list_num[a];
    if (2 * p + 1 < cnt) {
      b = list[2 * p + 1];
      n += list_num[b];
    }
  }
  cap[q] = n;
  ab_out[q] = make_int2(a, b);
}


// ---- halo overlap (resident loop): the PAIR ROWS are put in [interior | boundary] order per element before the union
// rows are built -- pairs keep their composition (consecutive centers of the ascending lists), only their slot changes.
// A pair is interior when both centers are farther than g.r from every face of the sub-domain (lamda space if
// triclinic), so that no candidate of its union row is a ghost: its LJ launch may run while the forward halo is in flight.
__device__ __forceinline__ bool split_interior(const SplitGeom &g, const double4 &p)
{
  double c0 = p.x, c1 = p.y, c2 = p.z;
  if (g.triclinic) {
    const double d0 = p.x - g.boxlo[0], d1 = p.y - g.boxlo[1], d2 = p.z - g.boxlo[2];
    c0 = g.h_inv[0] * d0 + g.h_inv[5] * d1 + g.h_inv[4] * d2;
    c1 = g.h_inv[1] * d1 + g.h_inv[3] * d2;
    c2 = g.h_inv[2] * d2;
  }
  return c0 - g.lo[0] > g.r[0] && g.hi[0] - c0 > g.r[0] && c1 - g.lo[1] > g.r[1] && g.hi[1] - c1 > g.r[1] &&
         c2 - g.lo[2] > g.r[2] && g.hi[2] - c2 > g.r[2];
}
// scan keys of the two parts: element 0 counts in the low CEN_SHIFT bits, element 1 above (as in cen_key)
__global__ void __launch_bounds__(BLOCK) ljpair_split_key_kernel(const int2 *__restrict__ ab, int P,
                                                                 const double4 *__restrict__ xq,
                                                                 const __grid_constant__ SplitGeom g,
                                                                 int *__restrict__ keyI, int *__restrict__ keyB)
{
  const int q = blockIdx.x * BLOCK + threadIdx.x;
  if (q >= 2 * P) return;
  const int2 p = ab[q];
  int kI = 0, kB = 0;
  if (p.x >= 0) {
    bool interior = split_interior(g, xq[p.x]);
    if (interior && p.y >= 0) interior = split_interior(g, xq[p.y]);
    const int key = (q < P) ? 1 : (1 << CEN_SHIFT);
    if (interior) kI = key;
    else kB = key;
  }
  keyI[q] = kI;
  keyB[q] = kB;
}
__global__ void __launch_bounds__(BLOCK) ljpair_split_permute_kernel(const int2 *__restrict__ ab, const int *__restrict__ cap,
                                                                     int P, const int *__restrict__ keyI,
                                                                     const int *__restrict__ keyB,
                                                                     const long long *__restrict__ scanI,
                                                                     const long long *__restrict__ scanB,
                                                                     int2 *__restrict__ ab_out, int *__restrict__ cap_out,
                                                                     long long *__restrict__ split)
{
  const int q = blockIdx.x * BLOCK + threadIdx.x;
  if (q >= 2 * P) return;
  const long long totI = scanI[2 * P], totB = scanB[2 * P];    // exclusive scans carry the totals in the last entry
  const int nI0 = (int) (totI & CEN_MASK), nI1 = (int) (totI >> CEN_SHIFT);
  const int E = q < P ? 0 : 1;
  if (keyI[q]) {
    const int pos = E ? (int) (scanI[q] >> CEN_SHIFT) : (int) (scanI[q] & CEN_MASK);
    ab_out[E * P + pos] = ab[q];
    cap_out[E * P + pos] = cap[q];
  } else if (keyB[q]) {
    const int pos = (E ? nI1 : nI0) + (E ? (int) (scanB[q] >> CEN_SHIFT) : (int) (scanB[q] & CEN_MASK));
    ab_out[E * P + pos] = ab[q];
    cap_out[E * P + pos] = cap[q];
  }
  if (q == 0) {
    // lj_pair_kernel turns list positions [s0, s1) into pairs [(s0+1)/2, (s1+1)/2): positions = 2 x pair slots
    const int n0 = nI0 + (int) (totB & CEN_MASK), n1 = nI1 + (int) (totB >> CEN_SHIFT);
    split[0] = 0;
    split[1] = (long long) (2 * nI0) | ((long long) (2 * nI1) << CEN_SHIFT);
    split[2] = (long long) (2 * n0) | ((long long) (2 * n1) << CEN_SHIFT);
  }
}

// union rows: one warp per pair, order-preserving ballot compaction, partners segmented by element as in lj rows.
// pass 1: candidates of a's master row within the margin sphere of a (remembered in a per-warp hash set in shared
// memory); pass 2: candidates of b's master row within the margin sphere of b that pass 1 did not take.  The test is
// set membership, not geometry: after atoms have moved, an atom inside a's margin sphere need not be in a's master
// row (it cannot reach a's cutoff before the next master rebuild, but it may reach b's), and a itself never is.
#define LJP_BLOCK 128
#define LJP_HT 2048
__global__ void __launch_bounds__(LJP_BLOCK) build_ljpair_kernel(
    const __grid_constant__ RebomosDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ list_off,
    const int *__restrict__ list_num, const int *__restrict__ list_val, int P, const int2 *__restrict__ ljp_ab,
    const int64_t *__restrict__ ljp_off, int *__restrict__ ljp_num, int *__restrict__ lj_val, int *__restrict__ flags)
{
  __shared__ int s_tab[LJP_BLOCK / 32][LJP_HT];
  const int lane = threadIdx.x & 31;
  const int q = (int) (((size_t) blockIdx.x * LJP_BLOCK + threadIdx.x) >> 5);
  if (q >= 2 * P) return;
  const int2 ab = ljp_ab[q];
  if (ab.x < 0) {
    if (lane == 0) ljp_num[2 * q] = ljp_num[2 * q + 1] = 0;
    return;
  }
  int *tab = s_tab[threadIdx.x >> 5];
  for (int k = lane; k < LJP_HT; k += 32) tab[k] = -1;
  __syncwarp();
  const int ti = q / P;
  const double4 xa = xq[ab.x];
  const int64_t base = ljp_off[q];
  const int cap = (int) (ljp_off[q + 1] - base);
  const unsigned lt = (1u << lane) - 1u;
  int nlA = 0, nlB = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int c = pass ? ab.y : ab.x;
    if (c < 0) break;
    const double4 xc = pass ? xq[c] : xa;
    const int n = list_num[c];
    const int64_t mb = list_off[c];
    for (int e0 = 0; e0 < n; e0 += 32) {
      const int e = e0 + lane;
      bool pl = false;
      int j = 0, tj = 0;
      if (e < n) {
        j = ld_stream_int(list_val + mb + e) & B200MD_NEIGHMASK;
        const double4 xj = xq[j];
        tj = elem_of(xj);
        if (tj >= 0) {
          const double dx = xc.x - xj.x, dy = xc.y - xj.y, dz = xc.z - xj.z;
          pl = dx * dx + dy * dy + dz * dz <= par.ljsq[ti * 2 + tj];
        }
        if (pl) {
          unsigned h = ((unsigned) j * 2654435761u) >> 21;    // 11 bits
          if (pass == 0) {
            while (atomicCAS(&tab[h], -1, j) != -1) h = (h + 1) & (LJP_HT - 1);
          } else {
            int v;
            while ((v = tab[h]) != -1) {
              if (v == j) {
                pl = false;
                break;
              }
              h = (h + 1) & (LJP_HT - 1);
            }
          }
        }
      }
      const unsigned mA = __ballot_sync(0xffffffffu, pl && tj == 0);
      const unsigned mB = __ballot_sync(0xffffffffu, pl && tj == 1);
      if (pl) {
        if (tj == 0) lj_val[base + nlA + __popc(mA & lt)] = j;
        else lj_val[base + cap - 1 - (nlB + __popc(mB & lt))] = j;
      }
      nlA += __popc(mA);
      nlB += __popc(mB);
      if (pass == 0 && nlA + nlB > LJP_HT * 3 / 4 - 32) {    // warp-uniform: the hash set would fill up
        if (lane == 0) flags[0] = 1;
        break;
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    ljp_num[2 * q] = nlA;
    ljp_num[2 * q + 1] = nlB;
    atomicAdd(&flags[5], nlA + nlB);
  }
}

// plugin-mode upload pipelining: for every range of owned centers, the largest OWNED atom index its bond-order
// evaluation reads (the centers themselves and their short-row candidates; ghosts are uploaded first).  Positions
// arrive in ascending index order, so a range may start as soon as the piece holding that index is on the device.
struct ChunkBounds {
  int t[B200MD_MAX_D2H_CHUNKS + 1];
  int K;
};
// dep[k] = largest atom index the short rows of range k name.  With a straggler list (strag_flag != NULL) a range is
// allowed its own piece and the next one; atoms named from farther away -- the neighbors of an atom that was wrapped
// through a periodic face since the last Atom::sort keeps its index and now sits at the other end of the box -- go to the
// list instead (each once), and their positions travel ahead of the pieces.  Without it ONE such atom made range 0 wait
// for the last piece: the whole upload in front of the first kernel, 1.77 -> 2.12 ms per call after 1000 steps.
__global__ void __launch_bounds__(BLOCK) dep_range_kernel(const int *__restrict__ short_idx,
                                                          const int *__restrict__ short_num, int inum,
                                                          const __grid_constant__ ChunkBounds cb, int *__restrict__ dep,
                                                          int *__restrict__ strag_flag, int *__restrict__ strag_list,
                                                          int *__restrict__ strag_count, int strag_cap)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= inum) return;
  int k = 0;
  while (k + 1 < cb.K && i >= cb.t[k + 1]) k++;
  const int far = strag_flag ? cb.t[min(k + 2, cb.K)] : inum;    // first index beyond the pieces range k may wait for
  int m = i;
  const int n = short_num[i];
  for (int e = 0; e < n; e++) {
    const int j = short_idx[(size_t) i * B200MD_SHORT_WIDTH + e];
    if (j >= inum) continue;    // ghosts travel first
    if (j >= far) {
      if (atomicExch(&strag_flag[j], 1) == 0) {
        const int pos = atomicAdd(strag_count, 1);
        if (pos < strag_cap) strag_list[pos] = j;
      }
    } else if (j > m)
      m = j;
  }
  atomicMax(&dep[k], m);
}
// stragglers: positions gathered on the host into a small pinned buffer, scattered into xq here
__global__ void __launch_bounds__(BLOCK) strag_scatter_kernel(const double *__restrict__ buf, const int *__restrict__ list,
                                                              int n, const int *__restrict__ type,
                                                              const int *__restrict__ map, int ntypes,
                                                              double4 *__restrict__ xq)
{
  const int q = blockIdx.x * BLOCK + threadIdx.x;
  if (q >= n) return;
  const int j = list[q];
  const int t = type[j];
  const int e = (t >= 1 && t <= ntypes) ? map[t] : -1;
  xq[j] = make_double4(buf[3 * (size_t) q], buf[3 * (size_t) q + 1], buf[3 * (size_t) q + 2], (double) e);
}

// ================================================================== tight rows (third list level)
// The force kernels stream rows filtered to rcut + margin_t.  They are derived from the inner ("wide") rows, never from
// the master rows: a pair that reaches the cutoff before the next tight derive is within rcut + margin_t now (each atom
// moves less than margin_t/2 until then), and it is in the wide rows because those hold every master-list pair that
// was within rcut + margin when they were derived and are themselves re-derived at margin/2.  Order inside a row (and
// inside each element segment of an LJ pair row) is preserved, so results do not depend on the level structure.
struct TightLimits {
  double shortsq[4], ljsq[4];
};
__global__ void __launch_bounds__(BLOCK) derive_tight_short_kernel(const double4 *__restrict__ xq,
                                                                   const int *__restrict__ short_idx,
                                                                   const int *__restrict__ short_num, int inum,
                                                                   const __grid_constant__ TightLimits lim,
                                                                   int *__restrict__ out_idx, int *__restrict__ out_num)
{
  // 8 lanes per owned row, ballot compaction inside the lane group.  All candidates of a row (<= 64) are fetched before
  // the first is tested -- 4 gathers in flight per lane; one at a time the kernel ran at the latency of 3 dependent round
  // trips per row -- and the trip count is the warp's largest, so the votes run on the full warp.  The limits are a
  // __grid_constant__: passed by value and indexed at run time they were copied to local memory by every thread
  // (64 M local-store sectors per launch of the LJ pass, ncu r02).
  const int sub = threadIdx.x & 7;
  const int i = (blockIdx.x * BLOCK + threadIdx.x) >> 3;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gshift = lane & ~7u;
  const bool live = i < inum;
  const double4 xi = xq[live ? i : 0];
  const int ti = elem_of(xi);
  const int n = (live && ti >= 0) ? short_num[i] : 0;
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  const int *row = short_idx + (size_t) (live ? i : 0) * B200MD_SHORT_WIDTH;
  int *orow = out_idx + (size_t) (live ? i : 0) * B200MD_SHORT_WIDTH;
  int cnt = 0;
  constexpr int U = 4;
  for (int e0 = 0; e0 < nmax; e0 += 8 * U) {
    int j[U];
    double4 xj[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e = e0 + u * 8 + sub;
      j[u] = (e < n) ? row[e] : -1;
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (j[u] >= 0) xj[u] = ld_sector(xq + j[u]);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (e0 + u * 8 >= nmax) break;    // warp-uniform
      bool keep = false;
      if (j[u] >= 0) {
        const int tj = elem_of(xj[u]);
        const double dx = xi.x - xj[u].x, dy = xi.y - xj[u].y, dz = xi.z - xj[u].z;
        keep = tj >= 0 && dx * dx + dy * dy + dz * dz <= lim.shortsq[ti * 2 + (tj & 1)];
      }
      const unsigned bits = (__ballot_sync(0xffffffffu, keep) >> gshift) & 0xffu;
      if (keep) orow[cnt + __popc(bits & ((1u << sub) - 1u))] = j[u];
      cnt += __popc(bits);
    }
  }
  if (sub == 0 && live) out_num[i] = cnt;
}

template <int TU, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) derive_tight_lj_kernel(const double4 *__restrict__ xq,
                                                                      const int64_t *__restrict__ ljp_off,
                                                                      const int *__restrict__ ljp_num,
                                                                      const int2 *__restrict__ ljp_ab,
                                                                      const int *__restrict__ lj_val, int P,
                                                                      const int *__restrict__ cen_counts,
                                                                      const __grid_constant__ TightLimits lim,
                                                                      int *__restrict__ out_val, int *__restrict__ out_num)
{
  // One warp per pair row.  The row is walked as ONE sequence -- the Mo partners from the front of the slot range, then
  // the S partners from its back -- TU x 32 entries per trip, and the indices of the next trip are fetched before the
  // positions of this one are waited for: the kernel is bound by the chain index -> position -> vote (ncu r02: 17.7
  // long-scoreboard stalls per issue, 0.77 ms per 1 M atoms with one segment and one trip at a time).
  // Only the slots that hold a pair get a warp: ceil(nMo / 2) from slot 0 and ceil(nS / 2) from slot P (a warp per slot
  // of the 2 P, half of them empty, spent 13 % of the kernel's stall samples waiting to learn that its slot was empty)
  const int lane = threadIdx.x & 31;
  const int w = (int) (((size_t) blockIdx.x * BLOCK + threadIdx.x) >> 5);
  const int p0 = (cen_counts[0] + 1) >> 1, p1 = (cen_counts[1] + 1) >> 1;
  if (w >= p0 + p1) return;
  const int q = w < p0 ? w : P + (w - p0);
  const int2 ab = ljp_ab[q];
  if (ab.x < 0) {
    if (lane == 0) out_num[2 * q] = out_num[2 * q + 1] = 0;
    return;
  }
  const int64_t base = ljp_off[q];
  const int cap = (int) (ljp_off[q + 1] - base);
  const int nA = ljp_num[2 * q], nB = ljp_num[2 * q + 1];
  const int ntot = nA + nB;
  const int *row = lj_val + base;
  int *orow = out_val + base;
  // entry e of the sequence sits at row[e] (e < nA) or row[cap - 1 - (e - nA)] (S partners fill the slot range from its back)
  int jn[TU];
#pragma unroll
  for (int u = 0; u < TU; u++) {
    const int e = 32 * u + lane;
    jn[u] = (e < ntot) ? row[e < nA ? e : cap - 1 - (e - nA)] : -1;
  }
  const int ti = q / P;
  const double4 xa = xq[ab.x];
  double4 xb = make_double4(1.0e30, 1.0e30, 1.0e30, 0.0);
  if (ab.y >= 0) xb = xq[ab.y];
  const double limA = lim.ljsq[ti * 2], limB = lim.ljsq[ti * 2 + 1];
  const unsigned lt = (1u << lane) - 1u;
  int cntA = 0, cntB = 0;
  for (int e0 = 0; e0 < ntot; e0 += 32 * TU) {
    int j[TU];
    double4 xj[TU];
#pragma unroll
    for (int u = 0; u < TU; u++) {
      j[u] = jn[u];
      xj[u] = make_double4(1.0e30, 1.0e30, 1.0e30, 0.0);
      if (j[u] >= 0) xj[u] = ld_sector(xq + j[u]);
    }
#pragma unroll
    for (int u = 0; u < TU; u++) {
      const int e = e0 + 32 * (TU + u) + lane;
      jn[u] = (e < ntot) ? row[e < nA ? e : cap - 1 - (e - nA)] : -1;
    }
#pragma unroll
    for (int u = 0; u < TU; u++) {
      if (e0 + 32 * u >= ntot) break;    // warp-uniform
      const bool isA = e0 + 32 * u + lane < nA;
      double dx = xa.x - xj[u].x, dy = xa.y - xj[u].y, dz = xa.z - xj[u].z;
      const double ra = dx * dx + dy * dy + dz * dz;
      dx = xb.x - xj[u].x, dy = xb.y - xj[u].y, dz = xb.z - xj[u].z;
      const double rb = dx * dx + dy * dy + dz * dz;
      const double limit = isA ? limA : limB;
      const bool keep = (j[u] >= 0) && (ra <= limit || rb <= limit);
      const unsigned bA = __ballot_sync(0xffffffffu, keep && isA), bB = __ballot_sync(0xffffffffu, keep && !isA);
      if (keep) {
        if (isA) orow[cntA + __popc(bA & lt)] = j[u];
        else orow[cap - 1 - (cntB + __popc(bB & lt))] = j[u];
      }
      cntA += __popc(bA);
      cntB += __popc(bB);
    }
  }
  if (lane == 0) {
    out_num[2 * q] = cntA;
    out_num[2 * q + 1] = cntB;
  }
}

// ================================================================== host side
static inline int nblocks(long long n, int per) { return (int) ((n + per - 1) / per); }

// smallest double x with sqrt(x) > T (strict) or sqrt(x) >= T; sqrt is correctly rounded and monotonic on
// host and device, so `rsq >= this` is EXACTLY the reference's comparison on rij = sqrt(rsq)
static double rsq_smallest_with_sqrt(double T, bool strict)
{
  auto ok = [&](double x) { return strict ? sqrt(x) > T : sqrt(x) >= T; };
  double x = T * T;
  while (ok(x)) x = nextafter(x, 0.0);
  while (!ok(x)) x = nextafter(x, 1.0e300);
  return x;
}

static void derive_params(b200md_ctx *c, const b200md_rebomos_params *p)
{
  RebomosDev &d = c->rp;
  for (int k = 0; k < 4; k++) {
    d.rcmin[k] = p->rcmin[k];
    d.rcmax[k] = p->rcmax[k];
    d.rcmaxsq[k] = p->rcmax[k] * p->rcmax[k];                  // pair_rebomos.cpp:974-977
    d.rcw[k] = p->rcmax[k] - p->rcmin[k];
    d.Q[k] = p->Q[k];
    d.alpha[k] = p->alpha[k];
    d.A[k] = p->A[k];
    d.BIJc[k] = p->BIJc[k];
    d.Beta[k] = p->Beta[k];
    d.rcLJmin[k] = p->rcLJmin[k];
    d.rcLJmax[k] = p->rcLJmax[k];
    const double eps = p->epsilon[k], sig = p->sigma[k];
    d.sig95[k] = 0.95 * sig;
    d.lj_out_hi[k] = rsq_smallest_with_sqrt(d.rcLJmax[k], true);     // sqrt(rsq) >  rcLJmax
    d.lj_in_lo[k] = rsq_smallest_with_sqrt(d.rcLJmin[k], false);     // sqrt(rsq) >= rcLJmin
    d.lj_s95[k] = rsq_smallest_with_sqrt(d.sig95[k], false);         // sqrt(rsq) >= 0.95 sigma
    // init_one (pair_rebomos.cpp:262-265): powint(sigma,12), powint(sigma,6)
    double s2 = sig * sig, s4 = s2 * s2, s8 = s4 * s4;
    double s6 = s2 * s4, s12 = s4 * s8;
    d.lj1[k] = 48.0 * eps * s12;
    d.lj2[k] = 24.0 * eps * s6;
    d.lj3[k] = 4.0 * eps * s12;
    d.lj4[k] = 4.0 * eps * s6;
    // cubic taper (pair_rebomos.cpp:533-538)
    const double dr = 0.95 * sig - p->rcLJmin[k];
    const double q = sig / (0.95 * sig);
    const double q2 = q * q;
    const double r6 = q2 * (q2 * q2);    // powint(q,6): yy = ww^2 * ww^4
    const double vdw = 4 * eps * r6 * (r6 - 1.0);
    const double dvdw = (-4 * eps / (0.95 * sig)) * r6 * (12.0 * r6 - 6.0);
    d.c2[k] = ((3.0 / dr) * vdw - dvdw) / dr;
    d.c3[k] = (vdw / (dr * dr) - d.c2[k]) / dr;
  }
  for (int e = 0; e < 2; e++) {
    for (int o = 0; o < 7; o++) {
      d.b[e][o] = p->b[o][e];
      d.bg[e][o] = p->bg[o][e];
    }
    for (int o = 0; o < 4; o++) d.a[e][o] = p->a[o][e];
  }
}

static void set_margin(b200md_ctx *c)
{
  double m = (c->margin_opt > 0.0) ? c->margin_opt : 0.5 * c->skin;    // default: two-level list, inner skin = skin/2
  if (m > c->skin) m = c->skin;    // the master rows only cover cut + skin
  c->margin = m;
  for (int k = 0; k < 4; k++) {
    const double s = c->rp.rcmax[k] + m;
    const double l = c->rp.rcLJmax[k] + m;
    c->rp.shortsq[k] = s * s;
    c->rp.ljsq[k] = l * l;
  }
}

extern "C" int b200md_rebomos_init(b200md_ctx *c, const b200md_rebomos_params *p, int ntypes,
                                   const int *map)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, p && map && ntypes >= 1 && ntypes <= B200MD_MAX_TYPES, "rebomos_init: ntypes must be 1..8");
  for (int k = 0; k < 4; k++)
    ARG_CHECK(c, p->rcmax[k] > p->rcmin[k] && p->rcmin[k] > 0.0 && p->sigma[k] > 0.0,
              "rebomos_init: need 0 < rcmin < rcmax and sigma > 0");
  CUDA_TRY(c, cudaSetDevice(c->device));
  derive_params(c, p);
  c->ntypes = ntypes;
  c->map_h[0] = -1;
  for (int t = 1; t <= ntypes; t++) {
    ARG_CHECK(c, map[t] >= -1 && map[t] <= 1, "rebomos_init: map entries must be -1, 0 (Mo) or 1 (S)");
    c->map_h[t] = map[t];
  }
  CUDA_TRY(c, c->map_d.reserve(B200MD_MAX_TYPES + 1));
  CUDA_TRY(c, cudaMemcpyAsync(c->map_d.p, c->map_h, (ntypes + 1) * sizeof(int), cudaMemcpyHostToDevice,
                              c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->rebomos_ready = true;
  c->type_on_device = c->tag_on_device = false;
  c->aeam_ready = false;
  c->inner_valid = false;
  return B200MD_OK;
}

// pack positions (and fail on bad types)
int b200md_rebomos_pack(b200md_ctx *c)
{
  if (c->nall == 0) return B200MD_OK;
  LaunchScope ls(c, "pack");
  pack_xq_kernel<<<nblocks(c->nall, BLOCK), BLOCK, 0, c->stream>>>(c->x_aos.p, c->type.p, c->map_d.p,
                                                                   c->ntypes, 0, c->nall, c->xq.p, c->flags.p);
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// (re)build the inner lists (short rows + directed LJ rows) from the master list on the device
int b200md_rebomos_build_inner(b200md_ctx *c)
{
  const int rows = c->list_inum + c->list_gnum;
  const int inum = c->list_inum;
  set_margin(c);
  CUDA_TRY(c, c->short_idx.reserve((size_t) B200MD_SHORT_WIDTH * rows + 64));
  CUDA_TRY(c, c->short_num.reserve((size_t) rows + 32));
  const bool pairs = c->lj_pairs != 0;
  const int P = inum / 2 + 2;    // pair slots per element
  c->ljp_P = P;
  CUDA_TRY(c, c->lj_off.reserve((size_t) (pairs ? 2 * P : inum) + 2));
  CUDA_TRY(c, c->lj_num.reserve((size_t) (pairs ? 4 * P : 2 * inum) + 32));
  int rc = B200MD_OK;
  if (!pairs && (rc = b200md_exclusive_scan_i64(c, c->list_num.p, c->lj_off.p, inum, 8))) return rc;
  // capacity bound without a host round trip: every row padded to a multiple of 8
  const int64_t cap = c->list_entries_used + 8 * (int64_t) (pairs ? 2 * P : inum) + 64;
  CUDA_TRY(c, c->lj_val.reserve((size_t) cap));
  c->lj_capacity = cap;
  CUDA_TRY(c, c->xhold.reserve(4 * (size_t) c->nall + 8));
  CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 4, 0, 2 * sizeof(int), c->stream));
  if (rows > 0) {
    LaunchScope ls(c, "build_inner");
    build_inner_kernel<<<nblocks((long long) rows * 32, BLOCK), BLOCK, 0, c->stream>>>(
        c->rp, c->xq.p, c->list_off.p, c->list_num.p, c->list_val.p, rows, pairs ? 0 : inum,
        c->short_idx.p, c->short_num.p, c->lj_off.p, c->lj_num.p, c->lj_val.p, c->flags.p);
    CUDA_TRY(c, cudaGetLastError());
  }
  CUDA_TRY(c, cudaMemcpyAsync(c->xhold.p, c->xq.p, (size_t) c->nall * sizeof(double4),
                              cudaMemcpyDeviceToDevice, c->stream));
  // owned centers by element for the fused bond-order launches (flags[12], [13] = counts)
  CUDA_TRY(c, c->cen_list.reserve(3 * (size_t) inum + 96));
  CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 12, 0, 2 * sizeof(int), c->stream));
  CUDA_TRY(c, c->cen_key.reserve((size_t) inum + 32));
  CUDA_TRY(c, c->cen_scan.reserve((size_t) inum + 2));
  if (inum > 0) {
    {
      LaunchScope ls(c, "center_lists");
      center_key_kernel<<<nblocks(inum, BLOCK), BLOCK, 0, c->stream>>>(c->xq.p, inum, c->cen_key.p);
    }
    if ((rc = b200md_exclusive_scan_i64(c, c->cen_key.p, c->cen_scan.p, inum, 1))) return rc;
    LaunchScope ls(c, "center_lists");
    center_lists_kernel<<<nblocks(inum, BLOCK), BLOCK, 0, c->stream>>>(
        c->cen_key.p, (const long long *) c->cen_scan.p, inum, c->cen_list.p, c->cen_list.p + inum + 32,
        c->flags.p + 12);
    CUDA_TRY(c, cudaGetLastError());
  } else
    CUDA_TRY(c, cudaMemsetAsync(c->cen_scan.p, 0, sizeof(int64_t), c->stream));
  if (pairs) {
    // LJ rows of center PAIRS (consecutive entries of the ascending center lists): capacities, offsets, union rows
    CUDA_TRY(c, c->ljp_ab.reserve(4 * (size_t) P + 8));
    CUDA_TRY(c, c->scan_tmp.reserve(2 * (size_t) P + 8));
    {
      LaunchScope ls(c, "build_inner");
      ljpair_cap_kernel<<<nblocks(2 * P, BLOCK), BLOCK, 0, c->stream>>>(c->cen_list.p, inum, P, c->flags.p + 12,
                                                                      c->list_num.p, c->scan_tmp.p, (int2 *) c->ljp_ab.p);
    }
    c->split_valid = false;
    if (c->split.on) {
      // [interior | boundary] order of the pair slots: flags -> two scans -> permuted (a, b) and capacities
      CUDA_TRY(c, c->ljp_tmp.reserve(2 * (size_t) (4 * P) + 2 * (size_t) (2 * P) + 64));
      CUDA_TRY(c, c->ljp_scan.reserve(2 * ((size_t) 2 * P + 2) + 8));
      int *keyI = c->ljp_tmp.p, *keyB = keyI + 2 * P, *cap2 = keyB + 2 * P;
      int2 *ab2 = (int2 *) (cap2 + 2 * P + (2 * P & 1));
      int64_t *scanI = c->ljp_scan.p, *scanB = scanI + 2 * P + 2, *split = scanB + 2 * P + 2;
      {
        LaunchScope ls(c, "build_inner");
        ljpair_split_key_kernel<<<nblocks(2 * P, BLOCK), BLOCK, 0, c->stream>>>((const int2 *) c->ljp_ab.p, P, c->xq.p, c->split,
                                                                              keyI, keyB);
      }
      if ((rc = b200md_exclusive_scan_i64(c, keyI, scanI, 2 * P, 1))) return rc;
      if ((rc = b200md_exclusive_scan_i64(c, keyB, scanB, 2 * P, 1))) return rc;
      CUDA_TRY(c, cudaMemsetAsync(ab2, 0xff, 2 * (size_t) P * sizeof(int2), c->stream));
      CUDA_TRY(c, cudaMemsetAsync(cap2, 0, 2 * (size_t) P * sizeof(int), c->stream));
      {
        LaunchScope ls(c, "build_inner");
        ljpair_split_permute_kernel<<<nblocks(2 * P, BLOCK), BLOCK, 0, c->stream>>>(
            (const int2 *) c->ljp_ab.p, c->scan_tmp.p, P, keyI, keyB, (const long long *) scanI, (const long long *) scanB, ab2,
            cap2, (long long *) split);
      }
      CUDA_TRY(c, cudaMemcpyAsync(c->ljp_ab.p, ab2, 2 * (size_t) P * sizeof(int2), cudaMemcpyDeviceToDevice, c->stream));
      CUDA_TRY(c, cudaMemcpyAsync(c->scan_tmp.p, cap2, 2 * (size_t) P * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
      c->ljp_split = split;
      c->split_valid = true;
    }
    if ((rc = b200md_exclusive_scan_i64(c, c->scan_tmp.p, c->lj_off.p, 2 * P, 8))) return rc;
    LaunchScope ls(c, "build_inner");
    build_ljpair_kernel<<<nblocks((long long) 2 * P * 32, LJP_BLOCK), LJP_BLOCK, 0, c->stream>>>(
        c->rp, c->xq.p, c->list_off.p, c->list_num.p, c->list_val.p, P, (const int2 *) c->ljp_ab.p, c->lj_off.p,
        c->lj_num.p, c->lj_val.p, c->flags.p);
    CUDA_TRY(c, cudaGetLastError());
  }
  c->h2d_ready = false;    // the upload dependencies below belong to these lists (plugin mode computes them on demand)
  c->tight_valid = false;  // ... and so do the tight rows (the resident loop derives them next)
  c->inner_valid = true;
  c->n_inner_rebuild++;
  return B200MD_OK;
}

int b200md_ensure_halo_stream(b200md_ctx *c);    // system.cu

// derive (or re-derive) the tight rows from the wide rows at the positions now on the device; GPU-resident loop only
int b200md_rebomos_derive_tight(b200md_ctx *c)
{
  c->tight_valid = false;
  if (!c->inner_valid || !c->lj_pairs || c->deterministic) return B200MD_OK;
  const double m = c->margin_t_opt;
  if (m <= 0.0 || m >= c->margin) return B200MD_OK;
  const int inum = c->list_inum, P = c->ljp_P;
  if (inum == 0) return B200MD_OK;
  c->margin_t = m;
  TightLimits lim;
  for (int k = 0; k < 4; k++) {
    const double a = c->rp.rcmax[k] + m, b = c->rp.rcLJmax[k] + m;
    lim.shortsq[k] = a * a;
    lim.ljsq[k] = b * b;
  }
  CUDA_TRY(c, c->short_idx_t.reserve((size_t) B200MD_SHORT_WIDTH * inum + 64));
  CUDA_TRY(c, c->short_num_t.reserve((size_t) inum + 32));
  CUDA_TRY(c, c->lj_val_t.reserve((size_t) c->lj_capacity));
  CUDA_TRY(c, c->lj_num_t.reserve(4 * (size_t) P + 32));
  CUDA_TRY(c, c->xhold_t.reserve(4 * (size_t) c->nall + 8));
  // the two passes are independent and neither fills the machine (gather latency): the short rows go to the side stream
  const bool side = !c->sync_timing && b200md_ensure_halo_stream(c) == B200MD_OK;
  cudaStream_t s_short = c->stream;
  if (side) {
    CUDA_TRY(c, cudaEventRecord(c->ev_ready, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->halo_stream, c->ev_ready, 0));
    s_short = c->halo_stream;
  }
  {
    LaunchScope ls(c, "derive_tight_short");
    derive_tight_short_kernel<<<nblocks((long long) inum * 8, BLOCK), BLOCK, 0, s_short>>>(
        c->xq.p, c->short_idx.p, c->short_num.p, inum, lim, c->short_idx_t.p, c->short_num_t.p);
  }
  if (side) CUDA_TRY(c, cudaEventRecord(c->ev_fwd, c->halo_stream));
  {
    LaunchScope ls(c, "derive_tight_lj");
    CUDA_TRY(c, cudaMemsetAsync(c->lj_num_t.p, 0, 4 * (size_t) P * sizeof(int), c->stream));    // slots without a pair
    const int nwarps = inum / 2 + 2;    // pairs of both elements together: ceil(nMo / 2) + ceil(nS / 2) <= inum / 2 + 1
#define DTL_ARGS c->xq.p, c->lj_off.p, c->lj_num.p, (const int2 *) c->ljp_ab.p, c->lj_val.p, P, c->flags.p + 12, lim, c->lj_val_t.p, c->lj_num_t.p
    // 4 trips in flight at 64 registers: 0.646 ms per 1 M atoms; 2 trips at 56 registers 0.675, at 40 (spills) 0.683
    derive_tight_lj_kernel<4, 4><<<nblocks((long long) nwarps * 32, BLOCK), BLOCK, 0, c->stream>>>(DTL_ARGS);
  }
  if (side) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_fwd, 0));
  CUDA_TRY(c, cudaGetLastError());
  CUDA_TRY(c, cudaMemcpyAsync(c->xhold_t.p, c->xq.p, (size_t) c->nall * sizeof(double4), cudaMemcpyDeviceToDevice,
                              c->stream));
  c->tight_valid = true;
  c->n_tight++;
  return B200MD_OK;
}

// upload pipelining (plugin mode): which piece of the position upload each range of centers has to wait for.  Computed
// from the short rows the first time a plugin-mode call finds fresh inner lists (one small kernel + one host read).
static int rebomos_prepare_pipeline(b200md_ctx *c)
{
  const int inum = c->list_inum;
  c->h2d_ready = false;
  if (c->h2d_chunks <= 1 || inum < c->d2h_min_atoms || inum == 0) return B200MD_OK;
  const int K = c->h2d_chunks;
  ChunkBounds cb;
  cb.K = K;
  // option "h2d_ramp": piece k is (ramp + k) / sum(ramp + j) of the atoms -- small pieces first (the first kernels start
  // sooner), big ones last (fewer, longer launches once the upload is ahead of the kernels); 0 = equal pieces
  {
    const double r = c->h2d_ramp > 0 ? (double) c->h2d_ramp : 0.0;
    double tot = 0.0, acc = 0.0;
    for (int k = 0; k < K; k++) tot += r > 0.0 ? r + k : 1.0;
    cb.t[0] = 0;
    for (int k = 0; k < K; k++) {
      acc += r > 0.0 ? r + k : 1.0;
      cb.t[k + 1] = (int) ((double) inum * (acc / tot));
    }
    cb.t[K] = inum;
  }
  int *dep = c->flags.p + 16;    // flags holds 16 + B200MD_MAX_D2H_CHUNKS ints
  int *pin = (int *) (c->pin_scal.p + 48);
  int *pin_cnt = (int *) (c->pin_scal.p + 60);
  const int cap = inum / 32 + 1024;    // more stragglers than this: the atom order is not spatial, nothing to gain
  c->n_strag = 0;
  CUDA_TRY(c, c->strag_flag.reserve((size_t) inum + 8));
  CUDA_TRY(c, c->strag_list.reserve((size_t) cap + 8));
  for (int pass = 0; pass < 2; pass++) {    // pass 0 with the straggler list, pass 1 (only if it overflowed) without
    CUDA_TRY(c, cudaMemsetAsync(dep, 0, K * sizeof(int), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->strag_flag.p, 0, ((size_t) inum + 1) * sizeof(int), c->stream));
    int *cnt = c->strag_flag.p + inum;
    {
      LaunchScope ls(c, "build_inner");
      dep_range_kernel<<<nblocks(inum, BLOCK), BLOCK, 0, c->stream>>>(c->short_idx.p, c->short_num.p, inum, cb, dep,
                                                                     pass == 0 ? c->strag_flag.p : nullptr, c->strag_list.p,
                                                                     cnt, cap);
    }
    CUDA_TRY(c, cudaMemcpyAsync(pin, dep, K * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(pin_cnt, cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (pass == 0 && *pin_cnt <= cap) {
      c->n_strag = *pin_cnt;
      break;
    }
  }
  if (c->n_strag) {
    c->strag_host.resize((size_t) c->n_strag);
    CUDA_TRY(c, cudaMemcpyAsync(c->strag_host.data(), c->strag_list.p, c->n_strag * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, c->strag_pin.reserve(3 * (size_t) c->n_strag + 8));
    CUDA_TRY(c, c->strag_dev.reserve(3 * (size_t) c->n_strag + 8));
  }
  for (int k = 0; k < K; k++) {
    int p = k;    // a range always needs its own piece
    while (p + 1 < K && pin[k] >= cb.t[p + 1]) p++;
    c->h2d_need[k] = p;
    c->h2d_t[k] = cb.t[k];
  }
  c->h2d_t[K] = inum;
  c->h2d_K = K;
  c->h2d_ready = true;
  return B200MD_OK;
}

// decide whether the inner lists are still valid for the positions now on the device
int b200md_rebomos_refresh_inner(b200md_ctx *c)
{
  const int rows = c->list_inum + c->list_gnum;
  ARG_CHECK(c, rows <= c->nall, "neighbor list has more rows than atoms");
  if (!c->inner_valid) return b200md_rebomos_build_inner(c);
  if (c->margin >= c->skin) return B200MD_OK;    // master-list owner (LAMMPS) enforces skin/2 itself
  const double half = 0.5 * c->margin;
  {
    LaunchScope ls(c, "check_disp");
    check_disp_kernel<<<nblocks(c->nall, BLOCK), BLOCK, 0, c->stream>>>(
        c->xq.p, (const double4 *) c->xhold.p, c->nall, half * half, c->flags.p);
    CUDA_TRY(c, cudaGetLastError());
  }
  int flag = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&flag, c->flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (flag) {
    CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 1, 0, sizeof(int), c->stream));
    return b200md_rebomos_build_inner(c);
  }
  return B200MD_OK;
}

// force kernels on whatever is resident: xq, inner lists.  f and scal must be zeroed by the caller.
template <bool EV, bool DET, bool ATOM>
static void launch_centers(b200md_ctx *c, const DetTables &det, int t_lo, int t_hi, bool overflow_pass)
{
  const int inum = c->list_inum;
  int *list0 = c->cen_list.p, *list1 = c->cen_list.p + inum + 32, *ovf = c->cen_list.p + 2 * ((size_t) inum + 32);
  int *cnt0 = c->flags.p + 12, *cnt1 = c->flags.p + 13, *cntO = c->flags.p + 14;
  const long long *scan = (const long long *) c->cen_scan.p;
  // Mo centers: 16 lanes, 16 staged bonds; S centers: 4 lanes, overflow to 16/16.
  // Grids cover the worst case (every atom of the range in one class); surplus groups see g >= count and leave.
  const int nr = t_hi - t_lo;
  const int grid0 = min(nblocks((long long) nr * 16, 128), c->num_sms * 48);
  const int grid1 = min(nblocks((long long) nr * 4, 128), c->num_sms * 48);
  const bool tight = c->tight_valid && !DET;
  const int *sidx = tight ? c->short_idx_t.p : c->short_idx.p, *snum = tight ? c->short_num_t.p : c->short_num.p;
#define RC_ARGS(list, cnt, sc, ol, oc) \
  c->rp, c->xq.p, sidx, snum, list, cnt, sc, t_lo, t_hi, ol, oc, c->f.p, det, b200md_scal_arg(c), c->flags.p, c->pa_e, c->pa_v
  // occupancy (r01 v6 sweeps at 995 904 atoms): force-only Mo launch 80 registers (6 CTAs/SM); force-only S launch stages
  // 4 bonds per center (bulk S has 3; more go to the overflow launch) which cuts its shared memory from 45 to 17 KB, and
  // runs at 72 registers (7 CTAs/SM): 0.257 -> 0.210 ms; at 64 registers 0.213, at 80: 0.223, unbounded (104): 0.280
  constexpr bool PLAIN = !EV && !DET && !ATOM;
  constexpr int MB_MO = PLAIN ? 6 : 5, MB_S = PLAIN ? 7 : 5, CAP_S = PLAIN ? 4 : 8;
  if (nr > 0) {
    {
      LaunchScope ls(c, PLAIN ? "rebo_center_mo" : "rebo_center_mo_ev");
      rebo_center_kernel<128, 16, 16, 0, EV, DET, ATOM, MB_MO><<<grid0, 128, 0, c->stream>>>(RC_ARGS(list0, cnt0, scan, nullptr, nullptr));
    }
    {
      LaunchScope ls(c, PLAIN ? "rebo_center_s" : "rebo_center_s_ev");
      rebo_center_kernel<128, 4, CAP_S, 1, EV, DET, ATOM, MB_S><<<grid1, 128, 0, c->stream>>>(RC_ARGS(list1, cnt1, scan, ovf, cntO));
    }
  }
  if (overflow_pass) {
    LaunchScope ls(c, "rebo_center_overflow");
    rebo_center_kernel<128, 16, 16, 1, EV, DET, ATOM, MB_MO><<<c->num_sms * 2, 128, 0, c->stream>>>(RC_ARGS(ovf, cntO, nullptr, nullptr, nullptr));
  }
}

// many-body part (REBO_neigh + FREBO + bondorder, fdotr virial) on c->stream, for the centers with atom index in
// [t_lo, t_hi).  first: set up this call's tables; last: overflow launch, deterministic gather, fdotr.
static int rebomos_forces_manybody(b200md_ctx *c, int eflag, int vflag, int t_lo, int t_hi, bool first, bool last)
{
  const int ncen = c->list_inum;
  const int rows = c->list_inum + c->list_gnum;
  if (first) CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 14, 0, sizeof(int), c->stream));
  if (ncen == 0) return B200MD_OK;
  DetTables det = {nullptr, nullptr, nullptr, nullptr};
  const bool detmode = c->deterministic != 0;
  if (detmode) {
    ARG_CHECK(c, c->list_gnum > 0 || c->nghost == 0,
              "deterministic mode needs the ghost rows of the neighbor list (REQ_GHOST / ghost_rows = 1)");
    CUDA_TRY(c, c->det_fb.reserve(3 * (size_t) ncen * (B200MD_MAX_REBO + 1) + 32));
    CUDA_TRY(c, c->det_j.reserve((size_t) ncen * (B200MD_MAX_REBO + 1) + 32));
    det.fb = c->det_fb.p;
    det.fi = c->det_fb.p + 3 * (size_t) ncen * B200MD_MAX_REBO;
    det.j = c->det_j.p;
    det.nb = c->det_j.p + (size_t) ncen * B200MD_MAX_REBO;
    if (first) {
      CUDA_TRY(c, cudaMemsetAsync(det.nb, 0, ncen * sizeof(int), c->stream));
      CUDA_TRY(c, cudaMemsetAsync(det.fi, 0, 3 * (size_t) ncen * sizeof(double), c->stream));
    }
  }
  const bool ev = eflag != 0;
  const bool atom = c->pa_e != nullptr;
  ARG_CHECK(c, !(atom && detmode), "per-atom energy/virial is not available in deterministic mode");
  if (atom) launch_centers<true, false, true>(c, det, t_lo, t_hi, last);
  else if (detmode) {
    if (ev) launch_centers<true, true, false>(c, det, t_lo, t_hi, last);
    else launch_centers<false, true, false>(c, det, t_lo, t_hi, last);
    if (last) {
      LaunchScope ls(c, "rebo_gather");
      rebo_gather_kernel<<<nblocks(rows, BLOCK), BLOCK, 0, c->stream>>>(c->short_idx.p, c->short_num.p, rows, ncen, det,
                                                                       c->f.p);
    }
  } else {
    if (ev) launch_centers<true, false, false>(c, det, t_lo, t_hi, last);
    else launch_centers<false, false, false>(c, det, t_lo, t_hi, last);
  }
  if (vflag && last) {
    LaunchScope ls(c, "fdotr");
    fdotr_kernel<<<c->num_sms * 4, BLOCK, 0, c->stream>>>(c->xq.p, c->f.p, c->nall, b200md_scal_arg(c));
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// tapered LJ for the owned atoms with index in [t_lo, t_hi) on c->stream: completes f of exactly those atoms
static int rebomos_forces_lj(b200md_ctx *c, int eflag, int vflag, int t_lo, int t_hi, int part = -1, int elems = 3)
{
  const int ncen = c->list_inum;
  if (t_hi <= t_lo) return B200MD_OK;
  const bool atom = c->pa_e != nullptr;
  int *list0 = c->cen_list.p, *list1 = c->cen_list.p + ncen + 32;
  // grids: 8 lanes per center (pair of centers), capped (grid-stride loops); list pieces are located on the device
  if (c->lj_pairs) {
    // split order (halo overlap): part 0 = interior pair rows, 1 = boundary rows, -1 = all; ljp_split plays cen_scan's role
    const long long ngroups = (c->split_valid ? ncen : t_hi - t_lo) / 2 + 2;
    const int *ljnum = c->tight_valid ? c->lj_num_t.p : c->lj_num.p, *ljval = c->tight_valid ? c->lj_val_t.p : c->lj_val.p;
    const long long *ljscan = (const long long *) c->cen_scan.p;
    if (c->split_valid) {
      ljscan = (const long long *) c->ljp_split;
      t_lo = part == 1 ? 1 : 0;
      t_hi = part == 0 ? 1 : 2;
    }
#define LJP_ARGS \
  c->rp, c->xq.p, c->lj_off.p, ljnum, (const int2 *) c->ljp_ab.p, ljval, c->ljp_P, ljscan, t_lo, t_hi, c->f.p, b200md_scal_arg(c), \
      c->pa_e, c->pa_v
#define LJP_LAUNCH(EVF, E, MB, AT, NTH) \
  lj_pair_kernel<EVF, E, 2, MB, AT, NTH><<<min(nblocks(ngroups * 8, NTH), c->num_sms * 64 * (256 / NTH)), NTH, 0, c->stream>>>(LJP_ARGS)
    // force-only instance: 2 position buffers per lane, 80 registers (3 CTAs/SM).  r01 v6 sweep at 995 904 atoms: D=2/80 regs
    // 0.634 ms; D=1 0.77; D=3/80 (spills) 0.85; D=4/118 regs (2 CTAs) 0.69; D=2/64 regs (spills) 0.78
    // CTAs of 128 threads at 80 registers (6 CTAs/SM): 0.605 ms; 256 threads x 3 CTAs: 0.619; 72 regs x 7 CTAs (spills): 0.683
#define LJP_FORCE(E) LJP_LAUNCH(false, E, 6, false, 128);
    const bool lj_ev = atom || eflag || vflag;    // thermo steps run the energy/virial instances: their own names
    if (elems & 1) {
      LaunchScope ls(c, lj_ev ? "lj_mo_ev" : "lj_mo");
      if (atom) LJP_LAUNCH(true, 0, 1, true, 256);
      else if (eflag || vflag) LJP_LAUNCH(true, 0, 2, false, 256);
      else LJP_FORCE(0)
    }
    if (elems & 2) {
      LaunchScope ls(c, lj_ev ? "lj_s_ev" : "lj_s");
      if (atom) LJP_LAUNCH(true, 1, 1, true, 256);
      else if (eflag || vflag) LJP_LAUNCH(true, 1, 2, false, 256);
      else LJP_FORCE(1)
    }
    CUDA_TRY(c, cudaGetLastError());
    return B200MD_OK;
  }
  // one row per center (lj_pairs = 0, the r01 v3 kernel).  Occupancy decided there too: 2 candidates in flight per lane at
  // 64 registers (4 CTAs/SM) 0.87 ms; 3 at 80: 0.94; 4 at 96 (2 CTAs): 1.10; 48 or 40 registers spill (1.4, 1.7 ms)
  const int grid = min(nblocks((long long) (t_hi - t_lo) * 8, BLOCK), c->num_sms * 64);
#define LJ_ARGS(list) \
  c->rp, c->xq.p, c->lj_off.p, c->lj_num.p, c->lj_val.p, list, (const long long *) c->cen_scan.p, t_lo, t_hi, c->f.p, \
      b200md_scal_arg(c), c->pa_e, c->pa_v
  {
    LaunchScope ls(c, "lj_mo");
    if (atom) lj_kernel<true, 0, 2, 2, true><<<grid, BLOCK, 0, c->stream>>>(LJ_ARGS(list0));
    else if (eflag || vflag) lj_kernel<true, 0, 2, 2, false><<<grid, BLOCK, 0, c->stream>>>(LJ_ARGS(list0));
    else lj_kernel<false, 0, 2, 4, false><<<grid, BLOCK, 0, c->stream>>>(LJ_ARGS(list0));
  }
  {
    LaunchScope ls(c, "lj_s");
    if (atom) lj_kernel<true, 1, 2, 2, true><<<grid, BLOCK, 0, c->stream>>>(LJ_ARGS(list1));
    else if (eflag || vflag) lj_kernel<true, 1, 2, 2, false><<<grid, BLOCK, 0, c->stream>>>(LJ_ARGS(list1));
    else lj_kernel<false, 1, 2, 4, false><<<grid, BLOCK, 0, c->stream>>>(LJ_ARGS(list1));
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

int b200md_rebomos_forces(b200md_ctx *c, int eflag, int vflag)
{
  int rc;
  if ((rc = rebomos_forces_manybody(c, eflag, vflag, 0, c->list_inum, true, true))) return rc;
  return rebomos_forces_lj(c, eflag, vflag, 0, c->list_inum);
}

// resident loop with halo overlap, force-only steps.  which = 0: the bond-order launches (all centers); which = 1: the LJ
// launches over the interior (part 0: no ghost among the candidates) or the boundary (part 1) pair rows
int b200md_rebomos_forces_part(b200md_ctx *c, int part, int which)
{
  ARG_CHECK(c, c->split_valid && c->lj_pairs, "rebomos_forces_part: the pair rows are not in split order");
  if (which == 0) return rebomos_forces_manybody(c, 0, 0, 0, c->list_inum, true, true);
  if (c->split_elems == 2) {
    // only the S rows are split (the larger launch): the interior S rows hide the forward halo, the Mo rows -- one launch,
    // all of them -- and the boundary S rows run beside the reverse halo: three LJ launches instead of four
    if (part == 0) return rebomos_forces_lj(c, 0, 0, 0, c->list_inum, 0, 2);
    int rc = rebomos_forces_lj(c, 0, 0, 0, c->list_inum, -1, 1);
    if (rc) return rc;
    return rebomos_forces_lj(c, 0, 0, 0, c->list_inum, 1, 2);
  }
  return rebomos_forces_lj(c, 0, 0, 0, c->list_inum, part);
}

static int check_flags(b200md_ctx *c, const int *fl)
{
  if (fl[4] || fl[5]) {
    c->n_short_entries = fl[4];
    c->n_lj_entries = fl[5];
  }
  if (fl[3]) {
    c->fail("atom type outside 1..ntypes");
    return B200MD_ERR_ARG;
  }
  if (fl[0]) {
    c->fail(fl[0] == 2 ? "REBO bond table overflow" : "REBO neighbor row overflow (more than " +
                std::to_string(B200MD_MAX_REBO) + " REBO neighbors or " +
                std::to_string(B200MD_SHORT_WIDTH) + " short-row entries on one atom)");
    cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream);
    return B200MD_ERR_OVERFLOW;
  }
  return B200MD_OK;
}

// Plugin-mode compute with the position upload, the force kernels and the force download all overlapped:
//   upload stream : ghosts | piece 0 | piece 1 | ... (each packed into xq as it lands), displacement check, its flag
//   compute stream: bond-order launches of center range k as soon as the piece holding the largest index it reads
//                   has arrived (h2d_need, computed when the inner lists were built); then the LJ ranges
//   copy stream   : ghost forces after the last bond-order launch, owned forces range by range behind the LJ launches
// The inner lists are used SPECULATIVELY: whether an atom has moved more than margin/2 since they were derived is only
// known once all positions are on the device.  If so (rare: every ~50+ steps at 300 K) the lists are re-derived and
// the forces recomputed by the plain path; nothing of the speculative pass has been added to the caller's f by then.
static int rebomos_compute_pipelined(b200md_ctx *c, int nlocal, int nghost, const double *x, int eflag, int vflag,
                                     double *f, double *eng_vdwl, double *virial, int *fl, bool *redo, int *redo_level)
{
  *redo = false;
  *redo_level = 0;
  const int K = c->h2d_K;
  const int nall = nlocal + nghost;
  if (!c->up_stream) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
  for (int k = 0; k <= K; k++)
    if (!c->up_ev[k]) CUDA_TRY(c, cudaEventCreateWithFlags(&c->up_ev[k], cudaEventDisableTiming));
  c->nlocal = nlocal;
  c->nghost = nghost;
  c->nall = nall;
  const bool tight = c->tight_valid;
  const bool need_check = c->margin < c->skin || tight;
  int *pin_flag = (int *) (c->pin_scal.p + 56);
  pin_flag[0] = pin_flag[1] = 0;
  // the tight rows may still be being re-derived from the previous call's positions (deferred derive, below): the
  // upload must not overwrite xq before that is done
  if (c->tight_derive_pending) {
    CUDA_TRY(c, cudaStreamWaitEvent(c->up_stream, c->ev_tight, 0));
    c->tight_derive_pending = false;
  }
  // ---- upload stream
  cudaStream_t compute_stream = c->stream;
  c->stream = c->up_stream;    // LaunchScope and the helpers below enqueue on c->stream
  int rc = B200MD_OK;
  auto piece = [&](int lo, int hi) -> int {
    if (hi <= lo) return B200MD_OK;
    CUDA_TRY(c, cudaMemcpyAsync(c->x_aos.p + 3 * (size_t) lo, x + 3 * (size_t) lo, 3 * (size_t) (hi - lo) * sizeof(double),
                                cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long) (3 * (size_t) (hi - lo) * sizeof(double));
    LaunchScope ls(c, "pack");
    pack_xq_kernel<<<nblocks(hi - lo, BLOCK), BLOCK, 0, c->stream>>>(c->x_aos.p, c->type.p, c->map_d.p, c->ntypes, lo,
                                                                    hi, c->xq.p, c->flags.p);
    return B200MD_OK;
  };
  if (c->n_strag) {    // stragglers first (see dep_range_kernel): a few KB gathered on the host
    const int ns = c->n_strag;
    double *sp = c->strag_pin.p;
    for (int q = 0; q < ns; q++) {
      const size_t j = (size_t) c->strag_host[q];
      sp[3 * (size_t) q] = x[3 * j];
      sp[3 * (size_t) q + 1] = x[3 * j + 1];
      sp[3 * (size_t) q + 2] = x[3 * j + 2];
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->strag_dev.p, sp, 3 * (size_t) ns * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long) (3 * (size_t) ns * sizeof(double));
    LaunchScope ls(c, "pack");
    strag_scatter_kernel<<<nblocks(ns, BLOCK), BLOCK, 0, c->stream>>>(c->strag_dev.p, c->strag_list.p, ns, c->type.p,
                                                                     c->map_d.p, c->ntypes, c->xq.p);
  }
  rc = piece(nlocal, nall);
  for (int p = 0; p < K && !rc; p++) {
    rc = piece(c->h2d_t[p], c->h2d_t[p + 1]);
    if (!rc && cudaEventRecord(c->up_ev[p], c->stream) != cudaSuccess) rc = B200MD_ERR_CUDA;
  }
  if (!rc && need_check) {
    const double half = c->margin < c->skin ? 0.5 * c->margin : 1.0e10;
    const double th = 0.5 * c->margin_t, soon = 0.8 * th;
    {
      LaunchScope ls(c, "check_disp");
      check_disp_kernel<<<nblocks(nall, BLOCK), BLOCK, 0, c->stream>>>(c->xq.p, (const double4 *) c->xhold.p, nall,
                                                                      half * half, c->flags.p,
                                                                      tight ? (const double4 *) c->xhold_t.p : nullptr,
                                                                      th * th, soon * soon);
    }
    if (cudaMemcpyAsync(pin_flag, c->flags.p + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      rc = B200MD_ERR_CUDA;
  }
  if (!rc && cudaEventRecord(c->up_ev[K], c->stream) != cudaSuccess) rc = B200MD_ERR_CUDA;
  c->stream = compute_stream;
  if (rc) return rc;
  // ---- compute stream
  const size_t n3 = 3 * (size_t) nall;
  CUDA_TRY(c, cudaMemsetAsync(c->f.p, 0, (n3 + 8) * sizeof(double), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), c->stream));
  for (int k = 0; k < K; k++) {
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->up_ev[c->h2d_need[k]], 0));
    if ((rc = rebomos_forces_manybody(c, eflag, vflag, c->h2d_t[k], c->h2d_t[k + 1], k == 0, k == K - 1))) return rc;
  }
  if ((rc = b200md_d2h_begin(c))) return rc;
  if ((rc = b200md_d2h_range(c, 0, f, 3 * (size_t) nlocal, n3))) return rc;
  const int KD = c->d2h_chunks;
  for (int k = 0; k < KD; k++) {
    const int t0 = (int) ((long long) nlocal * k / KD), t1 = (int) ((long long) nlocal * (k + 1) / KD);
    if ((rc = rebomos_forces_lj(c, eflag, vflag, t0, t1))) return rc;
    if ((rc = b200md_d2h_range(c, k + 1, f, 3 * (size_t) t0, 3 * (size_t) t1))) return rc;
  }
  // ---- the verdict on the lists arrives while the kernels are still running
  CUDA_TRY(c, cudaEventSynchronize(c->up_ev[K]));
  if (need_check && (pin_flag[0] || pin_flag[1] == 2)) {
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    b200md_collect_timers(c);
    CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 1, 0, 2 * sizeof(int), c->stream));
    *redo = true;
    *redo_level = pin_flag[0] ? 2 : 1;    // 2: re-derive the inner rows; 1: only the tight rows were overtaken
    return B200MD_OK;
  }
  rc = b200md_d2h_finish(c, eflag, vflag, f, eng_vdwl, virial, fl);
  if (!rc && need_check && pin_flag[1] == 1) {
    // an atom is 80 % of the way to the tight rows' limit: re-derive them now, after this call's work, from the positions
    // on the device -- the host integrates in the meantime, the next call's upload waits for the event
    CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 2, 0, sizeof(int), c->stream));
    if ((rc = b200md_rebomos_derive_tight(c))) return rc;
    if (!c->ev_tight) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_tight, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventRecord(c->ev_tight, c->stream));
    c->tight_derive_pending = true;
  }
  return rc;
}

extern "C" int b200md_rebomos_compute(b200md_ctx *c, int nlocal, int nghost, const double *x,
                                      const int *type, const int *tag, int eflag, int vflag, double *f,
                                      double *eng_vdwl, double *virial)
{
  return b200md_rebomos_compute_peratom(c, nlocal, nghost, x, type, tag, eflag, vflag, f, eng_vdwl, virial, nullptr,
                                        nullptr);
}

extern "C" int b200md_rebomos_compute_peratom(b200md_ctx *c, int nlocal, int nghost, const double *x,
                                              const int *type, const int *tag, int eflag, int vflag, double *f,
                                              double *eng_vdwl, double *virial, double *eatom, double *vatom)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->rebomos_ready, "rebomos_compute: call b200md_rebomos_init first");
  ARG_CHECK(c, c->list_valid, "rebomos_compute: no neighbor list (b200md_set_neighbor_list / b200md_neigh_build)");
  ARG_CHECK(c, c->list_inum == nlocal, "rebomos_compute: neighbor list was built for a different nlocal");
  ARG_CHECK(c, f != nullptr || nlocal + nghost == 0, "rebomos_compute: f is NULL");
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->n_compute++;
  if (nlocal + nghost == 0) {    // an empty rank (vacuum brick): nothing to upload, nothing to add
    if (eng_vdwl) *eng_vdwl = 0.0;
    if (virial)
      for (int k = 0; k < 6; k++) virial[k] = 0.0;
    return B200MD_OK;
  }
  // plugin mode keeps tight rows of its own (derived after every inner-list build, checked with the positions of every
  // call, re-derived one call ahead of need); a resident system's tight rows follow another schedule
  if (c->sys_owns_tight) {
    c->tight_valid = false;
    c->sys_owns_tight = false;
  }
  if (c->split.on || c->split_valid) {    // ... and so does the interior/boundary order of the pair rows
    c->split.on = 0;
    if (c->split_valid) c->inner_valid = false;
  }
  int rc;
  int fl[16];
  bool redo = false;
  if (!eatom && !vatom && !c->deterministic && c->inner_valid && c->h2d_ready && c->d2h_chunks > 1 &&
      c->type_on_device && (c->tag_on_device || !tag) && nlocal + nghost == c->ids_nall && c->nlocal == nlocal) {
    int level = 0;
    if ((rc = rebomos_compute_pipelined(c, nlocal, nghost, x, eflag, vflag, f, eng_vdwl, virial, fl, &redo, &level))) return rc;
    c->n_pipelined++;
    if (!redo) return check_flags(c, fl);
    c->n_redo++;
    // an atom had moved beyond the inner lists' margin (level 2) or beyond the tight rows' (level 1): positions are on
    // the device, re-derive and recompute
    if (level == 2 && (rc = b200md_rebomos_build_inner(c))) return rc;
    if ((rc = b200md_rebomos_derive_tight(c))) return rc;
  } else {
    if (c->tight_derive_pending) {
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      c->tight_derive_pending = false;
    }
    if ((rc = b200md_upload_atoms(c, nlocal, nghost, x, type, tag))) return rc;
    if ((rc = b200md_rebomos_pack(c))) return rc;
    if ((rc = b200md_rebomos_refresh_inner(c))) return rc;
    if ((rc = b200md_rebomos_derive_tight(c))) return rc;    // tight rows at these positions (0.5 ms per 1 M atoms)
  }
  if (!c->h2d_ready && !eatom && !vatom && !c->deterministic && (rc = rebomos_prepare_pipeline(c))) return rc;
  const size_t n3 = 3 * (size_t) c->nall;
  CUDA_TRY(c, cudaMemsetAsync(c->f.p, 0, (n3 + 8) * sizeof(double), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), c->stream));
  if ((rc = b200md_peratom_begin(c, eatom != nullptr || vatom != nullptr))) return rc;
  if (!eatom && !vatom && c->d2h_chunks > 1 && nlocal >= c->d2h_min_atoms) {
    // Forces go home while the LJ kernels are still running: the many-body launches scatter into owned and ghost
    // entries, so they run first; ghost forces are then final and travel during the first LJ range, and every LJ
    // range [t_k, t_k+1) completes f of exactly those owned atoms, which travel during the next range.
    if ((rc = rebomos_forces_manybody(c, eflag, vflag, 0, nlocal, true, true))) return rc;
    if ((rc = b200md_d2h_begin(c))) return rc;
    const int K = c->d2h_chunks;
    if ((rc = b200md_d2h_range(c, 0, f, 3 * (size_t) nlocal, 3 * (size_t) c->nall))) return rc;
    for (int k = 0; k < K; k++) {
      const int t0 = (int) ((long long) nlocal * k / K), t1 = (int) ((long long) nlocal * (k + 1) / K);
      if ((rc = rebomos_forces_lj(c, eflag, vflag, t0, t1))) return rc;
      if ((rc = b200md_d2h_range(c, k + 1, f, 3 * (size_t) t0, 3 * (size_t) t1))) return rc;
    }
    if ((rc = b200md_d2h_finish(c, eflag, vflag, f, eng_vdwl, virial, fl))) return rc;
    return check_flags(c, fl);
  }
  rc = b200md_rebomos_forces(c, (eatom || vatom) ? 1 : eflag, vflag);
  if (rc) {
    c->pa_e = c->pa_v = nullptr;
    return rc;
  }
  if ((rc = b200md_peratom_finish(c, eatom, vatom))) return rc;
  if ((rc = b200md_finish_compute(c, eflag, vflag, f, eng_vdwl, virial, fl))) return rc;
  // virial[0..5] = fdotr (xx,yy,zz,xy,xz,yz) of the many-body part + LJ pair virial in the same order
  if ((rc = check_flags(c, fl))) return rc;
  return B200MD_OK;
}

extern "C" int b200md_rebomos_neigh(b200md_ctx *c, int nlocal, int nghost, const double *x, const int *type,
                                    int stride, int *rebo_numneigh, int *rebo_rows, double *nM, double *nS)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->rebomos_ready && c->list_valid, "rebomos_neigh: init and neighbor list required");
  ARG_CHECK(c, c->list_inum == nlocal, "rebomos_neigh: neighbor list was built for a different nlocal");
  ARG_CHECK(c, stride >= 1 && rebo_numneigh && rebo_rows && nM && nS, "rebomos_neigh: NULL output");
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rc = b200md_upload_atoms(c, nlocal, nghost, x, type, nullptr);
  if (rc) return rc;
  if ((rc = b200md_rebomos_pack(c))) return rc;
  if ((rc = b200md_rebomos_refresh_inner(c))) return rc;
  const int rows = c->list_inum + c->list_gnum;    // ghost rows exist only if the list carries them
  DevBuf<int> d_num, d_rows;
  CUDA_TRY(c, d_num.reserve((size_t) rows + 8));
  CUDA_TRY(c, d_rows.reserve((size_t) rows * stride + 8));
  CUDA_TRY(c, c->nM.reserve((size_t) c->nall + 32));
  CUDA_TRY(c, c->nS.reserve((size_t) c->nall + 32));
  if (rows) {
    LaunchScope ls(c, "rebo_rows");
    rebo_rows_kernel<<<nblocks(rows, BLOCK), BLOCK, 0, c->stream>>>(
        c->rp, c->xq.p, c->short_idx.p, c->short_num.p, rows, stride, d_num.p, d_rows.p,
        c->nM.p, c->nS.p, c->flags.p);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(rebo_numneigh, d_num.p, rows * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(rebo_rows, d_rows.p, (size_t) rows * stride * sizeof(int),
                                cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(nM, c->nM.p, rows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(nS, c->nS.p, rows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  int fl[16];
  CUDA_TRY(c, cudaMemcpyAsync(fl, c->flags.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  d_num.release();
  d_rows.release();
  return check_flags(c, fl);
}
