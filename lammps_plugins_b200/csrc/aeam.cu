// b200md -- angular EAM (pair_style aeam) force path on sm_100a.
//
// Reference semantics: lammps-plugins USER-AEAM/pair_aeam.cpp
//   density pass :164-253, embedding :264-303, force pass :309-476,
//   file2array :752-872, array2spline :876-911, interpolate :915-942.
//
// Formulation (DESIGN.md "AEAM kernels"):
//   A1 aeam_density      8 lanes / owned non-angular atom: rho_i = sum_j f_ij(r), row streamed as sectors
//   A2 aeam_density_ang  warp / owned angular atom: rho_i = sum_{j<k} 2 f_ij f_ik (cos+1/3)^2
//   A3 aeam_embed        thread / owned atom: fp_i = F'(rho^ni), energy F
//   (fp of ghosts: halo exchange -- live here, numerically dead in the reference)
//   B1 aeam_force        8 lanes / owned atom, GATHER form: both directed visits (i->j and j->i) of a
//                        pair are evaluated from i's side using fp_j, so the pair term needs no atomics
//   B2 aeam_force_ang    warp / owned angular atom: 3-body forces, FP64 atomics on j,k (0.75 % of atoms)
// Spline rows live on the device as {c3,c4,c5,c6} (32 B = one sector per lookup); the derivative
// coefficients c0..c2 of the reference are exactly 3c3/d, 2c4/d, c5/d (pair_aeam.cpp:937-941).

#include "common.cuh"

#include <cmath>

#define BLOCK 256
#define ANG_CAP 96    // staged neighbors per angular center

// ================================================================== host: tables
// PairAEAM::interpolate (pair_aeam.cpp:915-942), 1-based rows, 7 coefficients
static void interpolate7(int n, double delta, const double *f /*1-based*/, double *spline /*(n+1)*7*/)
{
  auto S = [&](int m, int k) -> double & { return spline[(size_t) m * 7 + k]; };
  for (int m = 1; m <= n; m++) S(m, 6) = f[m];
  S(1, 5) = S(2, 6) - S(1, 6);
  S(2, 5) = 0.5 * (S(3, 6) - S(1, 6));
  S(n - 1, 5) = 0.5 * (S(n, 6) - S(n - 2, 6));
  S(n, 5) = S(n, 6) - S(n - 1, 6);
  for (int m = 3; m <= n - 2; m++)
    S(m, 5) = ((S(m - 2, 6) - S(m + 2, 6)) + 8.0 * (S(m + 1, 6) - S(m - 1, 6))) / 12.0;
  for (int m = 1; m <= n - 1; m++) {
    S(m, 4) = 3.0 * (S(m + 1, 6) - S(m, 6)) - 2.0 * S(m, 5) - S(m + 1, 5);
    S(m, 3) = S(m, 5) + S(m + 1, 5) - 2.0 * (S(m + 1, 6) - S(m, 6));
  }
  S(n, 4) = 0.0;
  S(n, 3) = 0.0;
  for (int m = 1; m <= n; m++) {
    S(m, 2) = S(m, 5) / delta;
    S(m, 1) = 2.0 * S(m, 4) / delta;
    S(m, 0) = 3.0 * S(m, 3) / delta;
  }
  for (int k = 0; k < 7; k++) S(0, k) = 0.0;
}

struct AeamHost {
  int nel = 0;
  std::vector<std::vector<double>> frho7, rhor7, z2r7;    // 7-coefficient tables (reference layout)
  std::vector<int> frho_n, rhor_n, z2r_n;
};

static int pack4(const std::vector<double> &s7, int n, std::vector<double> &out)
{
  int off = (int) (out.size() / 4);
  for (int m = 0; m <= n; m++)
    for (int k = 3; k <= 6; k++) out.push_back(s7[(size_t) m * 7 + k]);
  return off;
}

extern "C" int b200md_aeam_init(b200md_ctx *c, const b200md_aeam_tables *t)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, t && t->nelements >= 1 && t->nelements <= 4, "aeam_init: nelements must be 1..4");
  ARG_CHECK(c, t->nnonangular >= 0 && t->nnonangular <= t->nelements, "aeam_init: bad nnonangular");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int nel = t->nelements;
  if (!c->aeam_host) c->aeam_host = new AeamHost();
  AeamHost &H = *c->aeam_host;
  H = AeamHost();
  H.nel = nel;
  AeamDev &d = c->ap;
  memset(&d, 0, sizeof(d));
  d.nel = nel;
  d.nnonangular = t->nnonangular;

  std::vector<double> p_frho, p_rhor, p_z2r;
  // frho: one table per element (the reference's extra all-zero table serves pair hybrid only)
  for (int i = 0; i < nel; i++) {
    const int n = t->nrho[i];
    ARG_CHECK(c, n >= 5 && t->drho[i] > 0.0, "aeam_init: need nrho >= 5 and drho > 0");
    std::vector<double> f1((size_t) n + 1, 0.0);
    for (int m = 1; m <= n; m++) f1[m] = t->frho[i][m - 1];
    std::vector<double> s7((size_t) (n + 1) * 7);
    interpolate7(n, t->drho[i], f1.data(), s7.data());
    d.nrho[i] = n;
    d.rdrho[i] = 1.0 / t->drho[i];
    d.frho_off[i] = pack4(s7, n, p_frho);
    H.frho7.push_back(std::move(s7));
    H.frho_n.push_back(n);
  }
  // rhor: full matrix, table index i*nel+j (type2rhor, pair_aeam.cpp:816-821)
  for (int i = 0; i < nel; i++)
    for (int j = 0; j < nel; j++) {
      const int ij = i * nel + j;
      const int n = t->nr[ij];
      ARG_CHECK(c, n >= 5 && t->dr[ij] > 0.0 && t->cut[ij] > 0.0, "aeam_init: need nr >= 5, dr > 0, cut > 0");
      std::vector<double> f1((size_t) n + 1, 0.0);
      for (int m = 1; m <= n; m++) f1[m] = t->rhor[ij][m - 1];
      std::vector<double> s7((size_t) (n + 1) * 7);
      interpolate7(n, t->dr[ij], f1.data(), s7.data());
      d.nr[ij] = n;
      d.rdr[ij] = 1.0 / t->dr[ij];
      d.cut[ij] = t->cut[ij];
      d.rhor_off[ij] = pack4(s7, n, p_rhor);
      H.rhor7.push_back(std::move(s7));
      H.rhor_n.push_back(n);
    }
  // z2r: lower triangle (j <= i), interpolated with nr/dr of [i][j] (file2array :836-843);
  // pair (a,b) uses table (max,min) (type2z2r :853-871)
  std::vector<int> tri_off((size_t) nel * nel, 0);
  for (int i = 0; i < nel; i++)
    for (int j = 0; j <= i; j++) {
      const int ij = i * nel + j;
      const int n = t->nr[ij];
      std::vector<double> f1((size_t) n + 1, 0.0);
      for (int m = 1; m <= n; m++) f1[m] = t->z2r[ij][m - 1];
      std::vector<double> s7((size_t) (n + 1) * 7);
      interpolate7(n, t->dr[ij], f1.data(), s7.data());
      tri_off[ij] = pack4(s7, n, p_z2r);
      H.z2r7.push_back(std::move(s7));
      H.z2r_n.push_back(n);
    }
  for (int i = 0; i < nel; i++)
    for (int j = 0; j < nel; j++) {
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      d.z2r_off[i * nel + j] = tri_off[hi * nel + lo];
    }

  // fused per-pair rows for the density and force kernels: {rhor(i,j) c3..c6 | z2r(i,j) c3..c6} = 64 bytes,
  // one gather per neighbor.  Both halves are indexed with the SAME row m computed from nr/dr of (i,j), exactly
  // as the reference does (pair_aeam.cpp:352-372).
  std::vector<double> p_pair;
  for (int i = 0; i < nel; i++)
    for (int j = 0; j < nel; j++) {
      const int ij = i * nel + j;
      const int n = d.nr[ij];
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      const int ntri = t->nr[hi * nel + lo];
      d.pair_off[ij] = (int) (p_pair.size() / 8);
      d.z2r_n[ij] = ntri;
      d.z2r_k[ij] = t->dr[ij] / t->dr[hi * nel + lo];
      if (d.z2r_k[ij] != 1.0) d.asym_dr = 1;
      for (int m = 0; m <= n; m++) {
        for (int k = 0; k < 4; k++) p_pair.push_back(p_rhor[4 * ((size_t) d.rhor_off[ij] + m) + k]);
        const int mz = m < ntri ? m : ntri;
        for (int k = 0; k < 4; k++) p_pair.push_back(p_z2r[4 * ((size_t) d.z2r_off[ij] + mz) + k]);
      }
      // exact rsq form of the reference's `r > cut` (sqrt is monotone and correctly rounded)
      double x = t->cut[ij] * t->cut[ij];
      while (sqrt(x) > t->cut[ij]) x = nextafter(x, 0.0);
      while (!(sqrt(x) > t->cut[ij])) x = nextafter(x, 1.0e300);
      d.cut_gt_sq[ij] = x;
    }
  CUDA_TRY(c, c->spl_pair.reserve(p_pair.size() + 8));
  CUDA_TRY(c, cudaMemcpyAsync(c->spl_pair.p, p_pair.data(), p_pair.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, c->spl_frho.reserve(p_frho.size() + 8));
  CUDA_TRY(c, c->spl_rhor.reserve(p_rhor.size() + 8));
  CUDA_TRY(c, c->spl_z2r.reserve(p_z2r.size() + 8));
  CUDA_TRY(c, cudaMemcpyAsync(c->spl_frho.p, p_frho.data(), p_frho.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(c->spl_rhor.p, p_rhor.data(), p_rhor.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(c->spl_z2r.p, p_z2r.data(), p_z2r.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->ntypes = nel;
  c->aeam_ready = true;
  c->type_on_device = c->tag_on_device = false;
  c->rebomos_ready = false;
  c->inner_valid = false;
  return B200MD_OK;
}

extern "C" int b200md_aeam_get_spline(b200md_ctx *c, int kind, int index, double *out, int nrows)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->aeam_ready && out, "aeam_get_spline: call b200md_aeam_init first");
  ARG_CHECK(c, c->aeam_host, "aeam_get_spline: call b200md_aeam_init first");
  AeamHost &H = *c->aeam_host;
  const std::vector<std::vector<double>> *tab = kind == 0 ? &H.frho7 : kind == 1 ? &H.rhor7 : kind == 2 ? &H.z2r7 : nullptr;
  const std::vector<int> *ns = kind == 0 ? &H.frho_n : kind == 1 ? &H.rhor_n : &H.z2r_n;
  ARG_CHECK(c, tab && index >= 0 && index < (int) tab->size(), "aeam_get_spline: bad table");
  ARG_CHECK(c, nrows == (*ns)[index], "aeam_get_spline: nrows must equal the table length");
  memcpy(out, (*tab)[index].data(), (size_t) (nrows + 1) * 7 * sizeof(double));
  return B200MD_OK;
}

void b200md_aeam_forget(b200md_ctx *c)
{
  delete c->aeam_host;
  c->aeam_host = nullptr;
}

// ================================================================== device helpers
// xq.w of the AEAM path carries TWO things: the 0-based element in the two lowest mantissa bits and, in the
// remaining bits, the gated embedding derivative (1-del)*Fptmp*F'(rho) of the atom (aeam_gate_kernel), so the
// force kernel gets position, element and fp_j of a neighbor from ONE 32-byte sector.  Overwriting two mantissa
// bits perturbs fp by <= 3 ulp (7e-16 relative).
__device__ __forceinline__ int etype(const double4 &q) { return (int) (__double_as_longlong(q.w) & 3LL); }
__device__ __forceinline__ double w_encode(double g, int t)
{
  return __longlong_as_double((__double_as_longlong(g) & ~3LL) | (long long) t);
}
__device__ __forceinline__ double w_gate(const double4 &q) { return __longlong_as_double(__double_as_longlong(q.w) & ~3LL); }

__device__ __forceinline__ void spl_index(double x, double rdx, int n, int &m, double &p)
{
  p = x * rdx + 1.0;
  m = (int) p;
  m = min(m, n - 1);
  p -= m;
  p = fmin(p, 1.0);
}
__device__ __forceinline__ double spl_val(const double4 &c, double p)
{
  return ((c.x * p + c.y) * p + c.z) * p + c.w;
}
__device__ __forceinline__ double spl_der(const double4 &c, double p, double rdx)
{
  return ((3.0 * c.x * p + 2.0 * c.y) * p + c.z) * rdx;
}

// positions + 0-based element packed into one 32-byte sector
__global__ void __launch_bounds__(BLOCK) aeam_pack_kernel(const double *__restrict__ x,
                                                          const int *__restrict__ type, int nel, int nall,
                                                          double4 *__restrict__ xq, int *__restrict__ flags, int lo = 0)
{
  int i = lo + blockIdx.x * BLOCK + threadIdx.x;    // atoms [lo, nall)
  if (i >= nall) return;
  int t = type[i];
  if (t < 1 || t > nel) {
    flags[3] = 1;
    t = 1;
  }
  xq[i] = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], w_encode(0.0, t - 1));
}

// inner rows: master row filtered to r <= max(cut_ij, cut_ji) + margin, 8-aligned, order preserved
__global__ void __launch_bounds__(BLOCK) aeam_build_inner_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq,
    const int64_t *__restrict__ list_off, const int *__restrict__ list_num,
    const int *__restrict__ list_val, int inum, const int64_t *__restrict__ ea_off,
    int *__restrict__ ea_num, int *__restrict__ ea_val, int *__restrict__ ang_list, int *__restrict__ flags)
{
  const int lane = threadIdx.x & 31;
  const int i = (int) (((size_t) blockIdx.x * BLOCK + threadIdx.x) >> 5);
  if (i >= inum) return;
  const double4 xi = xq[i];
  const int ti = etype(xi);
  const int n = list_num[i];
  const int64_t base = list_off[i], obase = ea_off[i];
  const unsigned lt = (1u << lane) - 1u;
  int no = 0;
  // 4 trips of 32 entries in flight per lane (index -> position -> test is a dependent chain: one trip at a time the
  // kernel ran at the latency of 2 x 4 round trips per row, 1.8 ms per 2 M rows)
  constexpr int NU = 4;
  for (int e0 = 0; e0 < n; e0 += 32 * NU) {
    int jv[NU];
    double4 xv[NU];
#pragma unroll
    for (int u = 0; u < NU; u++) {
      const int e = e0 + u * 32 + lane;
      jv[u] = (e < n) ? (ld_stream_int(list_val + base + e) & B200MD_NEIGHMASK) : -1;
    }
#pragma unroll
    for (int u = 0; u < NU; u++)
      if (jv[u] >= 0) xv[u] = ld_sector(xq + jv[u]);
#pragma unroll
    for (int u = 0; u < NU; u++) {
      bool keep = false;
      if (jv[u] >= 0) {
        const double dx = xv[u].x - xi.x, dy = xv[u].y - xi.y, dz = xv[u].z - xi.z;
        keep = dx * dx + dy * dy + dz * dz <= par.cutsq_list[ti * par.nel + etype(xv[u])];
      }
      const unsigned mk = __ballot_sync(0xffffffffu, keep);
      if (keep) ea_val[obase + no + __popc(mk & lt)] = jv[u];
      no += __popc(mk);
    }
  }
  if (lane == 0) {
    ea_num[i] = no;
    atomicAdd(&flags[5], no);
  }
}

// per-type-pair constants staged in shared memory: indexing the kernel-parameter bank with a per-lane pair
// type serialises in the address-divergence unit (ncu r01, lj v1: pipe_adu 66 %); LDS with mostly equal
// addresses is a broadcast
struct PairPar {
  double cutgt, rdr;
  int nr, off;
  int roff, zoff, nz, pad;    // rows of the separate rhor / z2r tables, z2r row clamp
};
// The z2r table of an unlike pair is built with dr[max][min] (pair_aeam.cpp:836-843, 906-910) but indexed with m, p from
// dr[i][j] (:352-372): its derivative carries 1/dr[max][min], i.e. rdr[i][j] * zk with zk = dr[i][j] / dr[max][min] -- 1
// for like pairs and for files with symmetric dr, such as AlSi.aeam.  The hot loops run at their register limit (a few
// bytes of spill cost 5-20 % of a kernel), so the factor exists only in the ZK instances of the round-1 force kernel,
// which a file with asymmetric dr is routed to (AeamDev::asym_dr).
__device__ __forceinline__ double z_scale(const AeamDev &par, int pair) { return par.z2r_k[pair]; }
__device__ __forceinline__ void load_pair_par(const AeamDev &par, PairPar *sp, bool rhor_table = false)
{
  if (threadIdx.x < 16) {
    sp[threadIdx.x].cutgt = par.cut_gt_sq[threadIdx.x];
    sp[threadIdx.x].rdr = par.rdr[threadIdx.x];
    sp[threadIdx.x].nr = par.nr[threadIdx.x];
    sp[threadIdx.x].off = rhor_table ? par.rhor_off[threadIdx.x] : par.pair_off[threadIdx.x];
    sp[threadIdx.x].roff = par.rhor_off[threadIdx.x];
    sp[threadIdx.x].zoff = par.z2r_off[threadIdx.x];
    sp[threadIdx.x].nz = par.z2r_n[threadIdx.x];
    sp[threadIdx.x].pad = 0;
  }
  __syncthreads();
}

__device__ __forceinline__ void st_sector(double4 *p, double a, double b, double c, double d)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ double4 ld_stream_sector(const double4 *p)
{
  double4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p));
  return r;
}

template <int CPL> struct DfVec;
template <> struct DfVec<4> {
  static __device__ __forceinline__ void st(double *p, const double *o) { st_sector((double4 *) p, o[0], o[1], o[2], o[3]); }
  static __device__ __forceinline__ void ld(const double *p, double *o)
  {
    const double4 v = ld_stream_sector((const double4 *) p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct DfVec<2> {
  static __device__ __forceinline__ void st(double *p, const double *o)
  {
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(o[0]), "d"(o[1]) : "memory");
  }
  static __device__ __forceinline__ void ld(const double *p, double *o)
  {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "l"(p));
  }
};
template <> struct DfVec<1> {
  static __device__ __forceinline__ void st(double *p, const double *o)
  {
    asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(o[0]) : "memory");
  }
  static __device__ __forceinline__ void ld(const double *p, double *o)
  {
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(o[0]) : "l"(p));
  }
};

// ================================================================== A1: density of non-angular atoms
// Register-count sensitive (check -Xptxas -v after every change): the DF instance compiles to 64 registers without spill
// = 4 CTAs per SM, 1.27 ms at 2 M atoms; at 72 registers (3 CTAs) it takes 1.50 ms, forced to 64 with 12 bytes of
// spill 1.35 ms
// RANGED (plugin mode): centers [first, inum) -- an instance of its own, the resident loop's sits at its register limit.
template <bool DF, bool RANGED = false>
__global__ void __launch_bounds__(BLOCK) aeam_density_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ ea_off,
    const int *__restrict__ ea_num, const int *__restrict__ ea_val, const double4 *__restrict__ rhor,
    int inum, double *__restrict__ rho, double *__restrict__ ea_df, int first = 0)
{
  __shared__ PairPar sp[16];
  load_pair_par(par, sp, true);
  const int tid = blockIdx.x * BLOCK + threadIdx.x;
  const int i = (RANGED ? first : 0) + (tid >> 3), sub = tid & 7;
  double acc = 0.0;
  bool mine = false;
  if (i < inum) {
    const double4 xi = xq[i];
    const int ti = etype(xi);
    mine = ti < par.nnonangular;
    if (mine) {
      const int n = ea_num[i];
      const int *row = ea_val + ea_off[i];
      double *dfrow = DF ? ea_df + ea_off[i] : nullptr;    // f'_{ti,tj}(r) per entry, for the force pass
      const int tbase = ti * par.nel;
      for (int e0 = 0; e0 < n; e0 += 32) {
        int jj[4];
        double4 xj[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int e = e0 + u * 8 + sub;
          jj[u] = (e < n) ? ld_stream_int(row + e) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (jj[u] >= 0) xj[u] = ld_sector(xq + jj[u]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (jj[u] < 0) continue;
          const double dx = xj[u].x - xi.x, dy = xj[u].y - xi.y, dz = xj[u].z - xi.z;
          const double rsq = dx * dx + dy * dy + dz * dz;
          const PairPar pp = sp[tbase + etype(xj[u])];
          double dfv = 0.0;
          if (rsq < pp.cutgt) {    // else r > cut; i non-angular: CutDec = 0 (pair_aeam.cpp:187-194)
            const double r1 = sqrt(rsq);
            int m;
            double p;
            spl_index(r1, pp.rdr, pp.nr, m, p);
            const double4 cf = ld_sector(rhor + pp.off + m);
            acc += spl_val(cf, p);
            if (DF) dfv = spl_der(cf, p, pp.rdr);
          }
          if (DF) DfVec<1>::st(dfrow + e0 + u * 8 + sub, &dfv);
        }
      }
    }
  }
  acc = group_sum<8>(acc);
  if (mine && sub == 0) rho[i] = acc;
}


// ================================================================== cluster form (default)
// CL consecutive owned centers (a "cluster") share ONE union row: every gathered candidate position is tested against
// all CL centers, so a step gathers ~38 instead of ~81 position sectors per atom (fcc Al, cut + margin = 7 A), and the
// density pass stores f'_{ti,tj}(r) of every (entry, center) slot next to the row (one 32-byte store per entry), so the
// force pass reads it back as a coalesced stream instead of gathering the rhor row a second time: per in-range
// (center, candidate) the two passes gather 2 spline sectors instead of 3.  Gathers are the bound of these kernels
// (tools/microbench/gather.cu: 0.87 random sectors per cycle and SM, whatever the width), FP64 is at 10 %.
#define CL 4
#define CL_SHIFT 2
#define CL_HT 1024
#define CL_BLOCK 128

// capacity of a cluster's union row: the sum of its centers' master rows (upper bound)
__global__ void __launch_bounds__(BLOCK) aeam_cluster_cap_kernel(const int *__restrict__ list_num, int inum, int ncl,
                                                                 int *__restrict__ cap)
{
  const int q = blockIdx.x * BLOCK + threadIdx.x;
  if (q >= ncl) return;
  int s = 0;
#pragma unroll
  for (int c = 0; c < CL; c++)
    if (CL * q + c < inum) s += list_num[CL * q + c];
  cap[q] = s;
}

// one warp per cluster: union of the centers' master-row entries within max(cut_ij, cut_ji) + margin of THAT center,
// duplicates dropped through a per-warp hash set, order of first appearance (optionally sorted by atom index)
__global__ void __launch_bounds__(CL_BLOCK) aeam_build_cluster_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ list_off,
    const int *__restrict__ list_num, const int *__restrict__ list_val, int inum, int ncl,
    const int64_t *__restrict__ ec_off, int *__restrict__ ec_num, int *__restrict__ ec_val, int *__restrict__ ang_list,
    int *__restrict__ flags, int sort_rows)
{
  __shared__ int s_tab[CL_BLOCK / 32][CL_HT];
  const int lane = threadIdx.x & 31;
  const int q = (int) (((size_t) blockIdx.x * CL_BLOCK + threadIdx.x) >> 5);
  if (q >= ncl) return;
  int *tab = s_tab[threadIdx.x >> 5];
  for (int k = lane; k < CL_HT; k += 32) tab[k] = -1;
  __syncwarp();
  const int64_t base = ec_off[q];
  const unsigned lt = (1u << lane) - 1u;
  int no = 0;
  bool full = false;
  for (int cc = 0; cc < CL && !full; cc++) {
    const int i = CL * q + cc;
    if (i >= inum) break;
    const double4 xi = xq[i];
    const int ti = etype(xi);
    const int n = list_num[i];
    const int64_t mb = list_off[i];
    for (int e0 = 0; e0 < n; e0 += 32) {
      const int e = e0 + lane;
      bool keep = false;
      int j = 0;
      if (e < n) {
        j = ld_stream_int(list_val + mb + e) & B200MD_NEIGHMASK;
        const double4 xj = xq[j];
        const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
        keep = dx * dx + dy * dy + dz * dz <= par.cutsq_list[ti * par.nel + etype(xj)];
      }
      if (keep) {    // find or insert
        unsigned h = ((unsigned) j * 2654435761u) >> 22;    // 10 bits
        for (;;) {
          const int v = atomicCAS(&tab[h], -1, j);
          if (v == -1) break;
          if (v == j) {
            keep = false;
            break;
          }
          h = (h + 1) & (CL_HT - 1);
        }
      }
      const unsigned mk = __ballot_sync(0xffffffffu, keep);
      if (keep) ec_val[base + no + __popc(mk & lt)] = j;
      no += __popc(mk);
      if (no > CL_HT * 3 / 4 - 32) {    // warp-uniform: the hash set would fill up
        if (lane == 0) flags[0] = 4;
        full = true;
        break;
      }
    }
    __syncwarp();
  }
  if (sort_rows && !full && no > 1) {
    // bitonic sort of the row by atom index in the (now free) hash table: adjacent lanes of the compute kernels then
    // read adjacent sectors, which share 128-byte lines
    __syncwarp();
    int np = 32;
    while (np < no) np <<= 1;
    for (int k = lane; k < np; k += 32) tab[k] = k < no ? ec_val[base + k] : 0x7fffffff;
    __syncwarp();
    for (int kk = 2; kk <= np; kk <<= 1)
      for (int jj = kk >> 1; jj > 0; jj >>= 1) {
        for (int t = lane; t < np; t += 32) {
          const int u = t ^ jj;
          if (u > t) {
            const int a = tab[t], b = tab[u];
            const bool up = (t & kk) == 0;
            if ((a > b) == up) {
              tab[t] = b;
              tab[u] = a;
            }
          }
        }
        __syncwarp();
      }
    for (int k = lane; k < no; k += 32) ec_val[base + k] = tab[k];
  }
  if (lane == 0) {
    ec_num[q] = no;
    atomicAdd(&flags[5], no);
  }
}

// Lane layout of the cluster kernels: a cluster is served by 8 * LPE lanes; in one trip it takes 8 row entries, each
// by LPE adjacent lanes ("parts") that gather the SAME position sector (the L1 serves the duplicates at register
// write-back speed, tools/microbench/gather.cu) and test it against CPL = CL / LPE centers each.  LPE = 1 keeps all
// four centers in one lane (fewest gathered lanes, 128 registers, 14 resident warps: latency-bound, r02 ncu v1);
// LPE = 2 / 4 trade duplicate-lane gathers for half / a quarter of the per-lane state and 2-3x the resident warps.
// sum over the lanes of a cluster group that hold the same part (lanes l, l + LPE, l + 2 LPE, ...: 8 of them)
template <int LPE> __device__ __forceinline__ double part_sum(double v)
{
  constexpr int G = 8 * LPE;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned) (G - 1)));
#pragma unroll
  for (int o = 4 * LPE; o >= LPE; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// A1 (cluster form): density of the non-angular centers of a cluster.  Also writes f'_{ti,tj}(r) of every
// (entry, center) slot (0 where the density pass takes nothing) for the force pass.
template <int LPE, int U, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) aeam_density_cl_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ ec_off,
    const int *__restrict__ ec_num, const int *__restrict__ ec_val, const double4 *__restrict__ rhor, int inum,
    double *__restrict__ rho, double *__restrict__ ec_df)
{
  constexpr int CPL = CL / LPE, G = 8 * LPE;
  __shared__ PairPar sp[16];
  load_pair_par(par, sp);
  const int tid = blockIdx.x * BLOCK + threadIdx.x;
  const int q = tid / G, gl = tid % G;
  const int slot = gl / LPE, part = gl % LPE;
  const int i0 = q * CL, c0 = part * CPL;
  double acc[CPL], cx[CPL], cy[CPL], cz[CPL];
  int tb[CPL];    // ti * nel of a center that takes a density here, else -1 (angular center, or beyond inum)
#pragma unroll
  for (int c = 0; c < CPL; c++) {
    acc[c] = 0.0;
    cx[c] = cy[c] = cz[c] = 0.0;
    tb[c] = -1;
  }
  if (i0 < inum) {
#pragma unroll
    for (int c = 0; c < CPL; c++)
      if (i0 + c0 + c < inum) {
        const double4 xi = xq[i0 + c0 + c];
        cx[c] = xi.x;
        cy[c] = xi.y;
        cz[c] = xi.z;
        const int ti = etype(xi);
        tb[c] = ti < par.nnonangular ? ti * par.nel : -1;
      }
    const int n = ec_num[q];
    const int64_t base = ec_off[q];
    const int *row = ec_val + base;
    double *df = ec_df + CL * base + c0;
    for (int e0 = 0; e0 < n; e0 += 8 * U) {
      int jj[U];
      double4 xj[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int e = e0 + u * 8 + slot;
        jj[u] = (e < n) ? ld_stream_int(row + e) : -1;
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        if (jj[u] >= 0) xj[u] = ld_sector(xq + jj[u]);
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (jj[u] < 0) continue;
        const int tj = etype(xj[u]);
        bool hit[CPL];
        double pp[CPL], rd[CPL];
        const double4 *ra[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++) {
          const double dx = xj[u].x - cx[c], dy = xj[u].y - cy[c], dz = xj[u].z - cz[c];
          const double rsq = dx * dx + dy * dy + dz * dz;
          const PairPar &P = sp[(tb[c] < 0 ? 0 : tb[c]) + tj];
          // r > cut -> out; i non-angular: CutDec = 0 (pair_aeam.cpp:187-194)
          hit[c] = tb[c] >= 0 && jj[u] != i0 + c0 + c && rsq < P.cutgt;
          int m = 0;
          pp[c] = 0.0;
          if (CPL == 1 ? hit[c] : true) spl_index(rsq * rsqrt_nr(fmax(rsq, 1.0e-300)), P.rdr, P.nr, m, pp[c]);
          rd[c] = P.rdr;
          ra[c] = rhor + P.roff + m;
        }
        double4 rw[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++)
          if (hit[c]) rw[c] = ld_sector(ra[c]);
        double o[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++) {
          o[c] = 0.0;
          if (hit[c]) {
            acc[c] += spl_val(rw[c], pp[c]);
            o[c] = spl_der(rw[c], pp[c], rd[c]);
          }
        }
        DfVec<CPL>::st(df + (size_t) CL * (e0 + u * 8 + slot), o);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CPL; c++) acc[c] = part_sum<LPE>(acc[c]);
  if (slot == 0) {
#pragma unroll
    for (int c = 0; c < CPL; c++)
      if (tb[c] >= 0) rho[i0 + c0 + c] = acc[c];
  }
}

// B1 (cluster form): pair + embedding forces in gather form.  f'_{ij} comes from the density pass (ec_df); the phi
// row is the only spline gather of a same-element pair.
template <bool EV, bool ATOM, int LPE, int U, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) aeam_force_cl_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ ec_off,
    const int *__restrict__ ec_num, const int *__restrict__ ec_val, const double *__restrict__ ec_df,
    const double4 *__restrict__ rhor, const double4 *__restrict__ z2r, int inum, double *__restrict__ f,
    double *__restrict__ scal, double *__restrict__ pa_e, double *__restrict__ pa_v)
{
  constexpr int CPL = CL / LPE, G = 8 * LPE;
  __shared__ PairPar sp[16];
  load_pair_par(par, sp);
  const int tid = blockIdx.x * BLOCK + threadIdx.x;
  const int q = tid / G, gl = tid % G;
  const int slot = gl / LPE, part = gl % LPE;
  const int i0 = q * CL, c0 = part * CPL;
  const int nel = par.nel;
  double fx[CPL], fy[CPL], fz[CPL], cx[CPL], cy[CPL], cz[CPL], gi[CPL];
  int ti[CPL];
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  double ea[ATOM ? CPL : 1];
  double av[ATOM ? CPL : 1][6];
#pragma unroll
  for (int c = 0; c < CPL; c++) {
    fx[c] = fy[c] = fz[c] = cx[c] = cy[c] = cz[c] = gi[c] = 0.0;
    ti[c] = -1;
  }
  if (ATOM) {
#pragma unroll
    for (int c = 0; c < CPL; c++) {
      ea[ATOM ? c : 0] = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) av[ATOM ? c : 0][k] = 0.0;
    }
  }
  if (i0 < inum) {
#pragma unroll
    for (int c = 0; c < CPL; c++)
      if (i0 + c0 + c < inum) {
        const double4 xi = xq[i0 + c0 + c];
        cx[c] = xi.x;
        cy[c] = xi.y;
        cz[c] = xi.z;
        ti[c] = etype(xi);
        gi[c] = w_gate(xi);
      }
    const int n = ec_num[q];
    const int64_t base = ec_off[q];
    const int *row = ec_val + base;
    const double *df = ec_df + CL * base + c0;
    for (int e0 = 0; e0 < n; e0 += 8 * U) {
      int jj[U];
      double4 xj[U];
      double dv[U][CPL];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int e = e0 + u * 8 + slot;
        jj[u] = (e < n) ? ld_stream_int(row + e) : -1;
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        if (jj[u] >= 0) {
          xj[u] = ld_sector(xq + jj[u]);
          DfVec<CPL>::ld(df + (size_t) CL * (e0 + u * 8 + slot), dv[u]);
        }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (jj[u] < 0) continue;
        const int tj = etype(xj[u]);
        const double gj = w_gate(xj[u]);
        // at most two centers at a time -- phase 1: geometry and the phi row; phase 2: the gathers; phase 3: the terms
        constexpr int HB = CPL < 2 ? CPL : 2;
#pragma unroll
        for (int h = 0; h < CPL; h += HB) {
          bool in_ij[HB];
          double rinv[HB], pp[HB];
          int mz[HB];
          double4 zw[HB];
#pragma unroll
          for (int k = 0; k < HB; k++) {
            const int c = h + k;
            const double dx = xj[u].x - cx[c], dy = xj[u].y - cy[c], dz = xj[u].z - cz[c];
            const double rsq = dx * dx + dy * dy + dz * dz;
            const PairPar &P = sp[(ti[c] < 0 ? 0 : ti[c]) * nel + tj];
            in_ij[k] = ti[c] >= 0 && jj[u] != i0 + c0 + c && rsq < P.cutgt;    // !(r > cut[ti][tj])
            rinv[k] = rsqrt_nr(fmax(rsq, 1.0e-300));
            int m;
            spl_index(rsq * rinv[k], P.rdr, P.nr, m, pp[k]);
            mz[k] = P.zoff + min(m, P.nz);
          }
#pragma unroll
          for (int k = 0; k < HB; k++)
            if (in_ij[k]) zw[k] = ld_sector(z2r + mz[k]);
#pragma unroll
          for (int k = 0; k < HB; k++) {
            const int c = h + k;
            if (ti[c] < 0 || jj[u] == i0 + c0 + c) continue;
            const bool same = (tj == ti[c]);
            if (same && !in_ij[k]) continue;
            const double dx = xj[u].x - cx[c], dy = xj[u].y - cy[c], dz = xj[u].z - cz[c];
            const double recip = rinv[k];
            double coef = 0.0;    // fpair of visit (i,j) + fpair of visit (j,i)
            if (in_ij[k]) {
              // visit (i,j): pair_aeam.cpp:350-393
              const PairPar &P = sp[ti[c] * nel + tj];
              const double dfij = dv[u][c];
              const double phip = spl_der(zw[k], pp[k], P.rdr);    // symmetric dr only (asym_dr -> round-1 kernel)
              const double fpair = -gi[c] * dfij * recip + 0.5 * (-phip * recip);
              coef = fpair;
              // same element: visit (j,i) evaluates the same two splines at the same (m, p)
              if (same) coef += -gj * dfij * recip + 0.5 * (-phip * recip);
              if (EV) {
                const double ph = 0.5 * spl_val(zw[k], pp[k]);
                ev[0] += ph;
                ev[1] += dx * dx * fpair;
                ev[2] += dy * dy * fpair;
                ev[3] += dz * dz * fpair;
                ev[4] += dx * dy * fpair;
                ev[5] += dx * dz * fpair;
                ev[6] += dy * dz * fpair;
                if (ATOM) ea[ATOM ? c : 0] += ph;
              }
            }
            if (!same) {
              const PairPar &Q = sp[tj * nel + ti[c]];
              const double rsq = dx * dx + dy * dy + dz * dz;
              if (rsq < Q.cutgt) {
                // visit (j,i), evaluated here instead of scattering from j's row (different elements: own tables)
                int m;
                double p;
                spl_index(rsq * recip, Q.rdr, Q.nr, m, p);
                const double dfji = (gj != 0.0) ? spl_der(ld_sector(rhor + Q.roff + m), p, Q.rdr) : 0.0;
                const double phip = spl_der(ld_sector(z2r + Q.zoff + min(m, Q.nz)), p, Q.rdr);
                coef += -gj * dfji * recip + 0.5 * (-phip * recip);
              }
            }
            fx[c] -= dx * coef;
            fy[c] -= dy * coef;
            fz[c] -= dz * coef;
            if (ATOM) {
              const double hh = 0.5 * coef;
              av[ATOM ? c : 0][0] += dx * dx * hh;
              av[ATOM ? c : 0][1] += dy * dy * hh;
              av[ATOM ? c : 0][2] += dz * dz * hh;
              av[ATOM ? c : 0][3] += dx * dy * hh;
              av[ATOM ? c : 0][4] += dx * dz * hh;
              av[ATOM ? c : 0][5] += dy * dz * hh;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CPL; c++) {
    fx[c] = part_sum<LPE>(fx[c]);
    fy[c] = part_sum<LPE>(fy[c]);
    fz[c] = part_sum<LPE>(fz[c]);
  }
  if (slot == 0) {
#pragma unroll
    for (int c = 0; c < CPL; c++)
      if (ti[c] >= 0) {
        f[3 * (size_t) (i0 + c0 + c)] += fx[c];    // one lane per center; B2 (atomics) runs after this kernel
        f[3 * (size_t) (i0 + c0 + c) + 1] += fy[c];
        f[3 * (size_t) (i0 + c0 + c) + 2] += fz[c];
      }
  }
  if (ATOM) {
#pragma unroll
    for (int c = 0; c < CPL; c++) {
      const double e1 = part_sum<LPE>(ea[ATOM ? c : 0]);
      double a6[6];
#pragma unroll
      for (int k = 0; k < 6; k++) a6[k] = part_sum<LPE>(av[ATOM ? c : 0][k]);
      if (slot == 0 && ti[c] >= 0) {
        pa_e[i0 + c0 + c] += e1;
#pragma unroll
        for (int k = 0; k < 6; k++) pa_v[6 * (size_t) (i0 + c0 + c) + k] += a6[k];
      }
    }
  }
  if (EV) block_accumulate<7, BLOCK>(ev, scal);
}

// ================================================================== A2 / B2: angular atoms (warp per atom)
struct AngStage {
  double dx[ANG_CAP], dy[ANG_CAP], dz[ANG_CAP], ri[ANG_CAP], f[ANG_CAP], df[ANG_CAP];    // ri = 1 / r
  int j[ANG_CAP];
  unsigned char inD[ANG_CAP];    // passes the density-pass cutoff (cut - CutDec)
};

// stage the in-range neighbors of angular center i in row order; returns count (uniform over the warp)
__device__ __forceinline__ int ang_stage(const AeamDev &par, const double4 *__restrict__ xq,
                                         const double4 *__restrict__ rhor, const int *row, int n,
                                         const double4 &xi, int ti, bool force_pass, AngStage &S, int lane,
                                         int *flags, int self)
{
  const unsigned lt = (1u << lane) - 1u;
  int ns = 0;
  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    bool keep = false, inD = false;
    int j = 0;
    double dx = 0, dy = 0, dz = 0, r1 = 0, fv = 0, dfv = 0;
    if (e < n) {
      j = row[e];
      const double4 xj = xq[j];
      dx = xj.x - xi.x;
      dy = xj.y - xi.y;
      dz = xj.z - xi.z;
      r1 = sqrt(dx * dx + dy * dy + dz * dz);
      const int tj = etype(xj);
      const int pt = ti * par.nel + tj;
      const double cutdec = (tj >= par.nnonangular) ? 1.5 : 0.0;    // i is angular here
      inD = !(r1 > par.cut[pt] - cutdec);
      // force pass: j-role uses the plain cutoff (pair_aeam.cpp:350), k-role the reduced one (:408-418)
      keep = force_pass ? !(r1 > par.cut[pt]) : inD;
      if (j == self) keep = inD = false;    // union rows of a cluster hold the cluster's own centers
      if (keep) {
        int m;
        double p;
        spl_index(r1, par.rdr[pt], par.nr[pt], m, p);
        const double4 cf = rhor[par.rhor_off[pt] + m];
        fv = spl_val(cf, p);
        dfv = spl_der(cf, p, par.rdr[pt]);
      }
    }
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = ns + __popc(mk & lt);
      if (pos < ANG_CAP) {
        S.dx[pos] = dx; S.dy[pos] = dy; S.dz[pos] = dz; S.ri[pos] = 1.0 / r1; S.f[pos] = fv; S.df[pos] = dfv;
        S.j[pos] = j;
        S.inD[pos] = inD ? 1 : 0;
      } else
        flags[0] = 3;
    }
    ns += __popc(mk);
  }
  __syncwarp();
  return min(ns, ANG_CAP);
}

__global__ void __launch_bounds__(128) aeam_density_ang_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ ea_off,
    const int *__restrict__ ea_num, const int *__restrict__ ea_val, const double4 *__restrict__ rhor,
    const int *__restrict__ ang_list, const int *__restrict__ n_ang_ptr, double *__restrict__ rho,
    int *__restrict__ flags, int rshift)
{
  __shared__ AngStage stage[4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_ang = *n_ang_ptr;
  for (int a = blockIdx.x * 4 + wid; a < n_ang; a += gridDim.x * 4) {
    const int i = ang_list[a];
    const double4 xi = xq[i];
    const int ti = etype(xi);
    AngStage &S = stage[wid];
    const int ns = ang_stage(par, xq, rhor, ea_val + ea_off[i >> rshift], ea_num[i >> rshift], xi, ti, false, S, lane,
                             flags, i);
    double acc = 0.0;
    for (int p = 0; p < ns; p++) {
      const double ri1 = S.ri[p], fij = S.f[p];
      const double rsq1 = S.dx[p] * S.dx[p] + S.dy[p] * S.dy[p] + S.dz[p] * S.dz[p];
      for (int q = p + 1 + lane; q < ns; q += 32) {
        const double ex = S.dx[q] - S.dx[p], ey = S.dy[q] - S.dy[p], ez = S.dz[q] - S.dz[p];
        const double rsq3 = ex * ex + ey * ey + ez * ez;
        const double rsq2 = S.dx[q] * S.dx[q] + S.dy[q] * S.dy[q] + S.dz[q] * S.dz[q];
        const double cs = 0.5 * (rsq1 + rsq2 - rsq3) * (ri1 * S.ri[q]);
        const double delcs = cs + (1.0 / 3.0);
        acc += 2 * fij * S.f[q] * (delcs * delcs);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) rho[i] = acc;
    __syncwarp();
  }
}

// ================================================================== A3: embedding
// Also gates the atom's own F' into xq.w (what aeam_gate_kernel does for the ghosts once the halo has brought their fp).
__global__ void __launch_bounds__(BLOCK) aeam_embed_kernel(const __grid_constant__ AeamDev par,
                                                           double4 *xq,
                                                           const double4 *__restrict__ frho,
                                                           const double *__restrict__ rho, int inum,
                                                           double *__restrict__ fp, double *__restrict__ scal,
                                                           double *__restrict__ pa_e)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  double e[1] = {0.0};
  if (i < inum) {
    const int ti = etype(xq[i]);
    const double rh = rho[i];
    // p = pow(rho, ni)/drho + 1 with ni = 1 (non-angular) or 1/2 (angular)   pair_aeam.cpp:274-288
    const double arg = (ti < par.nnonangular) ? rh : sqrt(rh);
    double p = arg * par.rdrho[ti] + 1.0;
    int m = (int) p;
    m = max(1, min(m, par.nrho[ti] - 1));
    p -= m;
    p = fmin(p, 1.0);
    const double4 cf = frho[par.frho_off[ti] + m];
    const double fpi = spl_der(cf, p, par.rdrho[ti]);
    fp[i] = fpi;
    xq[i].w = w_encode((ti < par.nnonangular && rh > 0.0000000000001) ? fpi : 0.0, ti);
    e[0] = spl_val(cf, p);
    // pair_aeam.cpp:295-300: an angular atom's own share of its embedding energy is one third
    if (pa_e) pa_e[i] += (ti < par.nnonangular) ? e[0] : e[0] * (1.0 / 3.0);
  }
  block_accumulate<1, BLOCK>(e, scal);
}

// single-rank convenience: ghost fp/rho from the owner with the same atom ID
__global__ void __launch_bounds__(BLOCK) aeam_tagmap_kernel(const int *__restrict__ tag, int nlocal,
                                                            int *__restrict__ owner_of_tag, int maxtag)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i < nlocal) {
    const int t = tag[i];
    if (t >= 0 && t <= maxtag) owner_of_tag[t] = i;
  }
}
__global__ void __launch_bounds__(BLOCK) aeam_ghost_fill_kernel(const int *__restrict__ tag, int nlocal,
                                                                int nall, const int *__restrict__ owner_of_tag,
                                                                int maxtag, double *__restrict__ fp,
                                                                double *__restrict__ rho, int *__restrict__ flags)
{
  const int g = nlocal + blockIdx.x * BLOCK + threadIdx.x;
  if (g >= nall) return;
  const int t = tag[g];
  const int o = (t >= 0 && t <= maxtag) ? owner_of_tag[t] : -1;
  if (o < 0) {
    flags[7] = 1;
    return;
  }
  fp[g] = fp[o];
  rho[g] = rho[o];
}

// (1-del) * Fptmp * fp of every owned and ghost atom in one array: zero for angular atoms and for
// rho <= minrho (pair_aeam.cpp:128,329-332), so the force kernel gathers ONE double per neighbor
// (atoms [first, nall): the ghosts -- owned atoms are gated by aeam_embed_kernel)
__global__ void __launch_bounds__(BLOCK) aeam_gate_kernel(double4 *__restrict__ xq, const double *__restrict__ rho,
                                                          const double *__restrict__ fp, int nna, int first, int nall)
{
  const int a = first + blockIdx.x * BLOCK + threadIdx.x;
  if (a >= nall) return;
  const int t = (int) (__double_as_longlong(xq[a].w) & 3LL);
  xq[a].w = w_encode((t < nna && rho[a] > 0.0000000000001) ? fp[a] : 0.0, t);
}

// ================================================================== B1: pair + embedding forces (gather)
template <bool EV, bool ATOM, bool ZK>
__global__ void __launch_bounds__(BLOCK) aeam_force_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ ea_off,
    const int *__restrict__ ea_num, const int *__restrict__ ea_val, const double4 *__restrict__ ptab,
    int inum, double *__restrict__ f, double *__restrict__ scal, double *__restrict__ pa_e,
    double *__restrict__ pa_v)
{
  __shared__ PairPar sp[16];
  load_pair_par(par, sp);
  const int tid = blockIdx.x * BLOCK + threadIdx.x;
  const int i = tid >> 3, sub = tid & 7;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  double av[6] = {0, 0, 0, 0, 0, 0};    // ATOM: ev_tally halves of visit (i,j) AND of visit (j,i) both land on i
  if (i < inum) {
    const double4 xi = xq[i];
    const int ti = etype(xi);
    const int nel = par.nel;
    const double gi = w_gate(xi);
    const int n = ea_num[i];
    const int *row = ea_val + ea_off[i];
    for (int e0 = 0; e0 < n; e0 += 32) {
      int jj[4];
      double4 xj[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * 8 + sub;
        jj[u] = (e < n) ? ld_stream_int(row + e) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (jj[u] >= 0) xj[u] = ld_sector(xq + jj[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (jj[u] < 0) continue;
        const double dx = xj[u].x - xi.x, dy = xj[u].y - xi.y, dz = xj[u].z - xi.z;
        const double rsq = dx * dx + dy * dy + dz * dz;
        const int tj = etype(xj[u]);
        const PairPar pij = sp[ti * nel + tj];
        const bool same = (tj == ti);
        const bool in_ij = rsq < pij.cutgt;    // !(r > cut[ti][tj])
        if (same && !in_ij) continue;
        const double r1 = sqrt(rsq);
        const double recip = 1.0 / r1;
        const double gj = w_gate(xj[u]);
        double coef = 0.0;    // fpair of visit (i,j) + fpair of visit (j,i)
        if (in_ij) {
          // visit (i,j): pair_aeam.cpp:350-393
          int m;
          double p;
          spl_index(r1, pij.rdr, pij.nr, m, p);
          const double4 cr = ld_sector(ptab + 2 * (size_t) (pij.off + m));
          const double4 cz = ld_sector(ptab + 2 * (size_t) (pij.off + m) + 1);
          const double dfij = spl_der(cr, p, pij.rdr);
          double phip = spl_der(cz, p, pij.rdr);
          if (ZK && !same) phip *= z_scale(par, ti * nel + tj);
          const double fpair = -gi * dfij * recip + 0.5 * (-phip * recip);
          coef = fpair;
          // same element: visit (j,i) reads the same rows with the same (m, p) -- evaluated without a
          // second gather (98.5 % of the pairs of the AlSi workload)
          if (same) coef += -gj * dfij * recip + 0.5 * (-phip * recip);
          if (EV) {
            ev[0] += 0.5 * spl_val(cz, p);
            ev[1] += dx * dx * fpair;
            ev[2] += dy * dy * fpair;
            ev[3] += dz * dz * fpair;
            ev[4] += dx * dy * fpair;
            ev[5] += dx * dz * fpair;
            ev[6] += dy * dz * fpair;
          }
        }
        if (!same) {
          const PairPar pji = sp[tj * nel + ti];
          if (rsq < pji.cutgt) {
            // visit (j,i), evaluated here instead of scattering from j's row
            int m;
            double p;
            spl_index(r1, pji.rdr, pji.nr, m, p);
            const double dfji = spl_der(ld_sector(ptab + 2 * (size_t) (pji.off + m)), p, pji.rdr);
            double phip = spl_der(ld_sector(ptab + 2 * (size_t) (pji.off + m) + 1), p, pji.rdr);
            if (ZK) phip *= z_scale(par, tj * nel + ti);
            coef += -gj * dfji * recip + 0.5 * (-phip * recip);
          }
        }
        fx -= dx * coef;
        fy -= dy * coef;
        fz -= dz * coef;
        if (ATOM) {
          const double h = 0.5 * coef;
          av[0] += dx * dx * h;
          av[1] += dy * dy * h;
          av[2] += dz * dz * h;
          av[3] += dx * dy * h;
          av[4] += dx * dz * h;
          av[5] += dy * dz * h;
        }
      }
    }
  }
  fx = group_sum<8>(fx);
  fy = group_sum<8>(fy);
  fz = group_sum<8>(fz);
  if (i < inum && sub == 0) {
    f[3 * (size_t) i] += fx;    // one group per atom; B2 (atomics) runs after this kernel
    f[3 * (size_t) i + 1] += fy;
    f[3 * (size_t) i + 2] += fz;
  }
  if (ATOM) {
    // pair_aeam.cpp:389-393: eatom[i] += phi/2 per visit (i,j); ev_tally(i,j,..,0,0,fpair,del): del (x) del fpair, half
    // to each end -- as i of its own visits and as j of its neighbors' visits this atom collects (fpair_ij + fpair_ji)/2
    const double ea = group_sum<8>(ev[0]);
#pragma unroll
    for (int k = 0; k < 6; k++) av[k] = group_sum<8>(av[k]);
    if (i < inum && sub == 0) {
      pa_e[i] += ea;
#pragma unroll
      for (int k = 0; k < 6; k++) pa_v[6 * (size_t) i + k] += av[k];
    }
  }
  if (EV) block_accumulate<7, BLOCK>(ev, scal);
}


// B1 with f' handed over by the density pass (option aeam_cluster = 2): per in-range pair ONE spline gather (the phi
// row of the separate z2r table) instead of the fused 64-byte {rho' | phi} row, i.e. 173 instead of 223 sectors per
// atom (ncu r02).  The kernel is latency-sensitive (every trip is row indices -> positions -> spline rows), so the
// next trip's indices are fetched a trip ahead, the phi rows of two candidates are in flight together, and bond length
// and reciprocal come from one rsqrt; 3 CTAs per SM.
template <bool EV, bool ATOM, int U, int MINB, bool RANGED = false>
__global__ void __launch_bounds__(BLOCK, MINB) aeam_force_df_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ ea_off,
    const int *__restrict__ ea_num, const int *__restrict__ ea_val, const double *__restrict__ ea_df,
    const double4 *__restrict__ rhor, const double4 *__restrict__ z2r, int first, int inum, double *__restrict__ f,
    double *__restrict__ scal, double *__restrict__ pa_e, double *__restrict__ pa_v)
{
  // centers [first, inum): all owned atoms, or -- plugin mode, RANGED -- one range of them (the forces of a range are
  // complete when its launch is, and travel to the host behind it while the next range computes).  An instance of its
  // own: the one more live value costs the resident loop's instance 4 bytes of spill and 1 % (hardware fact 7).
  __shared__ PairPar sp[16];
  load_pair_par(par, sp);
  const int tid = blockIdx.x * BLOCK + threadIdx.x;
  const int i = (RANGED ? first : 0) + (tid >> 3), sub = tid & 7;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  double av[6] = {0, 0, 0, 0, 0, 0};
  if (i < inum) {
    const double4 xi = xq[i];
    const int ti = etype(xi);
    const int nel = par.nel;
    const double gi = w_gate(xi);
    // an angular center stored nothing in the density pass (its density comes from the triplet kernel); its gate is 0
    const bool has_df = ti < par.nnonangular;
    const int n = ea_num[i];
    const int64_t off = ea_off[i];
    const int *row = ea_val + off;
    const double *dfrow = ea_df + off;
    const PairPar *spi = sp + ti * nel;
    int jn[U];
#pragma unroll
    for (int u = 0; u < U; u++) jn[u] = (u * 8 + sub < n) ? ld_stream_int(row + u * 8 + sub) : -1;
    for (int e0 = 0; e0 < n; e0 += 8 * U) {
      int jj[U];
      double4 xj[U];
      double dfs[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        jj[u] = jn[u];
        const int e = e0 + 8 * U + u * 8 + sub;
        jn[u] = (e < n) ? ld_stream_int(row + e) : -1;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        dfs[u] = 0.0;
        if (jj[u] >= 0) {
          xj[u] = ld_sector(xq + jj[u]);
          if (has_df) DfVec<1>::ld(dfrow + e0 + u * 8 + sub, &dfs[u]);
        }
      }
#pragma unroll
      for (int h = 0; h < U; h += 2) {
        bool in_ij[2];
        double rinv[2], pp[2];
        int za[2];
        double4 zw[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const int u = h + k;
          in_ij[k] = false;
          rinv[k] = pp[k] = 0.0;
          za[k] = 0;
          if (jj[u] < 0) continue;
          const double dx = xj[u].x - xi.x, dy = xj[u].y - xi.y, dz = xj[u].z - xi.z;
          const double rsq = dx * dx + dy * dy + dz * dz;
          const PairPar &P = spi[etype(xj[u])];
          in_ij[k] = rsq < P.cutgt;    // !(r > cut[ti][tj])
          rinv[k] = rsqrt_nr(rsq);
          int m;
          spl_index(rsq * rinv[k], P.rdr, P.nr, m, pp[k]);
          za[k] = P.zoff + min(m, P.nz);
        }
#pragma unroll
        for (int k = 0; k < 2; k++)
          if (in_ij[k]) zw[k] = ld_sector(z2r + za[k]);
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const int u = h + k;
          if (jj[u] < 0) continue;
          const int tj = etype(xj[u]);
          const bool same = (tj == ti);
          if (same && !in_ij[k]) continue;
          const double dx = xj[u].x - xi.x, dy = xj[u].y - xi.y, dz = xj[u].z - xi.z;
          const double recip = rinv[k];
          const double gj = w_gate(xj[u]);
          double coef = 0.0;    // fpair of visit (i,j) + fpair of visit (j,i)
          if (in_ij[k]) {
            // visit (i,j): pair_aeam.cpp:350-393
            const double dfij = dfs[u];
            const double phip = spl_der(zw[k], pp[k], spi[tj].rdr);    // symmetric dr only (asym_dr -> round-1 kernel)
            const double fpair = -gi * dfij * recip + 0.5 * (-phip * recip);
            coef = fpair;
            // same element: visit (j,i) evaluates the same two splines at the same (m, p)
            if (same) coef += -gj * dfij * recip + 0.5 * (-phip * recip);
            if (EV) {
              ev[0] += 0.5 * spl_val(zw[k], pp[k]);
              ev[1] += dx * dx * fpair;
              ev[2] += dy * dy * fpair;
              ev[3] += dz * dz * fpair;
              ev[4] += dx * dy * fpair;
              ev[5] += dx * dz * fpair;
              ev[6] += dy * dz * fpair;
            }
          }
          if (!same) {
            const PairPar &Q = sp[tj * nel + ti];
            const double rsq = dx * dx + dy * dy + dz * dz;
            if (rsq < Q.cutgt) {
              // visit (j,i), evaluated here instead of scattering from j's row (different elements: own tables)
              int m;
              double p;
              spl_index(rsq * recip, Q.rdr, Q.nr, m, p);
              const double dfji = (gj != 0.0) ? spl_der(ld_sector(rhor + Q.roff + m), p, Q.rdr) : 0.0;
              const double phip = spl_der(ld_sector(z2r + Q.zoff + min(m, Q.nz)), p, Q.rdr);
              coef += -gj * dfji * recip + 0.5 * (-phip * recip);
            }
          }
          fx -= dx * coef;
          fy -= dy * coef;
          fz -= dz * coef;
          if (ATOM) {
            const double hh = 0.5 * coef;
            av[0] += dx * dx * hh;
            av[1] += dy * dy * hh;
            av[2] += dz * dz * hh;
            av[3] += dx * dy * hh;
            av[4] += dx * dz * hh;
            av[5] += dy * dz * hh;
          }
        }
      }
    }
  }
  fx = group_sum<8>(fx);
  fy = group_sum<8>(fy);
  fz = group_sum<8>(fz);
  if (i < inum && sub == 0) {
    // one add per atom from this kernel; atomic because the angular kernel and -- with halo overlap -- the reverse-halo
    // folds on another stream add to the same atoms
    atomicAdd(&f[3 * (size_t) i], fx);
    atomicAdd(&f[3 * (size_t) i + 1], fy);
    atomicAdd(&f[3 * (size_t) i + 2], fz);
  }
  if (ATOM) {
    const double ea = group_sum<8>(ev[0]);
#pragma unroll
    for (int k = 0; k < 6; k++) av[k] = group_sum<8>(av[k]);
    if (i < inum && sub == 0) {
      pa_e[i] += ea;
#pragma unroll
      for (int k = 0; k < 6; k++) pa_v[6 * (size_t) i + k] += av[k];
    }
  }
  if (EV) block_accumulate<7, BLOCK>(ev, scal);
}

// ================================================================== B2: 3-body forces of angular atoms
// force on atom j from an angular center: FP64 atomics, or -- deterministic mode -- fixed point (2^-44 eV/A resolution,
// range +-5e5 eV/A) into an int64 shadow array, where the order of the adds does not matter
#define AEAM_FIX_SCALE 17592186044416.0
__device__ __forceinline__ void force_add(double *f, unsigned long long *ffix, int j, double ax, double ay, double az)
{
  if (ffix) {
    atomicAdd(&ffix[3 * (size_t) j], (unsigned long long) __double2ll_rn(ax * AEAM_FIX_SCALE));
    atomicAdd(&ffix[3 * (size_t) j + 1], (unsigned long long) __double2ll_rn(ay * AEAM_FIX_SCALE));
    atomicAdd(&ffix[3 * (size_t) j + 2], (unsigned long long) __double2ll_rn(az * AEAM_FIX_SCALE));
  } else {
    atomicAdd(&f[3 * (size_t) j], ax);
    atomicAdd(&f[3 * (size_t) j + 1], ay);
    atomicAdd(&f[3 * (size_t) j + 2], az);
  }
}
__global__ void __launch_bounds__(BLOCK) aeam_fold_fixed_kernel(const long long *__restrict__ ffix, size_t n3,
                                                                double *__restrict__ f)
{
  const size_t k = (size_t) blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n3) return;
  const long long v = ffix[k];
  if (v) f[k] += (double) v * (1.0 / AEAM_FIX_SCALE);
}

template <bool EV, bool ATOM>
__global__ void __launch_bounds__(128) aeam_force_ang_kernel(
    const __grid_constant__ AeamDev par, const double4 *__restrict__ xq, const int64_t *__restrict__ ea_off,
    const int *__restrict__ ea_num, const int *__restrict__ ea_val, const double4 *__restrict__ rhor,
    const int *__restrict__ ang_list, const int *__restrict__ n_ang_ptr, const double *__restrict__ rho,
    const double *__restrict__ fp, double *__restrict__ f, double *__restrict__ scal, int *__restrict__ flags,
    double *__restrict__ pa_v, int rshift, unsigned long long *__restrict__ ffix)
{
  __shared__ AngStage stage[4];
  __shared__ double fstage[4][3][ANG_CAP];
  const double minrho = 0.0000000000001;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_ang = *n_ang_ptr;
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int a = blockIdx.x * 4 + wid; a < n_ang; a += gridDim.x * 4) {
    const int i = ang_list[a];
    const double4 xi = xq[i];
    const int ti = etype(xi);
    AngStage &S = stage[wid];
    const int ns = ang_stage(par, xq, rhor, ea_val + ea_off[i >> rshift], ea_num[i >> rshift], xi, ti, true, S, lane,
                             flags, i);
    // Fptmp*fp[i], Fptmp = ni*rho^(ni-1) = 0.5/sqrt(rho)   (pair_aeam.cpp:329-332)
    const double rh = rho[i];
    const double G = (rh > minrho) ? 0.5 / sqrt(rh) * fp[i] : 0.0;
    const double ci = 2.0;
    double fix = 0.0, fiy = 0.0, fiz = 0.0;
    // forces on the staged neighbors are collected in shared memory and leave the warp ONCE per neighbor (the
    // reference scatters f[j], f[k] per triplet, pair_aeam.cpp:462-470: ~2400 triplets for a center with 70 neighbors):
    // 3 x 70 instead of 3 x 2500 FP64 atomics per center, and a fixed summation order inside the center
    for (int q = lane; q < ns; q += 32) fstage[wid][0][q] = fstage[wid][1][q] = fstage[wid][2][q] = 0.0;
    __syncwarp();
    for (int p = 0; p < ns; p++) {
      // every 1/r comes from the stage (one division per staged neighbor) or, for r_jk, from one rsqrt: the triplet loop
      // had 8 divisions and a sqrt per (j, k) -- most of its FP64 instructions
      const double ri1 = S.ri[p], fij = S.f[p], dfij = S.df[p];
      const double d1x = S.dx[p], d1y = S.dy[p], d1z = S.dz[p];
      const double rsq1 = d1x * d1x + d1y * d1y + d1z * d1z;
      double fjx = 0.0, fjy = 0.0, fjz = 0.0;
      for (int q = p + 1 + lane; q < ns; q += 32) {
        if (!S.inD[q]) continue;
        const double ri2 = S.ri[q], fik = S.f[q], dfik = S.df[q];
        const double d2x = S.dx[q], d2y = S.dy[q], d2z = S.dz[q];
        const double d3x = d2x - d1x, d3y = d2y - d1y, d3z = d2z - d1z;
        const double rsq2 = d2x * d2x + d2y * d2y + d2z * d2z;
        const double rsq3 = d3x * d3x + d3y * d3y + d3z * d3z;
        const double ri3 = rsqrt_nr(rsq3);
        const double r3 = rsq3 * ri3;
        const double i12 = ri1 * ri2;
        const double cs = 0.5 * (rsq1 + rsq2 - rsq3) * i12;
        const double dcosij = ri2 - cs * ri1;
        const double dcosik = ri1 - cs * ri2;
        const double dcosjk = -r3 * i12;
        const double delcs = cs + (1.0 / 3.0);
        const double ftet = delcs * delcs;
        const double delcs2 = 2 * delcs;
        const double DFij = ci * (fik * dfij * ftet + fij * fik * delcs2 * dcosij);
        const double DFik = ci * (fij * dfik * ftet + fij * fik * delcs2 * dcosik);
        const double DFjk = ci * fij * fik * delcs2 * dcosjk;
        const double FFij = -G * DFij * ri1;
        const double FFik = -G * DFik * ri2;
        const double FFjk = -G * DFjk * ri3;
        const double fj0 = d1x * FFij - d3x * FFjk, fj1 = d1y * FFij - d3y * FFjk, fj2 = d1z * FFij - d3z * FFjk;
        const double fk0 = d2x * FFik + d3x * FFjk, fk1 = d2y * FFik + d3y * FFjk, fk2 = d2z * FFik + d3z * FFjk;
        fjx += fj0; fjy += fj1; fjz += fj2;
        fix -= fj0 + fk0; fiy -= fj1 + fk1; fiz -= fj2 + fk2;
        // a lane owns slot q within this p (distinct q per lane), and the p iterations are ordered by the __syncwarp below
        fstage[wid][0][q] += fk0;
        fstage[wid][1][q] += fk1;
        fstage[wid][2][q] += fk2;
        if (ATOM) {    // ev_tally3: thirds to i, j, k
          const int k = S.j[q];
          const double t[6] = {(d1x * fj0 + d2x * fk0) * (1.0 / 3.0), (d1y * fj1 + d2y * fk1) * (1.0 / 3.0),
                               (d1z * fj2 + d2z * fk2) * (1.0 / 3.0), (d1x * fj1 + d2x * fk1) * (1.0 / 3.0),
                               (d1x * fj2 + d2x * fk2) * (1.0 / 3.0), (d1y * fj2 + d2y * fk2) * (1.0 / 3.0)};
          double *vi = pa_v + 6 * (size_t) i, *vj = pa_v + 6 * (size_t) S.j[p], *vk = pa_v + 6 * (size_t) k;
#pragma unroll
          for (int m = 0; m < 6; m++) {
            atomicAdd(vi + m, t[m]);
            atomicAdd(vj + m, t[m]);
            atomicAdd(vk + m, t[m]);
          }
        }
        if (EV) {    // ev_tally3 (delr1, delr2 are x_j - x_i, x_k - x_i)
          v[0] += d1x * fj0 + d2x * fk0;
          v[1] += d1y * fj1 + d2y * fk1;
          v[2] += d1z * fj2 + d2z * fk2;
          v[3] += d1x * fj1 + d2x * fk1;
          v[4] += d1x * fj2 + d2x * fk2;
          v[5] += d1y * fj2 + d2y * fk2;
        }
      }
      fjx = warp_sum(fjx);
      fjy = warp_sum(fjy);
      fjz = warp_sum(fjz);
      __syncwarp();
      if (lane == 0) {
        fstage[wid][0][p] += fjx;
        fstage[wid][1][p] += fjy;
        fstage[wid][2][p] += fjz;
      }
      __syncwarp();
    }
    fix = warp_sum(fix);
    fiy = warp_sum(fiy);
    fiz = warp_sum(fiz);
    // one update per staged neighbor and one for the center; in deterministic mode in fixed point (integer atomics are
    // associative), folded into f by aeam_fold_fixed_kernel afterwards
    for (int q = lane; q < ns; q += 32) {
      const double ax = fstage[wid][0][q], ay = fstage[wid][1][q], az = fstage[wid][2][q];
      if (ax != 0.0 || ay != 0.0 || az != 0.0) force_add(f, ffix, S.j[q], ax, ay, az);
    }
    if (lane == 0) force_add(f, ffix, i, fix, fiy, fiz);
    __syncwarp();
  }
  if (EV) {
    __shared__ double sh[6][4];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double s = warp_sum(v[k]);
      if (lane == 0) sh[k][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < 6)
      accumulate_global(scal, 1 + threadIdx.x, sh[threadIdx.x][0] + sh[threadIdx.x][1] + sh[threadIdx.x][2] + sh[threadIdx.x][3]);
  }
}

// ================================================================== host side
static inline int nblocks(long long n, int per) { return (int) ((n + per - 1) / per); }
// row form actually used: a potential file with asymmetric dr takes the round-1 kernels (the only ones with the
// dr[i][j] / dr[max][min] factor on the phi' of unlike pairs)
static inline int aeam_row_mode(const b200md_ctx *c) { return c->ap.asym_dr ? 0 : c->aeam_cluster; }

int b200md_aeam_pack(b200md_ctx *c)
{
  if (c->nall == 0) return B200MD_OK;
  LaunchScope ls(c, "pack");
  aeam_pack_kernel<<<nblocks(c->nall, BLOCK), BLOCK, 0, c->stream>>>(c->x_aos.p, c->type.p, c->ap.nel, c->nall,
                                                                     c->xq.p, c->flags.p);
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// The list of angular centers in ASCENDING atom order (a scan, not an atomic append): the order decides which warp sums
// which center's virial, so an append made deterministic mode depend on the launch that built the list.
__global__ void __launch_bounds__(BLOCK) aeam_ang_key_kernel(const double4 *__restrict__ xq, int inum, int nna,
                                                             int *__restrict__ key)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i < inum) key[i] = etype(xq[i]) >= nna ? 1 : 0;
}
__global__ void __launch_bounds__(BLOCK) aeam_ang_list_kernel(const int *__restrict__ key, const long long *__restrict__ pos,
                                                              int inum, int *__restrict__ ang_list, int *__restrict__ count)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= inum) return;
  if (key[i]) ang_list[pos[i]] = i;
  if (i == inum - 1) *count = (int) (pos[i] + key[i]);
}
static int aeam_build_ang_list(b200md_ctx *c, int inum)
{
  if (inum <= 0) return B200MD_OK;
  CUDA_TRY(c, c->ang_key.reserve((size_t) inum + 8));
  CUDA_TRY(c, c->ang_scan.reserve((size_t) inum + 2));
  {
    LaunchScope ls(c, "build_inner");
    aeam_ang_key_kernel<<<nblocks(inum, BLOCK), BLOCK, 0, c->stream>>>(c->xq.p, inum, c->ap.nnonangular, c->ang_key.p);
  }
  int rc = b200md_exclusive_scan_i64(c, c->ang_key.p, c->ang_scan.p, inum, 1);
  if (rc) return rc;
  LaunchScope ls(c, "build_inner");
  aeam_ang_list_kernel<<<nblocks(inum, BLOCK), BLOCK, 0, c->stream>>>(c->ang_key.p, (const long long *) c->ang_scan.p, inum,
                                                                     c->ang_list.p, c->flags.p + 6);
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

int b200md_aeam_build_inner(b200md_ctx *c)
{
  const int inum = c->list_inum;
  c->tight_valid = false;    // the third list level exists for rebomos only
  double m = (c->margin_opt > 0.0) ? c->margin_opt : 0.5 * c->skin;    // default: two-level list, inner skin = skin/2
  if (m > c->skin) m = c->skin;
  c->margin = m;
  const int nel = c->ap.nel;
  for (int i = 0; i < nel; i++)
    for (int j = 0; j < nel; j++) {
      const double cc = fmax(c->ap.cut[i * nel + j], c->ap.cut[j * nel + i]) + m;
      c->ap.cutsq_list[i * nel + j] = cc * cc;
    }
  CUDA_TRY(c, c->ang_list.reserve((size_t) inum + 32));
  CUDA_TRY(c, c->xhold.reserve(4 * (size_t) c->nall + 8));
  CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 4, 0, 3 * sizeof(int), c->stream));
  const int mode = aeam_row_mode(c);
  if (mode == 1) {
    const int ncl = (inum + CL - 1) / CL;
    CUDA_TRY(c, c->ec_off.reserve((size_t) ncl + 2));
    CUDA_TRY(c, c->ec_num.reserve((size_t) ncl + 32));
    CUDA_TRY(c, c->ec_cap.reserve((size_t) ncl + 32));
    if (ncl > 0) {
      LaunchScope ls(c, "build_inner");
      aeam_cluster_cap_kernel<<<nblocks(ncl, BLOCK), BLOCK, 0, c->stream>>>(c->list_num.p, inum, ncl, c->ec_cap.p);
    }
    int rc = b200md_exclusive_scan_i64(c, c->ec_cap.p, c->ec_off.p, ncl, 8);
    if (rc) return rc;
    const size_t cap = (size_t) (c->list_entries_used + 8 * (int64_t) ncl + 64);
    CUDA_TRY(c, c->ec_val.reserve(cap));
    CUDA_TRY(c, c->ec_df.reserve(CL * cap));
    if (ncl > 0) {
      LaunchScope ls(c, "build_inner");
      aeam_build_cluster_kernel<<<nblocks((long long) ncl * 32, CL_BLOCK), CL_BLOCK, 0, c->stream>>>(
          c->ap, c->xq.p, c->list_off.p, c->list_num.p, c->list_val.p, inum, ncl, c->ec_off.p, c->ec_num.p, c->ec_val.p,
          c->ang_list.p, c->flags.p, c->aeam_sort_rows);
      CUDA_TRY(c, cudaGetLastError());
    }
  } else {
    CUDA_TRY(c, c->ea_off.reserve((size_t) inum + 2));
    CUDA_TRY(c, c->ea_num.reserve((size_t) inum + 32));
    int rc = b200md_exclusive_scan_i64(c, c->list_num.p, c->ea_off.p, inum, 8);
    if (rc) return rc;
    CUDA_TRY(c, c->ea_val.reserve((size_t) (c->list_entries_used + 8 * (int64_t) inum + 64)));
    if (mode == 2) CUDA_TRY(c, c->ec_df.reserve((size_t) (c->list_entries_used + 8 * (int64_t) inum + 64)));
    if (inum > 0) {
      LaunchScope ls(c, "build_inner");
      aeam_build_inner_kernel<<<nblocks((long long) inum * 32, BLOCK), BLOCK, 0, c->stream>>>(
          c->ap, c->xq.p, c->list_off.p, c->list_num.p, c->list_val.p, inum, c->ea_off.p, c->ea_num.p,
          c->ea_val.p, c->ang_list.p, c->flags.p);
      CUDA_TRY(c, cudaGetLastError());
    }
  }
  {
    int rc = aeam_build_ang_list(c, inum);
    if (rc) return rc;
  }
  CUDA_TRY(c, cudaMemcpyAsync(c->xhold.p, c->xq.p, (size_t) c->nall * sizeof(double4), cudaMemcpyDeviceToDevice,
                              c->stream));
  c->inner_valid = true;
  c->n_inner_rebuild++;
  return B200MD_OK;
}

__global__ void aeam_check_disp_kernel(const double4 *__restrict__ xq, const double4 *__restrict__ xhold,
                                       int nall, double thresh_sq, int *__restrict__ flags)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nall) return;
  double4 a = xq[i], b = xhold[i];
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  if (dx * dx + dy * dy + dz * dz > thresh_sq) flags[1] = 1;
}

int b200md_aeam_refresh_inner(b200md_ctx *c)
{
  if (!c->inner_valid) return b200md_aeam_build_inner(c);
  if (c->margin >= c->skin) return B200MD_OK;
  const double half = 0.5 * c->margin;
  {
    LaunchScope ls(c, "check_disp");
    aeam_check_disp_kernel<<<nblocks(c->nall, BLOCK), BLOCK, 0, c->stream>>>(
        c->xq.p, (const double4 *) c->xhold.p, c->nall, half * half, c->flags.p);
    CUDA_TRY(c, cudaGetLastError());
  }
  int flag = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&flag, c->flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (flag) {
    CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 1, 0, sizeof(int), c->stream));
    return b200md_aeam_build_inner(c);
  }
  return B200MD_OK;
}

static int aeam_density_tail(b200md_ctx *c);

// density + embedding for owned atoms (rho, fp valid for [0,inum) afterwards)
int b200md_aeam_density(b200md_ctx *c)
{
  const int inum = c->list_inum;
  CUDA_TRY(c, c->rho.reserve((size_t) c->nall + 32));
  CUDA_TRY(c, c->fp.reserve((size_t) c->nall + 32));
  if (inum == 0) return B200MD_OK;
  const double4 *rhor = (const double4 *) c->spl_rhor.p;
  const int mode = aeam_row_mode(c);
  const bool cl = mode == 1;
  {
    LaunchScope ls(c, "aeam_density");
    if (cl) {
      const int ncl = (inum + CL - 1) / CL;
#define ADC_ARGS c->ap, c->xq.p, c->ec_off.p, c->ec_num.p, c->ec_val.p, rhor, inum, c->rho.p, c->ec_df.p
#define ADC_LAUNCH(LPE, U, MINB) \
  aeam_density_cl_kernel<LPE, U, MINB><<<nblocks((long long) ncl * 8 * LPE, BLOCK), BLOCK, 0, c->stream>>>(ADC_ARGS)
      switch (c->aeam_variant % 10) {
        case 1: ADC_LAUNCH(1, 2, 2); break;
        case 2: ADC_LAUNCH(2, 1, 3); break;
        case 3: ADC_LAUNCH(2, 2, 3); break;
        case 4: ADC_LAUNCH(4, 1, 4); break;
        case 5: ADC_LAUNCH(4, 2, 4); break;
        case 6: ADC_LAUNCH(4, 4, 4); break;
        case 7: ADC_LAUNCH(2, 2, 2); break;
        default: ADC_LAUNCH(4, 2, 4); break;
      }
    } else if (mode == 2)
      aeam_density_kernel<true><<<nblocks((long long) inum * 8, BLOCK), BLOCK, 0, c->stream>>>(
          c->ap, c->xq.p, c->ea_off.p, c->ea_num.p, c->ea_val.p, rhor, inum, c->rho.p, c->ec_df.p);
    else
      aeam_density_kernel<false><<<nblocks((long long) inum * 8, BLOCK), BLOCK, 0, c->stream>>>(
          c->ap, c->xq.p, c->ea_off.p, c->ea_num.p, c->ea_val.p, rhor, inum, c->rho.p, nullptr);
  }
  return aeam_density_tail(c);
}

// density of the angular atoms + embedding (after the pair densities of all owned atoms)
static int aeam_density_tail(b200md_ctx *c)
{
  const int inum = c->list_inum;
  const double4 *rhor = (const double4 *) c->spl_rhor.p;
  const bool cl = aeam_row_mode(c) == 1;
  const int64_t *r_off = cl ? c->ec_off.p : c->ea_off.p;
  const int *r_num = cl ? c->ec_num.p : c->ea_num.p, *r_val = cl ? c->ec_val.p : c->ea_val.p;
  const int rshift = cl ? CL_SHIFT : 0;
  if (c->ap.nnonangular < c->ap.nel) {
    LaunchScope ls(c, "aeam_density_ang");
    aeam_density_ang_kernel<<<c->num_sms * c->ang_ctas, 128, 0, c->stream>>>(c->ap, c->xq.p, r_off, r_num, r_val, rhor,
                                                                  c->ang_list.p, c->flags.p + 6, c->rho.p, c->flags.p,
                                                                  rshift);
  }
  {
    LaunchScope ls(c, "aeam_embed");
    aeam_embed_kernel<<<nblocks(inum, BLOCK), BLOCK, 0, c->stream>>>(c->ap, c->xq.p, (const double4 *) c->spl_frho.p,
                                                                   c->rho.p, inum, c->fp.p, b200md_scal_arg(c), c->pa_e);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// ghost fp/rho from same-ID owners (valid when every ghost's owner is local: 1 rank, periodic images)
int b200md_aeam_fill_ghosts_by_tag(b200md_ctx *c, int maxtag)
{
  if (c->nghost == 0) return B200MD_OK;
  CUDA_TRY(c, c->scan_tmp.reserve((size_t) maxtag + 2));
  CUDA_TRY(c, cudaMemsetAsync(c->scan_tmp.p, 0xff, ((size_t) maxtag + 1) * sizeof(int), c->stream));
  {
    LaunchScope ls(c, "aeam_tagmap");
    aeam_tagmap_kernel<<<nblocks(c->nlocal, BLOCK), BLOCK, 0, c->stream>>>(c->tag.p, c->nlocal, c->scan_tmp.p, maxtag);
  }
  {
    LaunchScope ls(c, "aeam_ghost_fill");
    aeam_ghost_fill_kernel<<<nblocks(c->nghost, BLOCK), BLOCK, 0, c->stream>>>(
        c->tag.p, c->nlocal, c->nall, c->scan_tmp.p, maxtag, c->fp.p, c->rho.p, c->flags.p);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// forces; needs rho/fp of owned AND ghost atoms
// part 0: everything (gate, pair + embedding forces, angular forces).  Halo overlap (force-only steps, default row form):
// part 1 = gate + angular forces -- the only kernels that write ghost forces -- and part 2 = the pair kernel, which then
// runs beside the reverse halo (every kernel adds to f with atomics, so the order between them is free)
int b200md_aeam_forces(b200md_ctx *c, int eflag, int vflag, int part)
{
  const int inum = c->list_inum;
  if (inum == 0) return B200MD_OK;
  const double4 *rhor = (const double4 *) c->spl_rhor.p;
  const double4 *ptab = (const double4 *) c->spl_pair.p;
  const bool ev = eflag || vflag;
  if (part != 2 && c->nall > inum) {
    LaunchScope ls(c, "aeam_gate");
    aeam_gate_kernel<<<nblocks(c->nall - inum, BLOCK), BLOCK, 0, c->stream>>>(c->xq.p, c->rho.p, c->fp.p,
                                                                            c->ap.nnonangular, inum, c->nall);
  }
  const bool atom = c->pa_e != nullptr;
  const int mode = aeam_row_mode(c);
  const bool cl = mode == 1;
  const int64_t *r_off = cl ? c->ec_off.p : c->ea_off.p;
  const int *r_num = cl ? c->ec_num.p : c->ea_num.p, *r_val = cl ? c->ec_val.p : c->ea_val.p;
  const int rshift = cl ? CL_SHIFT : 0;
  if (part == 1) {
    // pair kernel later (part 2)
  } else if (cl) {
    LaunchScope ls(c, (atom || ev) ? "aeam_force_ev" : "aeam_force");
    const int ncl = (inum + CL - 1) / CL;
#define AFC_ARGS \
  c->ap, c->xq.p, c->ec_off.p, c->ec_num.p, c->ec_val.p, c->ec_df.p, rhor, (const double4 *) c->spl_z2r.p, \
      inum, c->f.p, b200md_scal_arg(c), c->pa_e, c->pa_v
#define AFC_LAUNCH(EV, ATOM, LPE, U, MINB) \
  aeam_force_cl_kernel<EV, ATOM, LPE, U, MINB><<<nblocks((long long) ncl * 8 * LPE, BLOCK), BLOCK, 0, c->stream>>>(AFC_ARGS)
    if (atom) AFC_LAUNCH(true, true, 4, 1, 2);
    else if (ev) AFC_LAUNCH(true, false, 4, 2, 3);
    else
      switch ((c->aeam_variant / 10) % 10) {
        case 1: AFC_LAUNCH(false, false, 1, 1, 2); break;
        case 2: AFC_LAUNCH(false, false, 2, 1, 3); break;
        case 3: AFC_LAUNCH(false, false, 2, 2, 3); break;
        case 4: AFC_LAUNCH(false, false, 4, 1, 4); break;
        case 5: AFC_LAUNCH(false, false, 4, 2, 4); break;
        case 6: AFC_LAUNCH(false, false, 4, 4, 4); break;
        case 7: AFC_LAUNCH(false, false, 2, 2, 2); break;
        default: AFC_LAUNCH(false, false, 4, 2, 4); break;
      }
  } else {
    LaunchScope ls(c, (atom || ev) ? "aeam_force_ev" : "aeam_force");    // thermo steps: the energy/virial instance
    // the default row form takes a range of centers (plugin mode: force download range by range)
    const bool ranged = mode == 2 && c->aeam_range_hi > c->aeam_range_lo;
    const int r_lo = ranged ? c->aeam_range_lo : 0, r_hi = ranged ? c->aeam_range_hi : inum;
    const int nb = nblocks((long long) (mode == 2 ? r_hi - r_lo : inum) * 8, BLOCK);
#define AF_ARGS c->ap, c->xq.p, c->ea_off.p, c->ea_num.p, c->ea_val.p, ptab, inum, c->f.p, b200md_scal_arg(c), c->pa_e, c->pa_v
#define AFD_ARGS \
  c->ap, c->xq.p, c->ea_off.p, c->ea_num.p, c->ea_val.p, c->ec_df.p, rhor, (const double4 *) c->spl_z2r.p, r_lo, r_hi, \
      c->f.p, b200md_scal_arg(c), c->pa_e, c->pa_v
    if (mode == 2) {
      if (atom) aeam_force_df_kernel<true, true, 2, 2><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS);
      else if (ev && ranged) aeam_force_df_kernel<true, false, 4, 2, true><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS);
      else if (ev) aeam_force_df_kernel<true, false, 4, 2><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS);
      else if (ranged) aeam_force_df_kernel<false, false, 2, 3, true><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS);
      else
        switch ((c->aeam_variant / 10) % 10) {
          case 1: aeam_force_df_kernel<false, false, 4, 2><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS); break;
          case 2: aeam_force_df_kernel<false, false, 4, 3><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS); break;
          case 3: aeam_force_df_kernel<false, false, 2, 4><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS); break;
          default: aeam_force_df_kernel<false, false, 2, 3><<<nb, BLOCK, 0, c->stream>>>(AFD_ARGS); break;
        }
    } else {
      if (c->ap.asym_dr) {
        if (atom) aeam_force_kernel<true, true, true><<<nb, BLOCK, 0, c->stream>>>(AF_ARGS);
        else if (ev) aeam_force_kernel<true, false, true><<<nb, BLOCK, 0, c->stream>>>(AF_ARGS);
        else aeam_force_kernel<false, false, true><<<nb, BLOCK, 0, c->stream>>>(AF_ARGS);
      } else {
        if (atom) aeam_force_kernel<true, true, false><<<nb, BLOCK, 0, c->stream>>>(AF_ARGS);
        else if (ev) aeam_force_kernel<true, false, false><<<nb, BLOCK, 0, c->stream>>>(AF_ARGS);
        else aeam_force_kernel<false, false, false><<<nb, BLOCK, 0, c->stream>>>(AF_ARGS);
      }
    }
  }
  if (part != 2 && c->ap.nnonangular < c->ap.nel) {
    LaunchScope ls(c, (atom || ev) ? "aeam_force_ang_ev" : "aeam_force_ang");
    unsigned long long *ffix = nullptr;
    const size_t n3 = 3 * (size_t) c->nall;
    if (c->deterministic) {
      CUDA_TRY(c, c->det_ffix.reserve(n3 + 8));
      CUDA_TRY(c, cudaMemsetAsync(c->det_ffix.p, 0, n3 * sizeof(long long), c->stream));
      ffix = (unsigned long long *) c->det_ffix.p;
    }
#define AA_ARGS \
  c->ap, c->xq.p, r_off, r_num, r_val, rhor, c->ang_list.p, c->flags.p + 6, c->rho.p, c->fp.p, c->f.p, b200md_scal_arg(c), \
      c->flags.p, c->pa_v, rshift, ffix
    if (atom) aeam_force_ang_kernel<true, true><<<c->num_sms * c->ang_ctas, 128, 0, c->stream>>>(AA_ARGS);
    else if (ev) aeam_force_ang_kernel<true, false><<<c->num_sms * c->ang_ctas, 128, 0, c->stream>>>(AA_ARGS);
    else aeam_force_ang_kernel<false, false><<<c->num_sms * c->ang_ctas, 128, 0, c->stream>>>(AA_ARGS);
    if (ffix) aeam_fold_fixed_kernel<<<nblocks((long long) n3, BLOCK), BLOCK, 0, c->stream>>>(c->det_ffix.p, n3, c->f.p);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

static int aeam_check_flags(b200md_ctx *c, const int *fl)
{
  if (fl[5]) c->n_lj_entries = fl[5];
  c->n_ang = fl[6];
  if (fl[3]) {
    c->fail("atom type outside 1..nelements (pair_coeff must map every type, in file order)");
    cudaMemsetAsync(c->flags.p + 3, 0, sizeof(int), c->stream);
    return B200MD_ERR_ARG;
  }
  if (fl[7]) {
    c->fail("aeam_compute: a ghost atom has no owner with the same ID on this rank; use the two-phase "
            "API (b200md_aeam_density / halo exchange of fp / b200md_aeam_force)");
    cudaMemsetAsync(c->flags.p + 7, 0, sizeof(int), c->stream);
    return B200MD_ERR_ARG;
  }
  if (fl[0] == 4) {
    c->fail("AEAM cluster row overflow (more than " + std::to_string(CL_HT * 3 / 4 - 32) +
            " distinct neighbors of 4 consecutive atoms within cut + margin); set option aeam_cluster = 0");
    cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream);
    return B200MD_ERR_OVERFLOW;
  }
  if (fl[0]) {
    c->fail("AEAM angular-neighbor staging overflow (more than " + std::to_string(ANG_CAP) + " neighbors in range)");
    cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream);
    return B200MD_ERR_OVERFLOW;
  }
  return B200MD_OK;
}

static int aeam_begin(b200md_ctx *c, int nlocal, int nghost, const double *x, const int *type, const int *tag)
{
  ARG_CHECK(c, c->aeam_ready, "aeam: call b200md_aeam_init first");
  ARG_CHECK(c, c->list_valid, "aeam: no neighbor list (b200md_set_neighbor_list / b200md_neigh_build)");
  ARG_CHECK(c, c->list_inum == nlocal, "aeam: neighbor list was built for a different nlocal");
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (c->tight_derive_pending) {    // a deferred re-derive of the pipelined path (it reads xq) is still in flight
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->tight_derive_pending = false;
  }
  int rc = b200md_upload_atoms(c, nlocal, nghost, x, type, tag);
  if (rc) return rc;
  const size_t n3 = 3 * (size_t) c->nall;
  CUDA_TRY(c, cudaMemsetAsync(c->f.p, 0, (n3 + 8) * sizeof(double), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), c->stream));
  if ((rc = b200md_aeam_pack(c))) return rc;
  return b200md_aeam_refresh_inner(c);
}

static int aeam_finish(b200md_ctx *c, int eflag, int vflag, double *f, double *eng_vdwl, double *virial)
{
  int fl[16];
  int rc = b200md_finish_compute(c, eflag, vflag, f, eng_vdwl, virial, fl);
  if (rc) return rc;
  return aeam_check_flags(c, fl);
}


// ================================================================== plugin mode: upload pipelined with the density pass
// The positions arrive in h2d_chunks pieces (ghosts first); the density launch of center range k starts as soon as the
// piece holding the largest atom index its rows name has landed.  The dependence is computed from the MASTER rows when a
// list is handed over (valid for every inner list derived from it; the inner rows are re-derived every ~12 calls at 863 K),
// with the straggler list of rebomos.cu's dep_range_kernel: atoms named from more than one piece away travel first.
struct AeamChunks {
  int t[B200MD_MAX_D2H_CHUNKS + 1];
  int K;
};
__global__ void __launch_bounds__(BLOCK) aeam_dep_kernel(const int64_t *__restrict__ list_off,
                                                         const int *__restrict__ list_num,
                                                         const int *__restrict__ list_val, int inum,
                                                         const __grid_constant__ AeamChunks cb, int *__restrict__ dep,
                                                         int *__restrict__ strag_flag, int *__restrict__ strag_list,
                                                         int *__restrict__ strag_count, int strag_cap)
{
  const int tid = blockIdx.x * BLOCK + threadIdx.x;
  const int i = tid >> 3, sub = tid & 7;
  if (i >= inum) return;
  int k = 0;
  while (k + 1 < cb.K && i >= cb.t[k + 1]) k++;
  const int far = strag_flag ? cb.t[min(k + 2, cb.K)] : inum;
  int m = i;
  const int n = list_num[i];
  const int *row = list_val + list_off[i];
  for (int e = sub; e < n; e += 8) {
    const int j = row[e] & B200MD_NEIGHMASK;
    if (j >= inum) continue;    // ghosts travel first
    if (j >= far) {
      if (atomicExch(&strag_flag[j], 1) == 0) {
        const int pos = atomicAdd(strag_count, 1);
        if (pos < strag_cap) strag_list[pos] = j;
      }
    } else if (j > m)
      m = j;
  }
  if (m > dep[k]) atomicMax(&dep[k], m);    // racy read, monotone value: almost every lane skips the atomic
}
__global__ void __launch_bounds__(BLOCK) aeam_strag_scatter_kernel(const double *__restrict__ buf,
                                                                   const int *__restrict__ list, int n,
                                                                   const int *__restrict__ type, int nel,
                                                                   double4 *__restrict__ xq)
{
  const int q = blockIdx.x * BLOCK + threadIdx.x;
  if (q >= n) return;
  const int j = list[q];
  int t = type[j];
  if (t < 1 || t > nel) t = 1;    // flagged by the pack kernel of the atom's piece
  xq[j] = make_double4(buf[3 * (size_t) q], buf[3 * (size_t) q + 1], buf[3 * (size_t) q + 2], w_encode(0.0, t - 1));
}
// flags[1] = 2: an atom moved more than margin/2 since the inner rows were derived (they may miss a pair: recompute);
// 1: more than 80 % of that (the rows are re-derived after this call, before they can miss one)
__global__ void __launch_bounds__(BLOCK) aeam_check_disp2_kernel(const double4 *__restrict__ xq,
                                                                 const double4 *__restrict__ xhold, int nall,
                                                                 double thresh_sq, double soon_sq, int *__restrict__ flags)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nall) return;
  const double4 a = xq[i], b = xhold[i];
  const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  const double d = dx * dx + dy * dy + dz * dz;
  if (d > soon_sq) atomicMax(&flags[1], d > thresh_sq ? 2 : 1);
}

static int aeam_prepare_pipeline(b200md_ctx *c)
{
  const int inum = c->list_inum;
  c->aeam_h2d_ready = false;
  c->aeam_h2d_list = c->n_list_upload;
  if (c->h2d_chunks <= 1 || c->d2h_chunks <= 1 || inum < c->d2h_min_atoms || inum == 0 || aeam_row_mode(c) != 2) return B200MD_OK;
  const int K = c->h2d_chunks;
  AeamChunks cb;
  cb.K = K;
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
  int *dep = c->flags.p + 16;
  int *pin = (int *) (c->pin_scal.p + 48), *pin_cnt = (int *) (c->pin_scal.p + 60);
  const int cap = inum / 32 + 1024;
  c->n_strag = 0;
  CUDA_TRY(c, c->strag_flag.reserve((size_t) inum + 8));
  CUDA_TRY(c, c->strag_list.reserve((size_t) cap + 8));
  for (int pass = 0; pass < 2; pass++) {
    CUDA_TRY(c, cudaMemsetAsync(dep, 0, K * sizeof(int), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->strag_flag.p, 0, ((size_t) inum + 1) * sizeof(int), c->stream));
    int *cnt = c->strag_flag.p + inum;
    {
      LaunchScope ls(c, "build_inner");
      aeam_dep_kernel<<<nblocks((long long) inum * 8, BLOCK), BLOCK, 0, c->stream>>>(
          c->list_off.p, c->list_num.p, c->list_val.p, inum, cb, dep, pass == 0 ? c->strag_flag.p : nullptr, c->strag_list.p,
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
    int p = k;
    while (p + 1 < K && pin[k] >= cb.t[p + 1]) p++;
    c->h2d_need[k] = p;
    c->h2d_t[k] = cb.t[k];
  }
  c->h2d_t[K] = inum;
  c->h2d_K = K;
  c->aeam_h2d_ready = true;
  return B200MD_OK;
}

static int aeam_forces_download_pipelined(b200md_ctx *c, int nlocal, int eflag, int vflag, double *f, double *eng_vdwl,
                                          double *virial, int *fl);

// One plugin-mode call with upload, kernels and download overlapped.  *redo: the inner rows were stale (an atom beyond
// margin/2: rare, the rows are re-derived one call ahead of need) -- nothing has been added to the caller's f, the
// positions are on the device, the caller re-derives and recomputes.
static int aeam_compute_pipelined(b200md_ctx *c, int nlocal, int nghost, const double *x, int eflag, int vflag, double *f,
                                  double *eng_vdwl, double *virial, int *fl, bool *redo)
{
  *redo = false;
  const int K = c->h2d_K;
  const int nall = nlocal + nghost;
  if (!c->up_stream) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
  for (int k = 0; k <= K; k++)
    if (!c->up_ev[k]) CUDA_TRY(c, cudaEventCreateWithFlags(&c->up_ev[k], cudaEventDisableTiming));
  c->nlocal = nlocal;
  c->nghost = nghost;
  c->nall = nall;
  const bool need_check = c->margin < c->skin;
  int *pin_flag = (int *) (c->pin_scal.p + 56);
  pin_flag[0] = 0;
  if (c->tight_derive_pending) {    // a deferred re-derive of the inner rows reads the previous call's positions
    CUDA_TRY(c, cudaStreamWaitEvent(c->up_stream, c->ev_tight, 0));
    c->tight_derive_pending = false;
  }
  // ---- upload stream
  cudaStream_t compute_stream = c->stream;
  c->stream = c->up_stream;
  int rc = B200MD_OK;
  auto piece = [&](int lo, int hi) -> int {
    if (hi <= lo) return B200MD_OK;
    CUDA_TRY(c, cudaMemcpyAsync(c->x_aos.p + 3 * (size_t) lo, x + 3 * (size_t) lo, 3 * (size_t) (hi - lo) * sizeof(double),
                                cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long) (3 * (size_t) (hi - lo) * sizeof(double));
    LaunchScope ls(c, "pack");
    aeam_pack_kernel<<<nblocks(hi - lo, BLOCK), BLOCK, 0, c->stream>>>(c->x_aos.p, c->type.p, c->ap.nel, hi, c->xq.p,
                                                                     c->flags.p, lo);
    return B200MD_OK;
  };
  if (c->n_strag) {
    const int ns = c->n_strag;
    double *sp = c->strag_pin.p;
    for (int q = 0; q < ns; q++) {
      const size_t j = (size_t) c->strag_host[q];
      sp[3 * (size_t) q] = x[3 * j];
      sp[3 * (size_t) q + 1] = x[3 * j + 1];
      sp[3 * (size_t) q + 2] = x[3 * j + 2];
    }
    if (cudaMemcpyAsync(c->strag_dev.p, sp, 3 * (size_t) ns * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
      rc = B200MD_ERR_CUDA;
    c->h2d_bytes += (long long) (3 * (size_t) ns * sizeof(double));
    LaunchScope ls(c, "pack");
    aeam_strag_scatter_kernel<<<nblocks(ns, BLOCK), BLOCK, 0, c->stream>>>(c->strag_dev.p, c->strag_list.p, ns, c->type.p,
                                                                          c->ap.nel, c->xq.p);
  }
  if (!rc) rc = piece(nlocal, nall);
  for (int p = 0; p < K && !rc; p++) {
    rc = piece(c->h2d_t[p], c->h2d_t[p + 1]);
    if (!rc && cudaEventRecord(c->up_ev[p], c->stream) != cudaSuccess) rc = B200MD_ERR_CUDA;
  }
  if (!rc && need_check) {
    const double half = 0.5 * c->margin, soon = 0.8 * half;
    if (cudaMemsetAsync(c->flags.p + 1, 0, sizeof(int), c->stream) != cudaSuccess) rc = B200MD_ERR_CUDA;
    {
      LaunchScope ls(c, "check_disp");
      aeam_check_disp2_kernel<<<nblocks(nall, BLOCK), BLOCK, 0, c->stream>>>(c->xq.p, (const double4 *) c->xhold.p, nall,
                                                                            half * half, soon * soon, c->flags.p);
    }
    if (cudaMemcpyAsync(pin_flag, c->flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      rc = B200MD_ERR_CUDA;
  }
  if (!rc && cudaEventRecord(c->up_ev[K], c->stream) != cudaSuccess) rc = B200MD_ERR_CUDA;
  c->stream = compute_stream;
  if (rc) return rc;
  // ---- compute stream
  const size_t n3 = 3 * (size_t) nall;
  CUDA_TRY(c, cudaMemsetAsync(c->f.p, 0, (n3 + 8) * sizeof(double), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), c->stream));
  CUDA_TRY(c, c->rho.reserve((size_t) nall + 32));
  CUDA_TRY(c, c->fp.reserve((size_t) nall + 32));
  const double4 *rhor = (const double4 *) c->spl_rhor.p;
  for (int k = 0; k < K; k++) {
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->up_ev[c->h2d_need[k]], 0));
    const int lo = c->h2d_t[k], hi = c->h2d_t[k + 1];
    if (hi <= lo) continue;
    LaunchScope ls(c, "aeam_density");
    aeam_density_kernel<true, true><<<nblocks((long long) (hi - lo) * 8, BLOCK), BLOCK, 0, c->stream>>>(
        c->ap, c->xq.p, c->ea_off.p, c->ea_num.p, c->ea_val.p, rhor, hi, c->rho.p, c->ec_df.p, lo);
  }
  CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->up_ev[K - 1], 0));    // the angular centers read anywhere
  if ((rc = aeam_density_tail(c))) return rc;
  if (nghost && (rc = b200md_aeam_fill_ghosts_by_tag(c, c->aeam_maxtag))) return rc;
  // ---- the verdict on the rows arrives while the density kernels are still running
  CUDA_TRY(c, cudaEventSynchronize(c->up_ev[K]));
  if (need_check && pin_flag[0] == 2) {
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    b200md_collect_timers(c);
    *redo = true;
    return B200MD_OK;
  }
  if ((rc = aeam_forces_download_pipelined(c, nlocal, eflag, vflag, f, eng_vdwl, virial, fl))) return rc;
  if (need_check && pin_flag[0] == 1) {
    // an atom is 80 % of the way to the rows' limit: re-derive them now, after this call's work, from the positions on
    // the device -- the host integrates in the meantime, the next call's upload waits for the event
    if ((rc = b200md_aeam_build_inner(c))) return rc;
    if (!c->ev_tight) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_tight, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventRecord(c->ev_tight, c->stream));
    c->tight_derive_pending = true;
  }
  return B200MD_OK;
}

// force download pipelined with the pair kernel (force-only or EV calls, default row form): the angular kernel -- the only
// one that writes ghost forces and forces of other centers -- runs first, the ghost forces leave, then the pair kernel
// runs over d2h_chunks ranges of centers and every range's forces leave behind its launch
static int aeam_forces_download_pipelined(b200md_ctx *c, int nlocal, int eflag, int vflag, double *f, double *eng_vdwl,
                                          double *virial, int *fl)
{
  const int KD = c->d2h_chunks;
  const size_t n3 = 3 * (size_t) c->nall;
  int rc;
  if ((rc = b200md_aeam_forces(c, eflag, vflag, 1))) return rc;
  if ((rc = b200md_d2h_begin(c))) return rc;
  if ((rc = b200md_d2h_range(c, 0, f, 3 * (size_t) nlocal, n3))) return rc;
  for (int k = 0; k < KD && !rc; k++) {
    c->aeam_range_lo = (int) ((long long) nlocal * k / KD);
    c->aeam_range_hi = (int) ((long long) nlocal * (k + 1) / KD);
    if (c->aeam_range_hi <= c->aeam_range_lo) continue;
    if (!(rc = b200md_aeam_forces(c, eflag, vflag, 2)))
      rc = b200md_d2h_range(c, k + 1, f, 3 * (size_t) c->aeam_range_lo, 3 * (size_t) c->aeam_range_hi);
  }
  c->aeam_range_lo = c->aeam_range_hi = 0;
  if (rc) return rc;
  return b200md_d2h_finish(c, eflag, vflag, f, eng_vdwl, virial, fl);
}

// one-shot: ghosts are periodic images of owned atoms (single rank); ghost fp via atom IDs
extern "C" int b200md_aeam_compute(b200md_ctx *c, int nlocal, int nghost, const double *x, const int *type,
                                   const int *tag, int eflag, int vflag, double *f, double *eng_vdwl,
                                   double *virial)
{
  return b200md_aeam_compute_peratom(c, nlocal, nghost, x, type, tag, eflag, vflag, f, eng_vdwl, virial, nullptr, nullptr);
}

extern "C" int b200md_aeam_compute_peratom(b200md_ctx *c, int nlocal, int nghost, const double *x, const int *type,
                                           const int *tag, int eflag, int vflag, double *f, double *eng_vdwl,
                                           double *virial, double *eatom, double *vatom)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, f != nullptr || nlocal + nghost == 0, "aeam_compute: f is NULL");
  ARG_CHECK(c, tag != nullptr || nghost == 0, "aeam_compute: atom IDs are needed to give ghosts their fp");
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->n_compute++;
  if (nlocal + nghost == 0) {    // an empty rank (vacuum brick)
    if (eng_vdwl) *eng_vdwl = 0.0;
    if (virial)
      for (int k = 0; k < 6; k++) virial[k] = 0.0;
    return B200MD_OK;
  }
  int rc;
  const bool pipe_ok = !eatom && !vatom && !c->deterministic && c->d2h_chunks > 1 && nlocal >= c->d2h_min_atoms &&
      aeam_row_mode(c) == 2;
  if (pipe_ok && c->list_valid && c->aeam_h2d_list != c->n_list_upload) c->aeam_h2d_ready = false;    // a new master list
  if (pipe_ok && c->inner_valid && c->aeam_h2d_ready && c->type_on_device && c->tag_on_device &&
      nlocal + nghost == c->ids_nall && c->nlocal == nlocal && c->list_inum == nlocal) {
    bool redo = false;
    int fl[16];
    if ((rc = aeam_compute_pipelined(c, nlocal, nghost, x, eflag, vflag, f, eng_vdwl, virial, fl, &redo))) return rc;
    c->n_pipelined++;
    if (!redo) return aeam_check_flags(c, fl);
    c->n_redo++;
    // stale rows: the positions are on the device; re-derive and recompute by the plain kernels
    if ((rc = b200md_aeam_build_inner(c))) return rc;
    const size_t n3 = 3 * (size_t) c->nall;
    CUDA_TRY(c, cudaMemsetAsync(c->f.p, 0, (n3 + 8) * sizeof(double), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), c->stream));
    if ((rc = b200md_aeam_density(c))) return rc;
    if (nghost && (rc = b200md_aeam_fill_ghosts_by_tag(c, c->aeam_maxtag))) return rc;
    if ((rc = aeam_forces_download_pipelined(c, nlocal, eflag, vflag, f, eng_vdwl, virial, fl))) return rc;
    return aeam_check_flags(c, fl);
  }
  rc = aeam_begin(c, nlocal, nghost, x, type, tag);
  if (rc) return rc;
  if (pipe_ok && !c->aeam_h2d_ready && (rc = aeam_prepare_pipeline(c))) return rc;
  if ((rc = b200md_peratom_begin(c, eatom != nullptr || vatom != nullptr))) return rc;
  rc = b200md_aeam_density(c);
  if (!rc && nghost) {
    int maxtag = 0;
    for (int i = 0; i < nlocal; i++) maxtag = tag[i] > maxtag ? tag[i] : maxtag;
    c->aeam_maxtag = maxtag;
    rc = b200md_aeam_fill_ghosts_by_tag(c, maxtag);
  }
  if (!rc && pipe_ok) {
    int fl[16];
    if ((rc = aeam_forces_download_pipelined(c, nlocal, eflag, vflag, f, eng_vdwl, virial, fl))) return rc;
    c->n_pipelined++;
    return aeam_check_flags(c, fl);
  }
  if (!rc) rc = b200md_aeam_forces(c, eflag, vflag, 0);
  if (rc) {
    c->pa_e = c->pa_v = nullptr;
    return rc;
  }
  if ((rc = b200md_peratom_finish(c, eatom, vatom))) return rc;
  return aeam_finish(c, eflag, vflag, f, eng_vdwl, virial);
}

// option "fp_gated": the density phase hands out (rho > minrho ? fp : 0) -- what a neighbor needs from an atom
// (pair_aeam.cpp:329-332) -- so that the host ships one double per ghost and tests nothing itself
__global__ void __launch_bounds__(BLOCK) aeam_fill_kernel(double *__restrict__ a, int n, double v)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void __launch_bounds__(BLOCK) aeam_fp_gated_kernel(const double *__restrict__ rho, const double *__restrict__ fp,
                                                              int n, double *__restrict__ out)
{
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i < n) out[i] = rho[i] > 0.0000000000001 ? fp[i] : 0.0;
}

// two-phase API for hosts that own the halo exchange (LAMMPS: comm->forward_comm(this) in between)
extern "C" int b200md_aeam_density_phase(b200md_ctx *c, int nlocal, int nghost, const double *x,
                                         const int *type, double *rho_out, double *fp_out)
{
  if (!c) return B200MD_ERR_ARG;
  c->n_compute++;
  int rc = aeam_begin(c, nlocal, nghost, x, type, nullptr);
  if (rc) return rc;
  // option "peratom": the embedding energy of this phase is tallied per atom and handed out by the force phase
  if ((rc = b200md_peratom_begin(c, c->peratom_opt != 0))) return rc;
  if ((rc = b200md_aeam_density(c))) return rc;
  if (nlocal) {
    if (rho_out) CUDA_TRY(c, cudaMemcpyAsync(rho_out, c->rho.p, nlocal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    const double *fp_src = c->fp.p;
    if (fp_out && c->fp_gated) {
      CUDA_TRY(c, c->fp_tmp.reserve((size_t) nlocal + 8));
      LaunchScope ls(c, "aeam_gate");
      aeam_fp_gated_kernel<<<nblocks(nlocal, BLOCK), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, nlocal, c->fp_tmp.p);
      fp_src = c->fp_tmp.p;
    }
    if (fp_out) CUDA_TRY(c, cudaMemcpyAsync(fp_out, fp_src, nlocal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  int fl[16];
  CUDA_TRY(c, cudaMemcpyAsync(fl, c->flags.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->d2h_bytes += (long long) ((rho_out ? 1 : 0) + (fp_out ? 1 : 0)) * nlocal * sizeof(double);
  return aeam_check_flags(c, fl);
}

extern "C" int b200md_aeam_force_phase(b200md_ctx *c, const double *rho_all, const double *fp_all, int eflag,
                                       int vflag, double *f, double *eng_vdwl, double *virial)
{
  return b200md_aeam_force_phase_peratom(c, rho_all, fp_all, eflag, vflag, f, eng_vdwl, virial, nullptr, nullptr);
}

extern "C" int b200md_aeam_force_phase_peratom(b200md_ctx *c, const double *rho_all, const double *fp_all, int eflag,
                                               int vflag, double *f, double *eng_vdwl, double *virial, double *eatom,
                                               double *vatom)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->aeam_ready && c->inner_valid && f && fp_all && (rho_all || c->fp_gated),
            "aeam_force_phase: call the density phase first (rho_all may be NULL only with option fp_gated)");
  ARG_CHECK(c, !(eatom || vatom) || c->pa_e,
            "aeam_force_phase: per-atom output needs option \"peratom\" = 1 before the density phase");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int ng = c->nghost;
  if (ng) {
    // owned entries are already on the device; ghosts come from the host's halo exchange
    if (rho_all) {
      CUDA_TRY(c, cudaMemcpyAsync(c->rho.p + c->nlocal, rho_all + c->nlocal, ng * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      c->h2d_bytes += (long long) ng * sizeof(double);
    } else {    // gated fp: the ghosts' rho only has to pass the minrho test
      LaunchScope ls(c, "aeam_gate");
      aeam_fill_kernel<<<nblocks(ng, BLOCK), BLOCK, 0, c->stream>>>(c->rho.p + c->nlocal, ng, 1.0);
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->fp.p + c->nlocal, fp_all + c->nlocal, ng * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long) ng * sizeof(double);
  }
  int rc;
  if (!eatom && !vatom && !c->pa_e && !c->deterministic && c->d2h_chunks > 1 && c->nlocal >= c->d2h_min_atoms &&
      aeam_row_mode(c) == 2) {    // force download pipelined with the pair kernel, as in the one-shot call
    int fl[16];
    if ((rc = aeam_forces_download_pipelined(c, c->nlocal, eflag, vflag, f, eng_vdwl, virial, fl))) return rc;
    c->n_pipelined++;
    return aeam_check_flags(c, fl);
  }
  rc = b200md_aeam_forces(c, eflag, vflag, 0);
  if (rc) {
    c->pa_e = c->pa_v = nullptr;
    return rc;
  }
  if ((rc = b200md_peratom_finish(c, eatom, vatom))) return rc;
  return aeam_finish(c, eflag, vflag, f, eng_vdwl, virial);
}

extern "C" int b200md_aeam_get_rho_fp(b200md_ctx *c, int nlocal, double *rho, double *fp)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->aeam_ready && nlocal <= c->nlocal && c->rho.p, "aeam_get_rho_fp: nothing computed yet");
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (nlocal) {
    if (rho) CUDA_TRY(c, cudaMemcpyAsync(rho, c->rho.p, nlocal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (fp) CUDA_TRY(c, cudaMemcpyAsync(fp, c->fp.p, nlocal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200MD_OK;
}
