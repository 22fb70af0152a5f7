// b200md -- GPU binned cell list and full neighbor build ("full/bin/ghost", "full/bin/atomonly").
//
// Restates LAMMPS-core (stable_2Aug2023) NBinStandard::setup_bins/coord2bin/bin_atoms,
// NStencilFullBin3d / NStencilFullGhostBin3d::create and NPairFullBin / NPairFullBinGhost::build
// (the reference logs show these choices: log.rebomos-bulk.1:45-51).
// The contract is BIT-EXACT rows: same members, same order as the host build --
//   * bins are visited in stencil order (k,j,i nested, z slowest), atoms within a bin in ascending
//     local index (LAMMPS fills its bin linked lists in reverse, so they read ascending);
//   * distance tests use the reference operation order with no FMA contraction.
// K1 bin_index / bin_fill / bin_sort   thread per atom / per bin
// K2 neigh_rows<COUNT|FILL>            warp per row: stencil x-runs are contiguous in the sorted
//                                      atom array, lanes take consecutive atoms, ballot+popc append

#include "common.cuh"

#include <cmath>

#define BLOCK 256

struct BinGeom {
  double bboxlo[3], bboxhi[3];
  double bininv[3];
  int nbin[3];
  int mbinlo[3];
  int mbin[3];
  int mbins;
  int nruns;
  int ntypes;
  // per-row stencil clipping: origin of bin (0,0,0) of the local grid, bin size, padded largest cutoff squared, slack
  double org[3], bsz[3];
  double clipsq, eps;
};

// ---------------------------------------------------------------- device: coord2bin
__device__ __forceinline__ int coord2bin_dim(double x, double lo, double hi, double inv, int nbin)
{
  int ix;
  if (x >= hi) ix = (int) ((x - hi) * inv) + nbin;
  else if (x >= lo) {
    ix = (int) ((x - lo) * inv);
    ix = min(ix, nbin - 1);
  } else
    ix = (int) ((x - lo) * inv) - 1;
  return ix;
}

__global__ void __launch_bounds__(BLOCK) nb_pack_kernel(const double *__restrict__ x,
                                                        const int *__restrict__ type, int nall,
                                                        double4 *__restrict__ xt)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nall) return;
  xt[i] = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], (double) type[i]);
}

__global__ void __launch_bounds__(BLOCK) bin_index_kernel(const __grid_constant__ BinGeom g,
                                                          const double4 *__restrict__ xt, int nall,
                                                          int *__restrict__ bin_of, int *__restrict__ bin_count,
                                                          int *__restrict__ flags)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nall) return;
  const double4 p = xt[i];
  int ix = coord2bin_dim(p.x, g.bboxlo[0], g.bboxhi[0], g.bininv[0], g.nbin[0]) - g.mbinlo[0];
  int iy = coord2bin_dim(p.y, g.bboxlo[1], g.bboxhi[1], g.bininv[1], g.nbin[1]) - g.mbinlo[1];
  int iz = coord2bin_dim(p.z, g.bboxlo[2], g.bboxhi[2], g.bininv[2], g.nbin[2]) - g.mbinlo[2];
  if (ix < 0 || ix >= g.mbin[0] || iy < 0 || iy >= g.mbin[1] || iz < 0 || iz >= g.mbin[2] ||
      !(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z)) {
    flags[8] = 1;    // atom outside the bin grid (lost atom / non-numeric position)
    bin_of[i] = -1;
    return;
  }
  const int b = iz * g.mbin[1] * g.mbin[0] + iy * g.mbin[0] + ix;
  bin_of[i] = b;
  atomicAdd(&bin_count[b], 1);
}

__global__ void __launch_bounds__(BLOCK) bin_fill_kernel(const int *__restrict__ bin_of, int nall,
                                                         const int64_t *__restrict__ bin_start,
                                                         int *__restrict__ bin_cursor, int *__restrict__ bin_atoms)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nall) return;
  const int b = bin_of[i];
  if (b < 0) return;
  const int pos = atomicAdd(&bin_cursor[b], 1);
  bin_atoms[bin_start[b] + pos] = i;
}

// positions in bin order: the row kernels stream their candidates from it
__global__ void __launch_bounds__(BLOCK) bin_gather_kernel(const double4 *__restrict__ xt,
                                                           const int *__restrict__ bin_atoms, int n,
                                                           double4 *__restrict__ xs)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int j = bin_atoms[k];    // the tail is unwritten when an atom fell outside the grid (the build fails then)
  if ((unsigned) j < (unsigned) n) xs[k] = xt[j];
}

// ascending local index inside every bin (insertion sort; bins hold ~15 atoms)
__global__ void __launch_bounds__(BLOCK) bin_sort_kernel(const int64_t *__restrict__ bin_start, int mbins,
                                                         int *__restrict__ bin_atoms)
{
  int b = blockIdx.x * BLOCK + threadIdx.x;
  if (b >= mbins) return;
  const int s = (int) bin_start[b], e = (int) bin_start[b + 1];
  for (int a = s + 1; a < e; a++) {
    const int v = bin_atoms[a];
    int q = a - 1;
    while (q >= s && bin_atoms[q] > v) {
      bin_atoms[q + 1] = bin_atoms[q];
      q--;
    }
    bin_atoms[q + 1] = v;
  }
}

// ---------------------------------------------------------------- K2: rows
// A stencil "run" = maximal set of stencil bins with equal (j,k) and consecutive i: run[4] = {i0, i1, j, k}.
// MODE 0: count the rows (row_num).  MODE 1: fill them at the offsets of the scan in between (dense CSR, two stencil
// walks).  MODE 2: ONE walk into rows of a fixed stride (row i at i * stride; the stride comes from the longest row of
// the previous build of the same system): rows, counts and offsets in one pass -- a row longer than the stride raises
// flags[15] and the caller falls back to the two-pass build.  Row contents are identical in all modes.
template <int MODE, int NU, bool UNI>
__global__ void __launch_bounds__(BLOCK) neigh_rows_kernel(
    const __grid_constant__ BinGeom g, const double4 *__restrict__ xt, const double4 *__restrict__ xs,
    const int *__restrict__ bin_of, const int64_t *__restrict__ bin_start, const int *__restrict__ bin_atoms,
    const int4 *__restrict__ runs,
    const double *__restrict__ cutsq, const double *__restrict__ cutghostsq, int nlocal, int nrows,
    int64_t *__restrict__ row_off, int *__restrict__ row_num, int *__restrict__ row_val, int stride,
    int *__restrict__ flags, unsigned long long *__restrict__ total_used)
{
  constexpr bool FILL = MODE != 0;
  __shared__ int s_max, s_sum;
  __shared__ int2 s_run[BLOCK / 32][32];
  if (threadIdx.x == 0) s_max = s_sum = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = (int) (((size_t) blockIdx.x * BLOCK + threadIdx.x) >> 5);
  const bool live = i < nrows;
  const double4 xi = xt[live ? i : 0];
  const int itype = __double2int_rn(xi.w);
  const bool owned = i < nlocal;
  const double *cut = (owned ? cutsq : cutghostsq) + (size_t) itype * (g.ntypes + 1);
  const double cut_uni = UNI ? cut[1] : 0.0;    // UNI: every type pair has the same list cutoff (both pair styles here)
  const int ib = live ? bin_of[i] : -1;
  int n = 0;
  if (ib >= 0) {
    const int mx = g.mbin[0], my = g.mbin[1], mz = g.mbin[2];
    const int zb = ib / (my * mx);
    const int yb = (ib - zb * my * mx) / mx;
    const int xb = ib - zb * my * mx - yb * mx;
    const int64_t obase = MODE == 1 ? row_off[i] : (MODE == 2 ? (int64_t) i * stride : 0);
    const unsigned lt = (1u << lane) - 1u, le = 0xffffffffu >> (31 - lane);
    // position inside the own bin, single precision: the clip below is conservative by epsf = 1e-3 bin sizes
    const float fx = (float) (xi.x - (g.org[0] + xb * g.bsz[0])), fy = (float) (xi.y - (g.org[1] + yb * g.bsz[1])),
                fz = (float) (xi.z - (g.org[2] + zb * g.bsz[2]));
    const float bx = (float) g.bsz[0], by = (float) g.bsz[1], bz = (float) g.bsz[2], epsf = (float) g.eps;
    const float clipf = (float) g.clipsq * 1.0001f;
    // The candidates of a row are the concatenation, in stencil-run order, of the runs' slices of the bin-sorted atom
    // array.  32 runs at a time: lane k looks up run k's slice (all bin_start lookups in flight at once instead of one
    // dependent chain per run), the non-empty slices are compacted, a warp scan turns their lengths into offsets, and the
    // warp then walks the CONCATENATED candidate sequence 32 at a time.  A lane finds its (run, position) from a bit mask
    // of the run starts inside the 32-candidate window (one REDUX + one POPC; a 5-step search over the offsets cost 30 of
    // the ~60 instructions of a trip).  Same candidates in the same order as the run-by-run walk (rows stay bit-exact).
    for (int r0 = 0; r0 < g.nruns; r0 += 32) {
      int len = 0, s0 = 0;
      if (r0 + lane < g.nruns) {
        const int4 run = runs[r0 + lane];
        int x0 = xb + run.x, x1 = xb + run.y;
        const int y = yb + run.z, z = zb + run.w;
        // ghost atoms: skip stencil bins outside the local bin grid (NPairFullBinGhost).  For owned
        // atoms the grid always covers the stencil, so the same clip is a no-op that guards memory.
        if (y >= 0 && y < my && z >= 0 && z < mz) {
          // Clip the run to the bins this atom's cutoff sphere can reach (the stencil is the union over all positions
          // inside the atom's bin: 125 bins against ~85 for one position).  Skipped bins hold no neighbor, so the row
          // keeps its members and their order.  Bin y covers [org + y b, org + (y + 1) b) up to the rounding of
          // coord2bin; epsf (1e-3 b) covers that and the single-precision arithmetic here.
          float gy = 0.0f, gz = 0.0f;
          if (run.z > 0) gy = run.z * by - fy;
          else if (run.z < 0) gy = fy + (-run.z - 1) * by;
          if (run.w > 0) gz = run.w * bz - fz;
          else if (run.w < 0) gz = fz + (-run.w - 1) * bz;
          gy = fmaxf(gy - epsf, 0.0f);
          gz = fmaxf(gz - epsf, 0.0f);
          const float rem = clipf - gy * gy - gz * gz;
          if (rem < 0.0f) x1 = x0 - 1;
          else {
            const float rx = sqrtf(rem) + epsf;
            x0 = max(x0, xb + (int) floorf((fx - rx) / bx));
            x1 = min(x1, xb + (int) floorf((fx + rx) / bx));
          }
          x0 = max(x0, 0);
          x1 = min(x1, mx - 1);
          if (x0 <= x1) {
            const int b0 = z * my * mx + y * mx + x0;
            s0 = (int) bin_start[b0];
            len = (int) bin_start[b0 + (x1 - x0) + 1] - s0;
          }
        }
      }
      // compact the non-empty slices (their start offsets are then strictly increasing)
      const unsigned ne = __ballot_sync(0xffffffffu, len > 0);
      const int nrun = __popc(ne);
      if (len > 0) s_run[wid][__popc(ne & lt)] = make_int2(s0, len);
      __syncwarp();
      s0 = 0, len = 0;
      if (lane < nrun) {
        const int2 r = s_run[wid][lane];
        s0 = r.x, len = r.y;
      }
      __syncwarp();
      int inc = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const int total = __shfl_sync(0xffffffffu, inc, 31);
      const int exc = inc - len;
      int kb = 0;    // slices that start before the window
      for (int c0 = 0; c0 < total; c0 += 32 * NU) {
        int jv[NU];
        double4 xv[NU];
#pragma unroll
        for (int u = 0; u < NU; u++) {
          const int w0 = c0 + u * 32, cidx = w0 + lane;
          const int rel = exc - w0;
          const unsigned starts = __reduce_or_sync(0xffffffffu, (lane < nrun && rel >= 0 && rel < 32) ? (1u << rel) : 0u);
          const int k = kb + __popc(starts & le) - 1;    // last slice that starts at or before cidx
          kb += __popc(starts);
          const int sk = __shfl_sync(0xffffffffu, s0, k & 31), ek = __shfl_sync(0xffffffffu, exc, k & 31);
          jv[u] = -1;
          if (cidx < total) {
            // candidates are consecutive entries of the bin-sorted arrays: both loads are coalesced streams (xs =
            // positions in bin order; gathering xt[j] cost one L1 wavefront per lane)
            const int pos = sk + (cidx - ek);
            jv[u] = bin_atoms[pos];
            xv[u] = ld_sector(xs + pos);
          }
        }
#pragma unroll
        for (int u = 0; u < NU; u++) {
          bool keep = false;
          if (jv[u] >= 0 && jv[u] != i) {
            const double dx = xi.x - xv[u].x, dy = xi.y - xv[u].y, dz = xi.z - xv[u].z;
            const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            keep = rsq <= (UNI ? cut_uni : cut[__double2int_rn(xv[u].w)]);
          }
          const unsigned mk = __ballot_sync(0xffffffffu, keep);
          if (FILL && keep && (MODE == 1 || n + __popc(mk & lt) < stride)) row_val[obase + n + __popc(mk & lt)] = jv[u];
          n += __popc(mk);
        }
      }
    }
  }
  if (lane == 0 && live) {
    if (MODE != 1) {
      row_num[i] = n;
      atomicMax(&s_max, n);
      if (MODE == 2) atomicAdd(&s_sum, n);
    }
    if (MODE == 2) {
      row_off[i] = (int64_t) i * stride;
      if (i == nrows - 1) row_off[nrows] = (int64_t) nrows * stride;
      if (n > stride) flags[15] = 1;
    }
  }
  if (MODE != 1) {    // longest row (the stride hint of the next build): one global atomic per block, not per row
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicMax(&flags[10], s_max);
      if (MODE == 2) atomicAdd(total_used, (unsigned long long) s_sum);    // entries the strided rows really hold
    }
  }
}

// ---------------------------------------------------------------- host: bins + stencil (NBinStandard / NStencil)
static void triclinic_bbox(const b200md_box &b, const double *lo, const double *hi, double *blo, double *bhi)
{
  const double h0 = b.boxhi[0] - b.boxlo[0], h1 = b.boxhi[1] - b.boxlo[1], h2 = b.boxhi[2] - b.boxlo[2];
  for (int d = 0; d < 3; d++) {
    blo[d] = 1.0e30;
    bhi[d] = -1.0e30;
  }
  for (int c = 0; c < 8; c++) {
    const double l0 = (c & 1) ? hi[0] : lo[0], l1 = (c & 2) ? hi[1] : lo[1], l2 = (c & 4) ? hi[2] : lo[2];
    double x[3];
    x[0] = h0 * l0 + b.xy * l1 + b.xz * l2 + b.boxlo[0];    // Domain::lamda2x
    x[1] = h1 * l1 + b.yz * l2 + b.boxlo[1];
    x[2] = h2 * l2 + b.boxlo[2];
    for (int d = 0; d < 3; d++) {
      blo[d] = x[d] < blo[d] ? x[d] : blo[d];
      bhi[d] = x[d] > bhi[d] ? x[d] : bhi[d];
    }
  }
}

int b200md_neigh_setup_bins(b200md_ctx *c, const b200md_box &box, int ntypes, BinGeom &g,
                            std::vector<int4> &runs)
{
  const double SMALL = 1.0e-6, CUT2BIN_RATIO = 100.0;
  double bsublo[3], bsubhi[3];
  if (!box.triclinic) {
    for (int d = 0; d < 3; d++) {
      g.bboxlo[d] = box.boxlo[d];
      g.bboxhi[d] = box.boxhi[d];
      bsublo[d] = box.sublo[d] - box.cutghost[d];
      bsubhi[d] = box.subhi[d] + box.cutghost[d];
    }
  } else {
    // Domain::set_global_box: boxlo_bound / boxhi_bound
    g.bboxlo[0] = fmin(box.boxlo[0], box.boxlo[0] + box.xy);
    g.bboxlo[0] = fmin(g.bboxlo[0], g.bboxlo[0] + box.xz);
    g.bboxlo[1] = fmin(box.boxlo[1], box.boxlo[1] + box.yz);
    g.bboxlo[2] = box.boxlo[2];
    g.bboxhi[0] = fmax(box.boxhi[0], box.boxhi[0] + box.xy);
    g.bboxhi[0] = fmax(g.bboxhi[0], g.bboxhi[0] + box.xz);
    g.bboxhi[1] = fmax(box.boxhi[1], box.boxhi[1] + box.yz);
    g.bboxhi[2] = box.boxhi[2];
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = box.sublo[d] - box.cutghost[d];
      hi[d] = box.subhi[d] + box.cutghost[d];
    }
    triclinic_bbox(box, lo, hi, bsublo, bsubhi);
  }
  double bbox[3], binsize[3];
  for (int d = 0; d < 3; d++) bbox[d] = g.bboxhi[d] - g.bboxlo[d];
  double binsize_optimal = 0.5 * box.cutneighmax;
  if (binsize_optimal == 0.0) binsize_optimal = bbox[0];
  const double binsizeinv = 1.0 / binsize_optimal;
  for (int d = 0; d < 3; d++) {
    ARG_CHECK(c, bbox[d] * binsizeinv < 2.0e9, "Domain too large for neighbor bins");
    g.nbin[d] = static_cast<int>(bbox[d] * binsizeinv);
    if (g.nbin[d] == 0) g.nbin[d] = 1;
    binsize[d] = bbox[d] / g.nbin[d];
    g.bininv[d] = 1.0 / binsize[d];
    ARG_CHECK(c, binsize_optimal * g.bininv[d] <= CUT2BIN_RATIO, "Cannot use neighbor bins - box size << cutoff");
    double coord = bsublo[d] - SMALL * bbox[d];
    int lo = static_cast<int>((coord - g.bboxlo[d]) * g.bininv[d]);
    if (coord < g.bboxlo[d]) lo = lo - 1;
    coord = bsubhi[d] + SMALL * bbox[d];
    int hi = static_cast<int>((coord - g.bboxlo[d]) * g.bininv[d]);
    lo = lo - 1;
    hi = hi + 1;
    g.mbinlo[d] = lo;
    g.mbin[d] = hi - lo + 1;
  }
  const long long bbin = (long long) g.mbin[0] * g.mbin[1] * g.mbin[2] + 1;
  ARG_CHECK(c, bbin < 2000000000LL, "Too many neighbor bins");
  g.mbins = (int) bbin;
  g.ntypes = ntypes;
  for (int d = 0; d < 3; d++) {
    g.bsz[d] = binsize[d];
    g.org[d] = g.bboxlo[d] + g.mbinlo[d] * binsize[d];
  }
  g.eps = 1.0e-6 * fmax(binsize[0], fmax(binsize[1], binsize[2]));
  g.clipsq = box.cutneighmax * box.cutneighmax;    // raised to the largest list cutoff by the caller

  // stencil (NStencil::create_setup + NStencilFull[Ghost]Bin3d::create), grouped into x-runs
  int s[3];
  for (int d = 0; d < 3; d++) {
    s[d] = static_cast<int>(box.cutneighmax * g.bininv[d]);
    if (s[d] * binsize[d] < box.cutneighmax) s[d]++;
  }
  auto bin_distance = [&](int i, int j, int k) {
    double delx, dely, delz;
    if (i > 0) delx = (i - 1) * binsize[0];
    else if (i == 0) delx = 0.0;
    else delx = (i + 1) * binsize[0];
    if (j > 0) dely = (j - 1) * binsize[1];
    else if (j == 0) dely = 0.0;
    else dely = (j + 1) * binsize[1];
    if (k > 0) delz = (k - 1) * binsize[2];
    else if (k == 0) delz = 0.0;
    else delz = (k + 1) * binsize[2];
    return delx * delx + dely * dely + delz * delz;
  };
  const double cutsqmax = box.cutneighmax * box.cutneighmax;
  runs.clear();
  for (int k = -s[2]; k <= s[2]; k++)
    for (int j = -s[1]; j <= s[1]; j++) {
      int start = 0;
      bool open = false;
      for (int i = -s[0]; i <= s[0]; i++) {
        const bool in = bin_distance(i, j, k) < cutsqmax;
        if (in && !open) {
          start = i;
          open = true;
        }
        if (!in && open) {
          runs.push_back(make_int4(start, i - 1, j, k));
          open = false;
        }
      }
      if (open) runs.push_back(make_int4(start, s[0], j, k));
    }
  g.nruns = (int) runs.size();
  return B200MD_OK;
}

struct NeighScratch {
  DevBuf<double4> xt, xs;
  DevBuf<unsigned long long> used;
  DevBuf<int4> runs;
  DevBuf<double> cutsq, cutghostsq;
  DevBuf<int64_t> bin_start;
};
// per-context scratch (contexts are driven by different host threads in multi-rank runs: no shared state)
static NeighScratch &scratch_of(b200md_ctx *c)
{
  if (!c->neigh_scratch) c->neigh_scratch = new NeighScratch();
  return *c->neigh_scratch;
}

void b200md_neigh_forget(b200md_ctx *c)
{
  if (!c->neigh_scratch) return;
  NeighScratch &S = *c->neigh_scratch;
  S.xt.release();
  S.xs.release();
  S.used.release();
  S.runs.release();
  S.cutsq.release();
  S.cutghostsq.release();
  S.bin_start.release();
  delete c->neigh_scratch;
  c->neigh_scratch = nullptr;
}

static inline int nblocks(long long n, int per) { return (int) ((n + per - 1) / per); }
// option "neigh_unroll": trips of 32 candidates a lane keeps in flight (A/B knob; default 1)
#define NEIGH_LAUNCH(MODE, ...)                                                                                     \
  do {                                                                                                              \
    const int nb_ = nblocks((long long) nrows * 32, BLOCK);                                                         \
    if (uniform_cut) neigh_rows_kernel<MODE, 1, true><<<nb_, BLOCK, 0, c->stream>>>(__VA_ARGS__);                   \
    else if (c->neigh_unroll >= 2) neigh_rows_kernel<MODE, 2, false><<<nb_, BLOCK, 0, c->stream>>>(__VA_ARGS__);    \
    else neigh_rows_kernel<MODE, 1, false><<<nb_, BLOCK, 0, c->stream>>>(__VA_ARGS__);                              \
  } while (0)

// Build from device-resident positions xt = {x,y,z,(double) type}.  Leaves the dense CSR master list in
// c->list_{off,num,val}.  The only host round trip is the total entry count (to size the value array).
int b200md_neigh_build_device(b200md_ctx *c, const b200md_box &box, int ntypes, const double *cutneighsq_h,
                              const double *cutneighghostsq_h, int nlocal, int nghost, const double4 *xt,
                              int ghost_rows, double skin, bool one_pass)
{
  NeighScratch &S = scratch_of(c);
  BinGeom g;
  std::vector<int4> runs;
  int rc = b200md_neigh_setup_bins(c, box, ntypes, g, runs);
  if (rc) return rc;
  const int nall = nlocal + nghost;
  const int nrows = ghost_rows ? nall : nlocal;
  const size_t nt2 = (size_t) (ntypes + 1) * (ntypes + 1);
  for (size_t k = 0; k < nt2; k++) {    // the clip radius covers every cutoff a row is filtered with
    g.clipsq = fmax(g.clipsq, cutneighsq_h[k]);
    if (cutneighghostsq_h) g.clipsq = fmax(g.clipsq, cutneighghostsq_h[k]);
  }
  g.clipsq *= 1.0 + 1.0e-12;
  g.eps = 1.0e-3 * fmax(g.bsz[0], fmax(g.bsz[1], g.bsz[2]));
  // one cutoff for every type pair (owned and ghost rows alike)?  Then the rows are filtered without the table lookup.
  bool uniform_cut = true;
  for (int a = 1; a <= ntypes; a++)
    for (int b = 1; b <= ntypes; b++) {
      const size_t k = (size_t) a * (ntypes + 1) + b;
      if (cutneighsq_h[k] != cutneighsq_h[(size_t) ntypes + 2]) uniform_cut = false;
      if (cutneighghostsq_h && cutneighghostsq_h[k] != cutneighsq_h[(size_t) ntypes + 2]) uniform_cut = false;
    }
  CUDA_TRY(c, S.xs.reserve((size_t) nall + 8));
  CUDA_TRY(c, S.used.reserve(2));
  CUDA_TRY(c, S.runs.reserve(runs.size() + 8));
  CUDA_TRY(c, S.cutsq.reserve(nt2 + 8));
  CUDA_TRY(c, S.cutghostsq.reserve(nt2 + 8));
  CUDA_TRY(c, S.bin_start.reserve((size_t) g.mbins + 8));
  CUDA_TRY(c, c->bin_of.reserve((size_t) nall + 32));
  CUDA_TRY(c, c->bin_count.reserve((size_t) g.mbins + 8));
  CUDA_TRY(c, c->bin_atoms.reserve((size_t) nall + 32));
  CUDA_TRY(c, c->list_off.reserve((size_t) nrows + 2));
  CUDA_TRY(c, c->list_num.reserve((size_t) nrows + 32));
  CUDA_TRY(c, cudaMemcpyAsync(S.runs.p, runs.data(), runs.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(S.cutsq.p, cutneighsq_h, nt2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(S.cutghostsq.p, cutneighghostsq_h ? cutneighghostsq_h : cutneighsq_h,
                              nt2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->bin_count.p, 0, (size_t) g.mbins * sizeof(int), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 8, 0, sizeof(int), c->stream));
  if (nall) {
    {
      LaunchScope ls(c, "bin_index");
      bin_index_kernel<<<nblocks(nall, BLOCK), BLOCK, 0, c->stream>>>(g, xt, nall, c->bin_of.p, c->bin_count.p,
                                                                     c->flags.p);
    }
    rc = b200md_exclusive_scan_i64(c, c->bin_count.p, S.bin_start.p, g.mbins, 1);
    if (rc) return rc;
    CUDA_TRY(c, cudaMemsetAsync(c->bin_count.p, 0, (size_t) g.mbins * sizeof(int), c->stream));
    {
      LaunchScope ls(c, "bin_fill");
      bin_fill_kernel<<<nblocks(nall, BLOCK), BLOCK, 0, c->stream>>>(c->bin_of.p, nall, S.bin_start.p,
                                                                    c->bin_count.p, c->bin_atoms.p);
    }
    {
      LaunchScope ls(c, "bin_sort");
      bin_sort_kernel<<<nblocks(g.mbins, BLOCK), BLOCK, 0, c->stream>>>(S.bin_start.p, g.mbins, c->bin_atoms.p);
    }
    {
      LaunchScope ls(c, "bin_gather");
      bin_gather_kernel<<<nblocks(nall, BLOCK), BLOCK, 0, c->stream>>>(xt, c->bin_atoms.p, nall, S.xs.p);
    }
  }
  int64_t total = 0, used = -1;
  CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 10, 0, sizeof(int), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 15, 0, sizeof(int), c->stream));
  bool done = false;
  if (nrows && one_pass && c->list_maxnum > 0) {
    // resident loop, second and later builds: one stencil walk into rows of a fixed stride
    // sticky and generous: rows of a lattice that is still heating up keep growing (fcc Al at 863 K: 98 -> 131 entries),
    // and a stride that follows the latest maximum re-allocates GB-sized buffers inside somebody's timed region
    int stride = ((int) (c->list_maxnum * (c->list_stride ? 1.25 : 1.5)) + 32) & ~7;    // first time: a cold lattice
    if (stride < c->list_stride) stride = c->list_stride;
    c->list_stride = stride;
    CUDA_TRY(c, cudaMemsetAsync(S.used.p, 0, sizeof(unsigned long long), c->stream));
    CUDA_TRY(c, c->list_val.reserve((size_t) nrows * stride + 64));
    {
      LaunchScope ls(c, "neigh_fill");
      NEIGH_LAUNCH(2, g, xt, S.xs.p, c->bin_of.p, S.bin_start.p, c->bin_atoms.p, S.runs.p, S.cutsq.p, S.cutghostsq.p, nlocal, nrows,
          c->list_off.p, c->list_num.p, c->list_val.p, stride, c->flags.p, S.used.p);
    }
    int fl[8];
    unsigned long long used_h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&used_h, S.used.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(fl, c->flags.p + 8, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (fl[0]) {
      c->fail("neighbor build: atom outside the bin grid (lost atom or non-numeric position)");
      return B200MD_ERR_ARG;
    }
    c->list_maxnum = fl[2];
    if (!fl[7]) {
      total = (int64_t) nrows * stride;    // capacity the rows occupy
      used = (int64_t) used_h;             // entries they hold (derived lists size themselves from it)
      done = true;
    } else {
      CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 15, 0, sizeof(int), c->stream));
      CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 10, 0, sizeof(int), c->stream));
    }
  }
  if (nrows && !done) {
    {
      LaunchScope ls(c, "neigh_count");
      NEIGH_LAUNCH(0, g, xt, S.xs.p, c->bin_of.p, S.bin_start.p, c->bin_atoms.p, S.runs.p, S.cutsq.p, S.cutghostsq.p, nlocal, nrows,
          nullptr, c->list_num.p, nullptr, 0, c->flags.p, nullptr);
    }
    rc = b200md_exclusive_scan_i64(c, c->list_num.p, c->list_off.p, nrows, 1);
    if (rc) return rc;
    int flag8 = 0, maxnum = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&total, c->list_off.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(&flag8, c->flags.p + 8, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(&maxnum, c->flags.p + 10, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->list_maxnum = maxnum;
    if (flag8) {
      c->fail("neighbor build: atom outside the bin grid (lost atom or non-numeric position)");
      return B200MD_ERR_ARG;
    }
    CUDA_TRY(c, c->list_val.reserve((size_t) total + 64));
    {
      LaunchScope ls(c, "neigh_fill");
      NEIGH_LAUNCH(1, g, xt, S.xs.p, c->bin_of.p, S.bin_start.p, c->bin_atoms.p, S.runs.p, S.cutsq.p, S.cutghostsq.p, nlocal, nrows,
          c->list_off.p, c->list_num.p, c->list_val.p, 0, c->flags.p, nullptr);
    }
    CUDA_TRY(c, cudaGetLastError());
  } else if (!nrows) {
    CUDA_TRY(c, cudaMemsetAsync(c->list_off.p, 0, sizeof(int64_t), c->stream));
  }
  c->list_inum = nlocal;
  c->list_gnum = ghost_rows ? nghost : 0;
  c->list_entries = total;
  c->list_entries_used = used >= 0 ? used : total;
  c->skin = skin;
  c->list_valid = true;
  c->inner_valid = false;
  c->n_list_upload++;    // counts every master list the context received, built here or handed over
  return B200MD_OK;
}

extern "C" int b200md_neigh_build(b200md_ctx *c, const b200md_box *box, int ntypes, const double *cutneighsq,
                                  const double *cutneighghostsq, int nlocal, int nghost, const double *x,
                                  const int *type, int ghost_rows, double skin)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, box && cutneighsq && x && type && ntypes >= 1 && nlocal >= 0 && nghost >= 0 && skin >= 0.0,
            "neigh_build: bad arguments");
  ARG_CHECK(c, box->cutneighmax > 0.0, "neigh_build: cutneighmax must be > 0");
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->type_on_device = c->tag_on_device = false;    // a new list means the host may have re-sorted its atoms
  int rc = b200md_upload_atoms(c, nlocal, nghost, x, type, nullptr);
  if (rc) return rc;
  NeighScratch &S = scratch_of(c);
  const int nall = nlocal + nghost;
  CUDA_TRY(c, S.xt.reserve((size_t) nall + 8));
  if (nall) {
    LaunchScope ls(c, "nb_pack");
    nb_pack_kernel<<<nblocks(nall, BLOCK), BLOCK, 0, c->stream>>>(c->x_aos.p, c->type.p, nall, S.xt.p);
  }
  rc = b200md_neigh_build_device(c, *box, ntypes, cutneighsq, cutneighghostsq, nlocal, nghost, S.xt.p, ghost_rows, skin, false);
  if (rc) return rc;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200MD_OK;
}
