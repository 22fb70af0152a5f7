// b200md -- common declarations for the sm_100a force-path library.
// Context object, growable device buffers, error plumbing, warp helpers.
#pragma once

#include "b200md.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#define B200MD_NEIGHMASK 0x1FFFFFFF
#define B200MD_SHORT_WIDTH 64      // max entries of a short (REBO-range + margin) row
#define B200MD_MAX_REBO 16         // max REBO neighbors of one atom (bulk MoS2: 12)
#define B200MD_MAX_TYPES 8

// ---------------------------------------------------------------- errors
#define CUDA_TRY(ctx, call)                                                                       \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      (ctx)->fail(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                  std::to_string(__LINE__) + " (" #call ")");                                     \
      return B200MD_ERR_CUDA;                                                                     \
    }                                                                                             \
  } while (0)

#define ARG_CHECK(ctx, cond, msg)                                                                 \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      (ctx)->fail(std::string("argument error: ") + (msg));                                       \
      return B200MD_ERR_ARG;                                                                      \
    }                                                                                             \
  } while (0)

// ---------------------------------------------------------------- device buffer
template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;    // elements
  cudaError_t reserve(size_t n, bool keep = false, cudaStream_t st = 0)
  {
    if (n <= cap) return cudaSuccess;
    size_t ncap = n + n / 4 + 256;    // a quarter of headroom: neighbor rows grow ~20 % from a cold lattice to a hot one
    static const bool trace = getenv("B200MD_ALLOC_TRACE") != nullptr;    // diagnostic: growth inside a timed region?
    if (trace) fprintf(stderr, "[alloc] %zu -> %zu elements of %zu bytes (%.1f MB)\n", cap, ncap, sizeof(T), ncap * sizeof(T) / 1.0e6);
    T *q = nullptr;
    cudaError_t e = cudaMalloc((void **) &q, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return e;
      cudaStreamSynchronize(st);
    }
    if (p) cudaFree(p);
    p = q;
    cap = ncap;
    return cudaSuccess;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T> struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) return cudaSuccess;
    size_t ncap = n + n / 8 + 256;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost((void **) &p, ncap * sizeof(T));
    if (e == cudaSuccess) cap = ncap;
    return e;
  }
  void release()
  {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// ---------------------------------------------------------------- potential parameter blocks
// REBOMoS constants, passed BY VALUE to kernels (lives in the kernel-parameter constant bank).
struct RebomosDev {
  double rcmin[4], rcmax[4], rcmaxsq[4], rcw[4];    // rcw = rcmax-rcmin
  double Q[4], alpha[4], A[4], BIJc[4], Beta[4];
  double b[2][7], bg[2][7];    // [elem][order]
  double a[2][4];
  double rcLJmin[4], rcLJmax[4], sig95[4];    // sig95 = 0.95*sigma
  // exact rsq equivalents of the reference's tests on rij = sqrt(rsq) (sqrt is monotonic and correctly
  // rounded on both sides): rij > rcLJmax <=> rsq >= lj_out_hi ; rij < rcLJmin <=> rsq < lj_in_lo ;
  // rij >= 0.95 sigma <=> rsq >= lj_s95
  double lj_out_hi[4], lj_in_lo[4], lj_s95[4];
  double lj1[4], lj2[4], lj3[4], lj4[4];
  double c2[4], c3[4];    // cubic taper coefficients (pair_rebomos.cpp:533-538)
  double shortsq[4];      // (rcmax + margin)^2
  double ljsq[4];         // (rcLJmax + margin)^2
};

struct AeamDev {
  int nel, nnonangular;
  // per type pair (row-major i*nel+j), nel <= 4
  double cut[16], rdr[16];
  int nr[16];
  int rhor_off[16];    // row offset of table (i,j) in the packed rhor spline array
  int z2r_off[16];     // row offset of the z2r table used by pair (i,j)
  double rdrho[4];
  int nrho[4];
  int frho_off[4];
  double cutsq_list[16];    // (cut+margin)^2 for the inner list
  int pair_off[16];         // row offset (in 64-byte rows) of the fused {rhor | z2r} table of pair (i,j)
  double cut_gt_sq[16];     // smallest rsq with sqrt(rsq) > cut[i][j]: `rsq >= this` is the reference's `r > cut`
  int z2r_n[16];            // rows of the z2r table pair (i,j) reads (the row index is clamped to it)
  double z2r_k[16];         // dr[i][j] / dr[max(i,j)][min(i,j)]: the z2r derivative is scaled with the lower-triangle dr
  int asym_dr;              // some z2r_k != 1: the file's dr matrix is not symmetric
};

// ---------------------------------------------------------------- context
struct TimedLaunch {
  int name_id;
  cudaEvent_t a, b;
};
struct KernelStat {
  double total_ms = 0.0, last_ms = 0.0;
  long long count = 0;
};

// Interior/boundary split of the owned centers (resident multi-stream loop): a center is INTERIOR when it is farther
// than `r` from every face of this rank's sub-domain, so that none of its list candidates is a ghost and its forces
// can be computed while the forward halo is still in flight.  lo/hi/r are in lamda units for a triclinic box.
struct SplitGeom {
  int on = 0, triclinic = 0;
  double boxlo[3] = {0, 0, 0}, h_inv[6] = {0, 0, 0, 0, 0, 0};
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, r[3] = {0, 0, 0};
};

struct SystemState;    // resident MD system (system.cu)
struct NeighScratch;   // device-build scratch (neigh.cu)
struct AeamHost;       // 7-coefficient spline tables kept for b200md_aeam_get_spline (aeam.cu)

#define B200MD_MAX_D2H_CHUNKS 16

struct b200md_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t up_stream = nullptr;      // plugin mode: position upload in pieces beside the bond-order launches
  cudaEvent_t up_ev[B200MD_MAX_D2H_CHUNKS + 2] = {};
  int h2d_chunks = 6, h2d_K = 0, h2d_ramp = 4;
  bool h2d_ready = false;
  int h2d_need[B200MD_MAX_D2H_CHUNKS + 1] = {}, h2d_t[B200MD_MAX_D2H_CHUNKS + 2] = {};
  // stragglers of the upload pipeline: atoms named by the short rows of a range from more than one piece away (wrapped
  // through a periodic face since the last sort); their positions travel ahead of the pieces
  int n_strag = 0;
  DevBuf<int> strag_flag, strag_list;
  DevBuf<double> strag_dev;
  PinBuf<double> strag_pin;
  std::vector<int> strag_host;
  cudaStream_t copy_stream = nullptr;    // plugin mode: D2H of finished force ranges runs beside the remaining kernels
  cudaEvent_t copy_ev[B200MD_MAX_D2H_CHUNKS + 2] = {};      // "range finished on the compute stream"
  cudaEvent_t copy_done[B200MD_MAX_D2H_CHUNKS + 2] = {};    // "range has arrived on the host"
  struct D2HRange {
    int slot;
    size_t lo, hi;
  };
  std::vector<D2HRange> d2h_ranges;
  std::string err;
  int num_sms = 148;

  // options
  int deterministic = 0;
  double margin_opt = 0.0;    // 0 -> use skin
  double margin_t_opt = 0.4;  // margin of the tight rows in A ("margin_tight", mA); 0 = no third level
  int sync_timing = 0;
  int f_overwrite = 0;
  int peratom_opt = 0;    // AEAM two-phase API: tally per-atom energy/virial from the density phase on
  int ang_ctas = 10;     // AEAM angular launches: CTAs (4 warps, one angular center each) per SM; latency-bound kernels:
                         // 2 -> 10 CTAs/SM: force_ang 0.30 -> 0.16 ms, density_ang 0.086 -> 0.041 ms at 15 360 Si atoms
  int lj_pairs = 1;      // LJ over pairs of neighboring centers sharing one union row (0: one row per center)
  int aeam_cluster = 2;  // AEAM row form.  2 (default): one row per center, the density pass hands f'(r) of every entry to the
                         // force pass (one spline gather per pair and pass).  1: clusters of 4 consecutive centers share one
                         // union row (fewest gathered sectors, but latency-bound: slower, profiles/r02_aeam_kernels.md).
                         // 0: one row per center, force pass re-gathers the fused {rho' | phi} row (the round-1 kernels)
  int force_rebuild = 0;     // resident loop: the next step rebuilds the master list whatever the displacements
  int aeam_variant = 0;      // tuning experiments: bit 0 = force kernel 2 entries per lane and trip, bit 1 = density 1
  int aeam_sort_rows = 0;    // AEAM cluster rows sorted by atom index (adjacent lanes then read adjacent sectors)
  int d2h_min_atoms = 65536;    // below this the ranged path is all launch latency
  int d2h_chunks = 4;    // plugin mode: owned-atom index ranges whose forces go home while the next range computes
  int p2p_halo = 1;    // multi-GPU halo through peer memory (CUDA IPC) when available; 0 = NCCL send/recv only
  long long n_p2p = 0;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  // counters
  long long n_launch = 0, n_list_upload = 0, n_inner_rebuild = 0, h2d_bytes = 0, d2h_bytes = 0;
  long long n_lj_entries = 0, n_short_entries = 0;
  long long n_compute = 0;    // force computations through the host-buffer entry points
  long long n_pipelined = 0, n_redo = 0;    // plugin-mode calls through the pipelined path / recomputed after a refresh
  // per-launch CUDA events while "sync_timing" is on; folded into kstat by b200md_collect_timers()
  std::vector<std::string> kname;
  std::map<std::string, int> kname_id;
  std::vector<TimedLaunch> pending;
  std::vector<cudaEvent_t> event_pool;
  std::map<std::string, KernelStat> kstat;

  // ---- atoms (device)
  int nlocal = 0, nghost = 0, nall = 0;
  bool type_on_device = false, tag_on_device = false;    // host type/tag arrays already uploaded for this list
  int ids_nall = -1;
  DevBuf<double> x_aos;      // [nall*3] staging of host x
  DevBuf<double4> xq;        // packed {x,y,z,elem-as-double}
  DevBuf<double> f;          // [nall*3]
  DevBuf<int> type, tag;
  PinBuf<double> pin_f;      // D2H staging
  PinBuf<double> pin_scal;   // small scalar results
  DevBuf<double> eatom_d, vatom_d;    // per-atom energy [nall] / virial [nall*6] of one call (on request)
  double *pa_e = nullptr, *pa_v = nullptr;    // non-null while a per-atom call is in flight
  PinBuf<double> pin_pa;
  DevBuf<double> scal;       // [64] device accumulators: 0 = energy, 1..6 virial, 8.. misc
  DevBuf<int> flags;         // [16] device flags: 0 = overflow, 1 = inner rebuild needed, 2.. counters

  // ---- master neighbor list (dense CSR on device)
  int list_inum = 0, list_gnum = 0;
  int64_t list_entries = 0;          // extent of list_val the rows occupy (dense CSR: their sum; strided rows: rows x stride)
  int64_t list_entries_used = 0;     // entries the rows hold
  int list_stride = 0;               // sticky stride of the one-pass builds
  int list_maxnum = 0;       // longest row of the last device build: stride hint of the next one-pass build
  int one_pass_neigh = 1;    // option "one_pass_neigh": resident-loop rebuilds walk the stencil once (fixed-stride rows)
  double skin = 0.0;
  double margin = 0.0;       // margin actually used by the inner lists
  bool list_valid = false, inner_valid = false;
  DevBuf<int64_t> list_off;  // [rows+1]
  DevBuf<int> list_num;      // [rows]
  DevBuf<int> list_val;      // [entries]
  DevBuf<double> xhold;      // [nall*3] positions at inner-list build

  // ---- REBOMoS
  bool rebomos_ready = false;
  RebomosDev rp;
  int ntypes = 0;
  int map_h[B200MD_MAX_TYPES + 1];
  DevBuf<int> map_d;
  // inner lists
  DevBuf<int> short_idx;     // [rows][B200MD_SHORT_WIDTH] candidates within rcmax + margin, master-row order
  DevBuf<int> short_num;     // [rows]
  DevBuf<int64_t> lj_off;    // [inum+1] 8-aligned row offsets
  DevBuf<int> lj_num;        // [inum]
  DevBuf<int> lj_val;        // directed LJ-window rows
  int64_t lj_capacity = 0;
  // third list level (GPU-resident loop only): the rows the force kernels actually stream, filtered from the inner
  // ("wide") rows above to rcut + margin_t and re-derived from them -- a pass over 128 entries per atom instead of the
  // 496 of the master rows -- whenever an atom moved margin_t/2
  DevBuf<int> short_idx_t, short_num_t, lj_val_t, lj_num_t;
  DevBuf<double> xhold_t;    // [nall*4] positions at the last tight derive
  double margin_t = 0.0;
  bool tight_valid = false;
  bool sys_owns_tight = false;       // the tight rows on the device were derived by the resident loop
  bool tight_derive_pending = false; // plugin mode: a deferred derive is in flight on the compute stream (ev_tight)
  cudaEvent_t ev_tight = nullptr;
  long long n_tight = 0;
  DevBuf<int> ljp_ab;        // pair mode: {a, b} atom indices of every pair row, [2][P] by element
  int ljp_P = 0;             // pair slots per element
  DevBuf<int> cen_list;              // owned centers by element: [Mo-like | S-like | overflow], ascending index
  DevBuf<int> cen_key;               // scan input: 1 per Mo-like center, 2^30 per S-like center
  DevBuf<int64_t> cen_scan;          // [inum+1] exclusive scan of cen_key: list position of every index threshold
  // resident loop with halo overlap: the LJ pair rows sit in [interior | boundary] order per element (interior = both
  // centers far from the faces of the sub-domain: no ghost among the candidates); ljp_split[0..2] = packed list
  // positions {0, end of the interior part, end} plays cen_scan's role for the "index ranges" [0,1) and [1,2)
  SplitGeom split;
  bool split_valid = false;          // the pair rows on the device are in split order
  DevBuf<int> ljp_tmp;
  DevBuf<int64_t> ljp_scan;
  int64_t *ljp_split = nullptr;
  cudaStream_t halo_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_fwd = nullptr, ev_reb = nullptr, ev_rev = nullptr;
  int overlap_halo = 1;              // option "overlap_halo": halos on their own stream beside the interior kernels
  int split_elems = 2;               // option "split_elems": 2 = only the S LJ rows are split into interior / boundary launches, 3 = both elements
  int flat_halo = 1;                 // option "flat_halo": one rank, the self halos as one gather / one fold
  int peer_vote = 1;                 // option "peer_vote": reneighbor vote through peer memory + a mapped host word
  int neigh_unroll = 1;              // option "neigh_unroll" (tuning)
  int fuse_integrate = 1;            // option "fuse_integrate": final_integrate(k) rides in initial_integrate(k+1)
  DevBuf<double> nM, nS;             // parity API (b200md_rebomos_neigh)
  DevBuf<double> det_fb;             // deterministic mode: per-bond and per-center forces
  DevBuf<int> det_j;

  // ---- AEAM
  bool aeam_ready = false;
  int fp_gated = 0;                  // option "fp_gated" (two-phase API)
  DevBuf<double> fp_tmp;
  bool aeam_h2d_ready = false;       // plugin mode: upload pieces and their dependences are set up for the current master list
  long long aeam_h2d_list = -1;      // ... n_list_upload they belong to
  int aeam_maxtag = 0;
  int aeam_range_lo = 0, aeam_range_hi = 0;    // plugin mode: the range of centers the next pair-kernel launch takes
  AeamDev ap;
  DevBuf<double> spl_frho, spl_rhor, spl_z2r;    // {c3,c4,c5,c6} per row
  DevBuf<double> spl_pair;                       // fused rows {rhor c3..c6 | z2r c3..c6} per ordered type pair
  std::vector<int> frho_rows, rhor_rows, z2r_rows;
  DevBuf<double> rho, fp;
  DevBuf<int64_t> ea_off;
  DevBuf<int> ea_num, ea_val;        // inner AEAM rows (filtered to cut+margin)
  DevBuf<int> ang_list;              // owned angular atoms
  DevBuf<int> ang_key;               // scan input / output of the ordered angular-center list
  DevBuf<int64_t> ang_scan;
  DevBuf<long long> det_ffix;        // deterministic mode: fixed-point forces of the angular triplets [nall*3]
  DevBuf<double> det_part;           // deterministic mode: per-block partial sums of the global accumulators
  DevBuf<int64_t> ec_off;            // cluster form: [ncl+1] 8-aligned offsets of the union rows
  DevBuf<int> ec_num, ec_val, ec_cap;    // union row of cluster q = centers 4q..4q+3
  DevBuf<double> ec_df;              // [4 * entries] f'_{ti,tj}(r) of (entry, center), written by the density pass
  int n_ang = 0;

  // ---- device-built list scratch (neigh.cu)
  DevBuf<int> bin_of, bin_count, bin_start, bin_atoms, stencil_d;
  DevBuf<int> scan_tmp;
  DevBuf<int64_t> scan_tmp64;

  SystemState *sys = nullptr;
  NeighScratch *neigh_scratch = nullptr;
  AeamHost *aeam_host = nullptr;

  int fail(const std::string &m)
  {
    err = m;
    return -1;
  }
};

// launch bookkeeping + optional per-kernel timing (events on the launching stream)
struct LaunchScope {
  b200md_ctx *c;
  cudaEvent_t eb = nullptr;
  static cudaEvent_t get_event(b200md_ctx *c)
  {
    if (!c->event_pool.empty()) {
      cudaEvent_t e = c->event_pool.back();
      c->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  LaunchScope(b200md_ctx *ctx, const char *name) : c(ctx)
  {
    c->n_launch++;
    if (c->sync_timing) {
      auto it = c->kname_id.find(name);
      int id;
      if (it == c->kname_id.end()) {
        id = (int) c->kname.size();
        c->kname.push_back(name);
        c->kname_id[name] = id;
      } else
        id = it->second;
      TimedLaunch t;
      t.name_id = id;
      t.a = get_event(c);
      t.b = get_event(c);
      cudaEventRecord(t.a, c->stream);
      eb = t.b;
      c->pending.push_back(t);
    }
  }
  ~LaunchScope()
  {
    if (eb) cudaEventRecord(eb, c->stream);
  }
};

int b200md_collect_timers(b200md_ctx *c);    // after a stream sync: fills last_ms
// D2H of forces/energy/virial/flags + host-side accumulate (shared by the compute entry points)
int b200md_finish_compute(b200md_ctx *c, int eflag, int vflag, double *f, double *eng_vdwl, double *virial,
                          int *flags_out);

// ranged D2H of finished forces on the copy stream (f_overwrite mode): begin, one call per finished range (each
// waits for what the compute stream has queued so far), finish = scalars + flags + join of both streams
int b200md_d2h_begin(b200md_ctx *c);
int b200md_d2h_range(b200md_ctx *c, int slot, double *f_host, size_t lo, size_t hi);
int b200md_d2h_finish(b200md_ctx *c, int eflag, int vflag, double *f_host, double *eng_vdwl, double *virial,
                      int *flags_out);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over an aligned group of W lanes.  The mask names ONLY the group: groups of one warp may sit in
// different iterations of a grid-stride loop (or have left it), and a full-warp mask would then wait for
// lanes that never arrive
template <int W> __device__ __forceinline__ double group_sum(double v)
{
  const unsigned lane = threadIdx.x & 31u;
  const unsigned mask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << (lane & ~(unsigned) (W - 1)));
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
// Global accumulators (energy, virial, kinetic energy).  Default: one FP64 atomic per block and value -- the sum then
// depends on the order in which blocks retire.  Deterministic mode: the kernels are handed a TAGGED pointer (bit 0 set)
// into a table of per-block partial sums, partial[blockIdx.x][16]; kernels of one stream run one after the other, so a
// block's slot is updated by one thread at a time, and b200md_det_fold() adds the table up in a fixed order (a fixed
// two-level tree) into scal[] before it is used: bitwise reproducible whatever the order in which blocks retire.
#define B200MD_DET_BLOCKS (1 << 18)    // rows of the table: enough for 8 M atoms per GPU at 8 lanes per atom
__device__ __forceinline__ void accumulate_global(double *out, int k, double s)
{
  const unsigned long long u = (unsigned long long) out;
  if (u & 1ull) {
    double *part = (double *) (u & ~7ull);
    part[(size_t) blockIdx.x * 16 + k] += s;
  } else
    atomicAdd(&out[k], s);
}
// block-level sum of `n` doubles per thread into global accumulators (one atomic per block per value)
template <int N, int BLOCK> __device__ __forceinline__ void block_accumulate(double (&v)[N], double *out)
{
  __shared__ double sh[N][BLOCK / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = warp_sum(v[k]);
    if (lane == 0) sh[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) s += sh[threadIdx.x][w];
    accumulate_global(out, threadIdx.x, s);
  }
}
// One 32-byte sector with ONE instruction (LDG.E.ENL2.256, sm_100+): a double4 gather written as `p[j]`
// compiles to two LDG.E.128 because double4 is only 16-byte aligned as a type; every element of the arrays
// gathered here sits on a 32-byte boundary.  Halves the LSU wavefronts of the neighbor gathers.
__device__ __forceinline__ double4 ld_sector(const double4 *p)
{
  double4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
// 1/sqrt(a) for a normal, positive a in a harmless range (bond lengths, 1 + S + P): MUFU.RSQ64H seed + two Newton
// steps, no slow path, <= 2 ulp; r = a * rsqrt(a).  The library sqrt and division each cost ~25 instructions with
// their special-case handling (6 % of the bond-order kernel's instructions went there, ncu source view).
__device__ __forceinline__ double rsqrt_nr(double a)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  return fma(y, e, y);
}
__device__ __forceinline__ int4 ld_stream_int4(const int4 *p)
{
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream_int(const int *p)
{
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// L2 evict-first policy for data that is streamed once per kernel (neighbor rows, per-entry scratch): the gathered
// data -- positions and spline tables, ~80 MB -- should own the 126 MB L2, not the gigabytes that pass through
__device__ __forceinline__ unsigned long long policy_evict_first()
{
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int ld_stream_int_ef(const int *p, unsigned long long pol)
{
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ double ld_stream_f64_ef(const double *p, unsigned long long pol)
{
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_stream_f64_ef(double *p, double v, unsigned long long pol)
{
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
#endif

// shared internal entry points
// accumulator pointer handed to the kernels: scal[], or the tagged fixed-point accumulators in deterministic mode
static inline double *b200md_scal_arg(b200md_ctx *c)
{
  return (c->deterministic && c->det_part.p) ? (double *) (((unsigned long long) c->det_part.p) | 1ull) : c->scal.p;
}
int b200md_det_fold(b200md_ctx *c);    // deterministic mode: per-block partial sums -> scal[0..15] in a fixed order (and zeroed)
int b200md_peratom_begin(b200md_ctx *c, bool wanted);                      // zeroed device arrays, sets pa_e / pa_v
int b200md_peratom_finish(b200md_ctx *c, double *eatom, double *vatom);    // D2H + accumulate into the host arrays
int b200md_upload_atoms(b200md_ctx *c, int nlocal, int nghost, const double *x, const int *type,
                        const int *tag);
int b200md_exclusive_scan_i64(b200md_ctx *c, const int *in, int64_t *out, int n, int align);
