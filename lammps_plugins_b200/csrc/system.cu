// b200md -- GPU-resident MD system: the run loop LAMMPS wraps around Pair::compute(), kept entirely
// on the device for the benchmark driver (one instance per GPU / rank).
//
// Restates LAMMPS-core (stable_2Aug2023) semantics, SURVEY.md A.3-A.5:
//   Verlet::setup / Verlet::run          setup() / b200md_system_run()
//   FixNVE::initial/final_integrate      k_initial_integrate / k_final_integrate
//   Neighbor::decide / check_distance    displacement flag fused into initial integrate
//   Domain::x2lamda / lamda2x / pbc      k_x2lamda / k_lamda2x / k_pbc
//   Atom::sort                           sort_atoms()
//   CommBrick::setup/borders/forward_comm/reverse_comm  (+ Pair fp forward)   halo_*()
//   CommBrick::exchange                  migrate()   (multi-GPU only)
//   compute temp / pressure / thermo     thermo()
// Halo swaps keep LAMMPS' staged x,y,z order and send-list order, so ghost ordering is identical to the
// host engine; a swap whose partner is this rank is a device gather, otherwise an NCCL send/recv pair
// over NVLink.  Arithmetic that decides membership (slab tests, lamda conversion) is written with
// explicit non-contracted FP64 ops so the ghost set and its coordinates match the host bit for bit.

#include "common.cuh"

#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <cmath>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <unordered_set>

#define BLOCK 256
#define BIGSLAB 1.0e20
#define LOCAL_TIMEOUT_S 120
#define P2P_SLOTS 18    // 3 message kinds x 3 dimensions x 2 directions
#define VOTE_MAXR 64    // ranks the peer-memory reneighbor vote is laid out for (flag words after the halo slots)

int b200md_rebomos_build_inner(b200md_ctx *c);
int b200md_rebomos_derive_tight(b200md_ctx *c);
int b200md_rebomos_forces(b200md_ctx *c, int eflag, int vflag);
int b200md_rebomos_forces_part(b200md_ctx *c, int part, int which);
int b200md_aeam_build_inner(b200md_ctx *c);
int b200md_aeam_density(b200md_ctx *c);
int b200md_aeam_forces(b200md_ctx *c, int eflag, int vflag, int part);
int b200md_neigh_build_device(b200md_ctx *c, const b200md_box &box, int ntypes, const double *cutneighsq_h,
                              const double *cutneighghostsq_h, int nlocal, int nghost, const double4 *xt,
                              int ghost_rows, double skin, bool one_pass);
void b200md_neigh_forget(b200md_ctx *c);
void b200md_aeam_forget(b200md_ctx *c);

// In-process rank group ("loopback" transport): every rank is a b200md_ctx driven by its own host thread
// in ONE process, on the same or on different GPUs; a message is a device-to-device copy from the sender's
// buffer.  It exists so that the multi-rank code (exchange/borders/forward/reverse with real partners) can
// be checked on a box with a single GPU, where NCCL refuses two ranks on one device.
struct LocalSlot {
  const double *ptr = nullptr;
  size_t n = 0;
  long long posted = 0, consumed = 0;
};
struct LocalGroup {
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<LocalSlot> slot;    // [src * nranks + dst]
  // host-side reductions
  std::vector<double> red;
  int red_arrived = 0;
  long long red_gen = 0;
  std::vector<double> red_out;
};
static std::mutex g_groups_mu;
static std::map<int, std::shared_ptr<LocalGroup>> g_groups;
static int g_next_group = 1;

struct Swap {
  int dim = 0, sendproc = 0, recvproc = 0;
  double lo = 0, hi = 0;
  int pbc_flag = 0, pbc[6] = {0, 0, 0, 0, 0, 0};
  double bshift[3] = {0, 0, 0};    // shift applied in borders() (lamda units if triclinic)
  double fshift[3] = {0, 0, 0};    // shift applied in forward_comm() (box units)
  int nsend = 0, nrecv = 0, firstrecv = 0;
  DevBuf<int> sendlist;
};

struct SystemState {
  b200md_system_desc d;
  std::vector<double> mass;
  // geometry
  int triclinic = 0;
  double boxlo[3], boxhi[3], prd[3], h[6], h_inv[6];
  double sublo[3], subhi[3];    // box coords, or lamda if triclinic
  double cutghost[3];
  double cutneighmax = 0.0;
  std::vector<double> cutneighsq, cutneighghostsq;
  int ghost_rows = 0;
  int myloc[3], procneigh[3][2];
  int nranks = 1, me = 0;
  int maxneed[3];
  std::vector<Swap> swaps;
  // atoms (positions/types/tags/forces live in the ctx buffers)
  int nlocal = 0, nghost = 0;
  long long natoms = 0;
  DevBuf<double> v;
  DevBuf<double4> xhold, xt;
  DevBuf<double> dmass;          // per type
  DevBuf<int> itmp, itmp2;
  DevBuf<double4> x4tmp;
  DevBuf<double> dtmp;
  DevBuf<int64_t> scan64;
  DevBuf<double> sendbuf, recvbuf;
  // run state
  long long step = 0, nbuild = 0, ndanger = 0, nextsort = 0;
  int ago = 0;
  std::vector<std::vector<double>> rows;    // thermo rows
  long long nmigrated = 0, ninner = 0, noverlap = 0;
  bool fold_atomic = false;    // reverse-halo folds use atomics (they run beside force kernels on another stream)
  // fix nvt (Nose-Hoover chain, LAMMPS defaults: tchain 3, tloop 1, no drag) -- FixNH restated for the resident loop
  struct NoseHoover {
    bool on = false;
    double t_start = 0, t_stop = 0, t_period = 0;
    double eta[3] = {0, 0, 0}, eta_dot[4] = {0, 0, 0, 0}, eta_dotdot[3] = {0, 0, 0}, eta_mass[3] = {0, 0, 0};
    double t_current = 0, t_target = 0, ke_target = 0, tdof = 0, t_freq = 0;
    long long begin = 0, end = 0;
  } nh;
  // transport between ranks: NCCL (one process per GPU) or the in-process loopback group
  ncclComm_t nccl = nullptr;
  std::shared_ptr<LocalGroup> local;
  DevBuf<double> xbuf;    // exchange() staging
  // peer-memory halo (one process per GPU, CUDA IPC over NVLink): windows the neighbors write into directly
  struct PeerHalo {
    bool ok = false;
    int want = 1;                  // option "p2p_halo"
    double *win = nullptr;         // [P2P_SLOTS * 2 parities * slot_cap]
    int *flag = nullptr;           // [P2P_SLOTS * 2] epoch of the last completed write, then [2 parities][VOTE_MAXR] votes
    DevBuf<int *> vote_peers;      // every rank's vote area (own included), index = rank
    int *done = nullptr;           // block-completion counter of the push kernels
    size_t slot_cap = 0;           // doubles per slot
    std::vector<double *> pwin;    // peers' windows mapped here (index = rank)
    std::vector<int *> pflag;
    int epoch[3] = {0, 0, 0};      // forward x, forward rho/fp, reverse f
  } p2p;
  // one rank: every ghost is a periodic image of an owned atom.  flat_src[g] = that atom, flat_code[g] = the swap (1-based,
  // 8 bits per dimension) whose shift the image received in each dimension: the staged x, y, z copies of a halo (6 swaps,
  // 3 dependent launches) become ONE gather that applies the same additions in the same order
  DevBuf<int> flat_src, flat_code;
  bool flat_ok = false;
  int *vote_host = nullptr;        // mapped pinned word the vote kernel reports to: (epoch << 3) | level
  int vote_epoch = 0;
};

// ================================================================== kernels
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }

struct Geom {
  double boxlo[3], h[6], h_inv[6];
};

// Domain::x2lamda / lamda2x with the reference operation order (no FMA contraction)
__global__ void __launch_bounds__(BLOCK) k_x2lamda(const __grid_constant__ Geom g, double4 *__restrict__ x, int n)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  double4 p = x[i];
  const double d0 = p.x - g.boxlo[0], d1 = p.y - g.boxlo[1], d2 = p.z - g.boxlo[2];
  p.x = add(add(mul(g.h_inv[0], d0), mul(g.h_inv[5], d1)), mul(g.h_inv[4], d2));
  p.y = add(mul(g.h_inv[1], d1), mul(g.h_inv[3], d2));
  p.z = mul(g.h_inv[2], d2);
  x[i] = p;
}
__global__ void __launch_bounds__(BLOCK) k_lamda2x(const __grid_constant__ Geom g, double4 *__restrict__ x, int n)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  double4 p = x[i];
  const double l0 = p.x, l1 = p.y, l2 = p.z;
  p.x = add(add(add(mul(g.h[0], l0), mul(g.h[5], l1)), mul(g.h[4], l2)), g.boxlo[0]);
  p.y = add(add(mul(g.h[1], l1), mul(g.h[3], l2)), g.boxlo[1]);
  p.z = add(mul(g.h[2], l2), g.boxlo[2]);
  x[i] = p;
}

// Domain::pbc on owned atoms
__global__ void __launch_bounds__(BLOCK) k_pbc(double4 *__restrict__ x, int n, double lo0, double lo1, double lo2,
                                               double hi0, double hi1, double hi2, double p0, double p1, double p2)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  double4 p = x[i];
  if (p.x < lo0) p.x += p0;
  if (p.x >= hi0) { p.x -= p0; p.x = fmax(p.x, lo0); }
  if (p.y < lo1) p.y += p1;
  if (p.y >= hi1) { p.y -= p1; p.y = fmax(p.y, lo1); }
  if (p.z < lo2) p.z += p2;
  if (p.z >= hi2) { p.z -= p2; p.z = fmax(p.z, lo2); }
  x[i] = p;
}

__device__ __forceinline__ double comp(const double4 &p, int dim) { return dim == 0 ? p.x : dim == 1 ? p.y : p.z; }

// borders(): flag atoms of [nfirst,nlast) inside the slab, using >= lo && <= hi
__global__ void __launch_bounds__(BLOCK) k_slab_flags(const double4 *__restrict__ x, int nfirst, int nlast, int dim,
                                                      double lo, double hi, int *__restrict__ flag)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= nlast - nfirst) return;
  const double v = comp(x[nfirst + k], dim);
  flag[k] = (v >= lo && v <= hi) ? 1 : 0;
}
__global__ void __launch_bounds__(BLOCK) k_scatter_list(const int *__restrict__ flag, const int64_t *__restrict__ pos,
                                                        int nfirst, int n, int *__restrict__ list)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  if (flag[k]) list[pos[k]] = nfirst + k;
}
// AtomVec::pack_border + unpack_border for a self swap: ghost = copy of list atom shifted by (dx,dy,dz)
__global__ void __launch_bounds__(BLOCK) k_border_copy(double4 *__restrict__ x, int *__restrict__ type,
                                                       int *__restrict__ tag, const int *__restrict__ list, int n,
                                                       int first, double dx, double dy, double dz, int pbc_flag)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int j = list[k];
  double4 p = x[j];
  if (pbc_flag) {
    p.x = add(p.x, dx);
    p.y = add(p.y, dy);
    p.z = add(p.z, dz);
  }
  x[first + k] = p;
  type[first + k] = type[j];
  tag[first + k] = tag[j];
}
// multi-GPU variants: pack to / unpack from a contiguous buffer of 6 doubles per atom (x,y,z,w,type,tag)
__global__ void __launch_bounds__(BLOCK) k_border_pack(const double4 *__restrict__ x, const int *__restrict__ type,
                                                       const int *__restrict__ tag, const int *__restrict__ list,
                                                       int n, double dx, double dy, double dz, int pbc_flag,
                                                       double *__restrict__ buf)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int j = list[k];
  double4 p = x[j];
  if (pbc_flag) {
    p.x = add(p.x, dx);
    p.y = add(p.y, dy);
    p.z = add(p.z, dz);
  }
  buf[6 * (size_t) k] = p.x;
  buf[6 * (size_t) k + 1] = p.y;
  buf[6 * (size_t) k + 2] = p.z;
  buf[6 * (size_t) k + 3] = p.w;
  buf[6 * (size_t) k + 4] = (double) type[j];
  buf[6 * (size_t) k + 5] = (double) tag[j];
}
__global__ void __launch_bounds__(BLOCK) k_border_unpack(double4 *__restrict__ x, int *__restrict__ type,
                                                         int *__restrict__ tag, int first, int n,
                                                         const double *__restrict__ buf)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  x[first + k] = make_double4(buf[6 * (size_t) k], buf[6 * (size_t) k + 1], buf[6 * (size_t) k + 2], buf[6 * (size_t) k + 3]);
  type[first + k] = (int) buf[6 * (size_t) k + 4];
  tag[first + k] = (int) buf[6 * (size_t) k + 5];
}

// CommBrick::exchange ------------------------------------------------------------------------------
// leavers of one dimension: flag owned atoms with x[dim] < lo || x[dim] >= hi
__global__ void __launch_bounds__(BLOCK) k_leave_flags(const double4 *__restrict__ x, int n, int dim, double lo,
                                                       double hi, int *__restrict__ flag)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const double v = comp(x[i], dim);
  flag[i] = (v < lo || v >= hi) ? 1 : 0;
}
// AtomVecAtomic::pack_exchange in the emission order computed on the host: 8 doubles per atom
// (x, y, z, vx, vy, vz, tag, type)
__global__ void __launch_bounds__(BLOCK) k_exchange_pack(const double4 *__restrict__ x, const double *__restrict__ v,
                                                         const int *__restrict__ type, const int *__restrict__ tag,
                                                         const int *__restrict__ order, int n, double *__restrict__ buf)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int i = order[k];
  const double4 p = x[i];
  double *b = buf + 8 * (size_t) k;
  b[0] = p.x;
  b[1] = p.y;
  b[2] = p.z;
  b[3] = v[3 * (size_t) i];
  b[4] = v[3 * (size_t) i + 1];
  b[5] = v[3 * (size_t) i + 2];
  b[6] = (double) tag[i];
  b[7] = (double) type[i];
}
// AtomVec::copy(src -> dst) for the (hole, tail stayer) pairs; sources and destinations are disjoint
__global__ void __launch_bounds__(BLOCK) k_exchange_fill(double4 *__restrict__ x, double *__restrict__ v,
                                                         int *__restrict__ type, int *__restrict__ tag,
                                                         const int *__restrict__ moves, int n)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int dst = moves[2 * k], src = moves[2 * k + 1];
  x[dst] = x[src];
  v[3 * (size_t) dst] = v[3 * (size_t) src];
  v[3 * (size_t) dst + 1] = v[3 * (size_t) src + 1];
  v[3 * (size_t) dst + 2] = v[3 * (size_t) src + 2];
  type[dst] = type[src];
  tag[dst] = tag[src];
}
// incoming atoms: keep those with lo <= x[dim] < hi, appended in buffer order (one warp, ballot compaction)
__global__ void __launch_bounds__(32) k_exchange_unpack(const double *__restrict__ buf, int nrecv, int dim, double lo,
                                                        double hi, double4 *__restrict__ x, double *__restrict__ v,
                                                        int *__restrict__ type, int *__restrict__ tag, int nlocal,
                                                        int *__restrict__ nlocal_out)
{
  const int lane = threadIdx.x;
  const unsigned lt = (1u << lane) - 1u;
  int n = nlocal;
  for (int k0 = 0; k0 < nrecv; k0 += 32) {
    const int k = k0 + lane;
    bool keep = false;
    const double *b = buf + 8 * (size_t) k;
    if (k < nrecv) {
      const double val = b[dim];
      keep = val >= lo && val < hi;
    }
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int j = n + __popc(mk & lt);
      x[j] = make_double4(b[0], b[1], b[2], 0.0);
      v[3 * (size_t) j] = b[3];
      v[3 * (size_t) j + 1] = b[4];
      v[3 * (size_t) j + 2] = b[5];
      tag[j] = (int) b[6];
      type[j] = (int) b[7];
    }
    n += __popc(mk);
  }
  if (lane == 0) *nlocal_out = n;
}

// forward_comm: positions of ghosts from their source atoms (self swap)
__global__ void __launch_bounds__(BLOCK) k_forward_x(double4 *__restrict__ x, const int *__restrict__ list, int n,
                                                     int first, double dx, double dy, double dz, int pbc_flag)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const double4 p = x[list[k]];
  double4 q = x[first + k];
  if (pbc_flag) {
    q.x = p.x + dx;
    q.y = p.y + dy;
    q.z = p.z + dz;
  } else {
    q.x = p.x;
    q.y = p.y;
    q.z = p.z;
  }
  x[first + k] = q;
}
__global__ void __launch_bounds__(BLOCK) k_forward_x_pack(const double4 *__restrict__ x, const int *__restrict__ list,
                                                          int n, double dx, double dy, double dz, int pbc_flag,
                                                          double *__restrict__ buf)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const double4 p = x[list[k]];
  buf[3 * (size_t) k] = pbc_flag ? p.x + dx : p.x;
  buf[3 * (size_t) k + 1] = pbc_flag ? p.y + dy : p.y;
  buf[3 * (size_t) k + 2] = pbc_flag ? p.z + dz : p.z;
}
__global__ void __launch_bounds__(BLOCK) k_forward_x_unpack(double4 *__restrict__ x, int first, int n,
                                                            const double *__restrict__ buf)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  double4 q = x[first + k];
  q.x = buf[3 * (size_t) k];
  q.y = buf[3 * (size_t) k + 1];
  q.z = buf[3 * (size_t) k + 2];
  x[first + k] = q;
}
// per-atom scalar pair (rho, fp) forward: PairAEAM::pack/unpack_forward_comm (pair_aeam.cpp:946-965)
__global__ void __launch_bounds__(BLOCK) k_forward_s2(double *__restrict__ a, double *__restrict__ b,
                                                      const int *__restrict__ list, int n, int first)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int j = list[k];
  a[first + k] = a[j];
  b[first + k] = b[j];
}
__global__ void __launch_bounds__(BLOCK) k_forward_s2_pack(const double *__restrict__ a, const double *__restrict__ b,
                                                           const int *__restrict__ list, int n, double *__restrict__ buf)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int j = list[k];
  buf[2 * (size_t) k] = a[j];
  buf[2 * (size_t) k + 1] = b[j];
}
__global__ void __launch_bounds__(BLOCK) k_forward_s2_unpack(double *__restrict__ a, double *__restrict__ b, int first,
                                                             int n, const double *__restrict__ buf)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  a[first + k] = buf[2 * (size_t) k];
  b[first + k] = buf[2 * (size_t) k + 1];
}
// reverse_comm: ghost forces summed into their source atoms; within one swap every source is unique
// ATOMIC: the fold runs on the halo stream while force kernels on the compute stream still add to the same owned atoms
__device__ __forceinline__ void fold3(double *f, size_t j, double a, double b, double c, bool atomic)
{
  if (atomic) {
    atomicAdd(&f[3 * j], a);
    atomicAdd(&f[3 * j + 1], b);
    atomicAdd(&f[3 * j + 2], c);
  } else {
    f[3 * j] += a;
    f[3 * j + 1] += b;
    f[3 * j + 2] += c;
  }
}
__global__ void __launch_bounds__(BLOCK) k_reverse_f(double *__restrict__ f, const int *__restrict__ list, int n,
                                                     int first, bool atomic)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const size_t j = list[k], g = (size_t) first + k;
  fold3(f, j, f[3 * g], f[3 * g + 1], f[3 * g + 2], atomic);
}
__global__ void __launch_bounds__(BLOCK) k_reverse_f_unpack(double *__restrict__ f, const int *__restrict__ list, int n,
                                                            const double *__restrict__ buf, bool atomic)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  fold3(f, list[k], buf[3 * (size_t) k], buf[3 * (size_t) k + 1], buf[3 * (size_t) k + 2], atomic);
}

// the -d and +d self swaps of one dimension in ONE launch (single rank: 3 launches per halo instead of 6; each of these
// kernels is a few microseconds of work behind ~5 microseconds of launch latency).  Both swaps scan the same, earlier
// atoms and own disjoint ghost ranges, so the forward copy needs no ordering; in the reverse fold an atom that sits in
// both send lists (a box thinner than twice the ghost cutoff) receives two adds, hence the atomics.
__global__ void __launch_bounds__(BLOCK) k_forward_x2(double4 *__restrict__ x, const int *__restrict__ list_a, int na,
                                                      int first_a, double ax, double ay, double az, int pbc_a,
                                                      const int *__restrict__ list_b, int nb, int first_b, double bx,
                                                      double by, double bz, int pbc_b)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= na + nb) return;
  const bool second = k >= na;
  if (second) k -= na;
  const int src = second ? list_b[k] : list_a[k];
  const int dst = (second ? first_b : first_a) + k;
  const double4 p = x[src];
  double4 q = x[dst];
  if (second ? pbc_b : pbc_a) {
    q.x = p.x + (second ? bx : ax);
    q.y = p.y + (second ? by : ay);
    q.z = p.z + (second ? bz : az);
  } else {
    q.x = p.x;
    q.y = p.y;
    q.z = p.z;
  }
  x[dst] = q;
}
__global__ void __launch_bounds__(BLOCK) k_reverse_f2(double *__restrict__ f, const int *__restrict__ list_a, int na,
                                                      int first_a, const int *__restrict__ list_b, int nb, int first_b)
{
  int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= na + nb) return;
  const bool second = k >= na;
  if (second) k -= na;
  const size_t j = second ? list_b[k] : list_a[k], g = (size_t) (second ? first_b : first_a) + k;
  atomicAdd(&f[3 * j], f[3 * g]);
  atomicAdd(&f[3 * j + 1], f[3 * g + 1]);
  atomicAdd(&f[3 * j + 2], f[3 * g + 2]);
}

// ------------------------------------------------------------------ one rank: flat self halo
struct FlatShifts {
  double s[6][3];
  int pbc[6];
};
__global__ void __launch_bounds__(BLOCK) k_flat_build(const int *__restrict__ list, int n, int first, int nlocal, int swap,
                                                      int dim, int *__restrict__ src, int *__restrict__ code)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  const int j = list[k], g = first + k - nlocal;
  int sj = j, cj = 0;
  if (j >= nlocal) {    // the image of an image (a ghost of an earlier dimension)
    sj = src[j - nlocal];
    cj = code[j - nlocal];
  }
  src[g] = sj;
  code[g] = cj | ((swap + 1) << (8 * dim));
}
__global__ void __launch_bounds__(BLOCK) k_flat_forward_x(double4 *__restrict__ x, const int *__restrict__ src,
                                                          const int *__restrict__ code, int ng, int nlocal,
                                                          const __grid_constant__ FlatShifts sh)
{
  const int g = blockIdx.x * BLOCK + threadIdx.x;
  if (g >= ng) return;
  double4 p = x[src[g]];
  const int cd = code[g];
#pragma unroll
  for (int d = 0; d < 3; d++) {    // the staged halo's additions, dimension by dimension
    const int sid = ((cd >> (8 * d)) & 255) - 1;
    if (sid >= 0 && sh.pbc[sid]) {
      p.x = p.x + sh.s[sid][0];
      p.y = p.y + sh.s[sid][1];
      p.z = p.z + sh.s[sid][2];
    }
  }
  double4 q = x[nlocal + g];
  q.x = p.x;
  q.y = p.y;
  q.z = p.z;
  x[nlocal + g] = q;
}
__global__ void __launch_bounds__(BLOCK) k_flat_forward_s2(double *__restrict__ a, double *__restrict__ b,
                                                           const int *__restrict__ src, int ng, int nlocal)
{
  const int g = blockIdx.x * BLOCK + threadIdx.x;
  if (g >= ng) return;
  const int j = src[g];
  a[nlocal + g] = a[j];
  b[nlocal + g] = b[j];
}
__global__ void __launch_bounds__(BLOCK) k_flat_reverse_f(double *__restrict__ f, const int *__restrict__ src, int ng,
                                                          int nlocal)
{
  const int g = blockIdx.x * BLOCK + threadIdx.x;
  if (g >= ng) return;
  const size_t j = (size_t) src[g], q = (size_t) nlocal + g;
  atomicAdd(&f[3 * j], f[3 * q]);
  atomicAdd(&f[3 * j + 1], f[3 * q + 1]);
  atomicAdd(&f[3 * j + 2], f[3 * q + 2]);
}

// ------------------------------------------------------------------ peer-memory halo kernels
// The sender packs straight into the RECEIVER's window over NVLink (no staging copy, no NCCL launch), then the last
// block to finish publishes the epoch in the receiver's flag; the receiver's unpack kernel waits for the epoch.
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
struct PushDesc {
  const int *list[2];    // send lists of the -d and +d swap (forward) -- unused for reverse
  int n[2];              // items
  int first[2];          // reverse: first ghost of the swap
  double *dst[2];        // slot in the peer's window
  int *flag[2];          // the peer's flag for that slot
  double shift[2][3];
  int pbc[2];
};
__device__ __forceinline__ void push_finish(const PushDesc &d, int epoch, int *done)
{
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(done, 1) == (int) gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      st_release_sys(d.flag[0], epoch);
      st_release_sys(d.flag[1], epoch);
    }
  }
}
// Push kernels: a capped grid of blocks walks the items in tiles of BLOCK.  A tile is gathered and packed into shared
// memory, then written to the peer window as 16-byte stores that a warp coalesces into 512 contiguous bytes -- element-wise
// 8-byte remote stores with a 24-byte stride reached 20-90 GB/s over NVLink (r01: 137 GB/s at best).  One system-scope
// fence per block (not per tile), then the last block publishes the epoch.
__device__ __forceinline__ void st_tile(double *dst, const double *tile, int ndbl)
{
  // dst is 16-byte aligned (slot bases are, and a tile starts at a multiple of BLOCK items)
  const int npair = ndbl >> 1;
  for (int t = threadIdx.x; t < npair; t += BLOCK) {
    const double2 v = make_double2(tile[2 * t], tile[2 * t + 1]);
    *reinterpret_cast<double2 *>(dst + 2 * t) = v;
  }
  if ((ndbl & 1) && threadIdx.x == 0) dst[ndbl - 1] = tile[ndbl - 1];
}
__global__ void __launch_bounds__(BLOCK) k_p2p_push_x(const double4 *__restrict__ x, const __grid_constant__ PushDesc d,
                                                      int epoch, int *done)
{
  __shared__ double tile[3 * BLOCK];
  for (int w = 0; w < 2; w++) {
    const int n = d.n[w];
    for (int base = blockIdx.x * BLOCK; base < n; base += gridDim.x * BLOCK) {
      const int q = base + threadIdx.x;
      if (q < n) {
        const double4 p = x[d.list[w][q]];
        tile[3 * threadIdx.x] = d.pbc[w] ? p.x + d.shift[w][0] : p.x;
        tile[3 * threadIdx.x + 1] = d.pbc[w] ? p.y + d.shift[w][1] : p.y;
        tile[3 * threadIdx.x + 2] = d.pbc[w] ? p.z + d.shift[w][2] : p.z;
      }
      __syncthreads();
      st_tile(d.dst[w] + 3 * (size_t) base, tile, 3 * min(BLOCK, n - base));
      __syncthreads();
    }
  }
  push_finish(d, epoch, done);
}
__global__ void __launch_bounds__(BLOCK) k_p2p_push_s2(const double *__restrict__ a, const double *__restrict__ b,
                                                       const __grid_constant__ PushDesc d, int epoch, int *done)
{
  for (int w = 0; w < 2; w++) {
    const int n = d.n[w];
    for (int q = blockIdx.x * BLOCK + threadIdx.x; q < n; q += gridDim.x * BLOCK) {
      const int j = d.list[w][q];
      *reinterpret_cast<double2 *>(d.dst[w] + 2 * (size_t) q) = make_double2(a[j], b[j]);    // 16 B per lane, coalesced
    }
  }
  push_finish(d, epoch, done);
}
// reverse: the ghost forces of a swap are contiguous at f[3*first ...]
__global__ void __launch_bounds__(BLOCK) k_p2p_push_f(const double *__restrict__ f, const __grid_constant__ PushDesc d,
                                                      int epoch, int *done)
{
  for (int w = 0; w < 2; w++) {
    const int nd = 3 * d.n[w];
    const double *src = f + 3 * (size_t) d.first[w];
    const int npair = nd >> 1;
    for (int t = blockIdx.x * BLOCK + threadIdx.x; t < npair; t += gridDim.x * BLOCK)
      *reinterpret_cast<double2 *>(d.dst[w] + 2 * (size_t) t) = make_double2(src[2 * t], src[2 * t + 1]);
    if ((nd & 1) && blockIdx.x == 0 && threadIdx.x == 0) d.dst[w][nd - 1] = src[nd - 1];
  }
  push_finish(d, epoch, done);
}
__device__ __forceinline__ void wait_epoch(const int *flag0, const int *flag1, int epoch)
{
  if (!flag0) return;    // a k_p2p_wait launch ahead of this kernel has seen the epoch
  if (threadIdx.x == 0) {
    while (ld_acquire_sys(flag0) != epoch) {}
    if (flag1) while (ld_acquire_sys(flag1) != epoch) {}
  }
  __syncthreads();
}
// The wait for the neighbors' epoch in a launch of its own, ONE warp: when every block of an unpack launch spun on the
// flag, hundreds of high-priority CTAs sat on the SMs, doing nothing, beside the force kernels of the compute stream
// (N = 8: interior + boundary LJ 0.60 ms against 0.50 ms on one GPU).  The unpack launch behind it starts when the data is there.
__global__ void __launch_bounds__(32) k_p2p_wait(const int *flag0, const int *flag1, int epoch)
{
  const int *f = threadIdx.x == 0 ? flag0 : (threadIdx.x == 1 ? flag1 : nullptr);
  if (f)
    while (ld_acquire_sys(f) != epoch) {}
}
__global__ void __launch_bounds__(BLOCK) k_p2p_unpack_x(double4 *__restrict__ x, const double *src0, int first0, int n0,
                                                        const int *flag0, const double *src1, int first1, int n1,
                                                        const int *flag1, int epoch)
{
  wait_epoch(flag0, flag1, epoch);
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  const int w = k < n0 ? 0 : 1;
  const int q = w ? k - n0 : k;
  if (q >= (w ? n1 : n0)) return;
  const double *b = (w ? src1 : src0) + 3 * (size_t) q;
  const size_t g = (size_t) (w ? first1 : first0) + q;
  double4 v = x[g];
  v.x = __ldcg(b);
  v.y = __ldcg(b + 1);
  v.z = __ldcg(b + 2);
  x[g] = v;
}
__global__ void __launch_bounds__(BLOCK) k_p2p_unpack_s2(double *__restrict__ a, double *__restrict__ bb, const double *src0,
                                                         int first0, int n0, const int *flag0, const double *src1,
                                                         int first1, int n1, const int *flag1, int epoch)
{
  wait_epoch(flag0, flag1, epoch);
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  const int w = k < n0 ? 0 : 1;
  const int q = w ? k - n0 : k;
  if (q >= (w ? n1 : n0)) return;
  const double *b = (w ? src1 : src0) + 2 * (size_t) q;
  const size_t g = (size_t) (w ? first1 : first0) + q;
  a[g] = __ldcg(b);
  bb[g] = __ldcg(b + 1);
}
__global__ void __launch_bounds__(BLOCK) k_p2p_unpack_f(double *__restrict__ f, const int *__restrict__ list, int n,
                                                        const double *src, const int *flag, int epoch, bool atomic)
{
  wait_epoch(flag, nullptr, epoch);
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= n) return;
  fold3(f, list[k], __ldcg(src + 3 * (size_t) k), __ldcg(src + 3 * (size_t) k + 1), __ldcg(src + 3 * (size_t) k + 2), atomic);
}

// Reneighbor vote without NCCL and without a stream synchronisation: thread r stores this rank's level
// (flags[9], set by k_initial_integrate) into rank r's vote area over NVLink, waits for rank r's vote of the same epoch
// in its own area, and thread 0 reports the maximum to a mapped host word the run loop spins on.  The ncclAllReduce of one
// int + cudaMemcpyAsync + cudaStreamSynchronize it replaces left the GPU idle for ~30 us per step at 2 ranks (step
// trace).  Two parities: a rank can be at most one vote ahead of a peer (it needs that peer's vote to go on).  One rank:
// only the report.  Level 7 = a peer's vote did not arrive (the host fails the run).
__global__ void __launch_bounds__(VOTE_MAXR) k_vote(int *__restrict__ flag9, int *const *__restrict__ peers, int me,
                                                    int nranks, int epoch, int *host_word)
{
  __shared__ int s_lvl[VOTE_MAXR];
  const int t = threadIdx.x;
  const int level = *flag9;
  int lvl = (t == 0) ? level : 0;
  if (nranks > 1 && t < nranks) {
    const int off = (epoch & 1) * VOTE_MAXR;
    st_release_sys(peers[t] + off + me, (epoch << 3) | level);
    const int *mine = peers[me] + off + t;
    int v = 0;
    long long spins = 0;
    while (((v = ld_acquire_sys(mine)) >> 3) != epoch && ++spins < (1LL << 27)) {}
    lvl = ((v >> 3) == epoch) ? (v & 7) : 7;
  }
  s_lvl[t] = lvl;
  __syncthreads();
  if (t == 0) {
    int m = 0;
    for (int r = 0; r < (nranks > 1 ? nranks : 1); r++) m = max(m, s_lvl[r]);
    *flag9 = 0;
    *((volatile int *) host_word) = (epoch << 3) | m;
    __threadfence_system();
  }
}

// FixNVE::initial_integrate fused with Neighbor::check_distance.  flags[9] = 3: some atom moved more than
// skin/2 since the master list was built (LAMMPS' rebuild rule); 2: some atom moved more than margin/2 since
// the inner lists were derived from it (only the inner filter is redone); 1: more than margin_t/2 since the tight rows
// were derived from the inner lists (only that short pass is redone)
__global__ void __launch_bounds__(BLOCK) k_initial_integrate(double4 *__restrict__ x, double *__restrict__ v,
                                                             const double *__restrict__ f, const int *__restrict__ type,
                                                             const double *__restrict__ mass, int nlocal, double dtf,
                                                             double dtv, const double4 *__restrict__ xhold,
                                                             double triggersq, const double4 *__restrict__ xhold_inner,
                                                             double innersq, const double4 *__restrict__ xhold_tight,
                                                             double tightsq, int *__restrict__ flags, int fuse_final)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nlocal) return;
  const double dtfm = dtf / mass[type[i]];
  double4 p = x[i];
  double vx = v[3 * (size_t) i], vy = v[3 * (size_t) i + 1], vz = v[3 * (size_t) i + 2];
  const double fx = f[3 * (size_t) i], fy = f[3 * (size_t) i + 1], fz = f[3 * (size_t) i + 2];
  if (fuse_final) {    // FixNVE::final_integrate of the previous step, same operations in the same order (v, f read once)
    vx += dtfm * fx;
    vy += dtfm * fy;
    vz += dtfm * fz;
  }
  vx += dtfm * fx;
  vy += dtfm * fy;
  vz += dtfm * fz;
  p.x += dtv * vx;
  p.y += dtv * vy;
  p.z += dtv * vz;
  v[3 * (size_t) i] = vx;
  v[3 * (size_t) i + 1] = vy;
  v[3 * (size_t) i + 2] = vz;
  x[i] = p;
  const double4 h = xhold[i];
  const double dx = p.x - h.x, dy = p.y - h.y, dz = p.z - h.z;
  int level = 0;
  if (dx * dx + dy * dy + dz * dz > triggersq) level = 3;
  if (level == 0 && xhold_inner) {
    const double4 g = xhold_inner[i];
    const double ex = p.x - g.x, ey = p.y - g.y, ez = p.z - g.z;
    if (ex * ex + ey * ey + ez * ez > innersq) level = 2;
  }
  if (level == 0 && xhold_tight) {    // also when the inner lists carry the full skin and have no trigger of their own
    const double4 t = xhold_tight[i];
    const double tx = p.x - t.x, ty = p.y - t.y, tz = p.z - t.z;
    if (tx * tx + ty * ty + tz * tz > tightsq) level = 1;
  }
  if (level) atomicMax(&flags[9], level);
}
__global__ void __launch_bounds__(BLOCK) k_final_integrate(double *__restrict__ v, const double *__restrict__ f,
                                                           const int *__restrict__ type, const double *__restrict__ mass,
                                                           int nlocal, double dtf)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= nlocal) return;
  const double dtfm = dtf / mass[type[i]];
  v[3 * (size_t) i] += dtfm * f[3 * (size_t) i];
  v[3 * (size_t) i + 1] += dtfm * f[3 * (size_t) i + 1];
  v[3 * (size_t) i + 2] += dtfm * f[3 * (size_t) i + 2];
}
// compute temp: sum m v^2 into scal[8]
__global__ void __launch_bounds__(BLOCK) k_ke(const double *__restrict__ v, const int *__restrict__ type,
                                              const double *__restrict__ mass, int nlocal, double *__restrict__ scal)
{
  double t[1] = {0.0};
  for (int i = blockIdx.x * BLOCK + threadIdx.x; i < nlocal; i += gridDim.x * BLOCK) {
    const double vx = v[3 * (size_t) i], vy = v[3 * (size_t) i + 1], vz = v[3 * (size_t) i + 2];
    t[0] += (vx * vx + vy * vy + vz * vz) * mass[type[i]];
  }
  block_accumulate<1, BLOCK>(t, scal + 8);
}
// FixNH::nh_v_temp: v *= exp(-dt/2 * eta_dot[0])
__global__ void __launch_bounds__(BLOCK) k_scale_v(double *__restrict__ v, size_t n3, double factor)
{
  const size_t k = (size_t) blockIdx.x * BLOCK + threadIdx.x;
  if (k < n3) v[k] *= factor;
}
// w component: potential-specific element code from the LAMMPS type
__global__ void __launch_bounds__(BLOCK) k_set_w(double4 *__restrict__ x, const int *__restrict__ type,
                                                 const int *__restrict__ map, int use_map, int n)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const int t = type[i];
  // rebomos: element code as a double; aeam: element in the two low mantissa bits (aeam.cu, w_encode)
  x[i].w = use_map ? (double) map[t] : __longlong_as_double((long long) (t - 1));
}
__global__ void __launch_bounds__(BLOCK) k_make_xt(const double4 *__restrict__ x, const int *__restrict__ type, int n,
                                                   double4 *__restrict__ xt)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  double4 p = x[i];
  p.w = (double) type[i];
  xt[i] = p;
}
__global__ void __launch_bounds__(BLOCK) k_upload_x(const double *__restrict__ xa, int n, double4 *__restrict__ x)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  x[i] = make_double4(xa[3 * (size_t) i], xa[3 * (size_t) i + 1], xa[3 * (size_t) i + 2], 0.0);
}
__global__ void __launch_bounds__(BLOCK) k_download_x(const double4 *__restrict__ x, int n, double *__restrict__ xa)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const double4 p = x[i];
  xa[3 * (size_t) i] = p.x;
  xa[3 * (size_t) i + 1] = p.y;
  xa[3 * (size_t) i + 2] = p.z;
}

// Atom::sort: bin index of owned atoms (box coords), clamped
struct SortGeom {
  double lo[3], inv[3];
  int nb[3];
};
__global__ void __launch_bounds__(BLOCK) k_sort_bins(const __grid_constant__ SortGeom g, const double4 *__restrict__ x,
                                                     int n, int *__restrict__ bin_of, int *__restrict__ count)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const double4 p = x[i];
  int ix = (int) ((p.x - g.lo[0]) * g.inv[0]);
  int iy = (int) ((p.y - g.lo[1]) * g.inv[1]);
  int iz = (int) ((p.z - g.lo[2]) * g.inv[2]);
  ix = min(max(ix, 0), g.nb[0] - 1);
  iy = min(max(iy, 0), g.nb[1] - 1);
  iz = min(max(iz, 0), g.nb[2] - 1);
  const int b = iz * g.nb[1] * g.nb[0] + iy * g.nb[0] + ix;
  bin_of[i] = b;
  atomicAdd(&count[b], 1);
}
__global__ void __launch_bounds__(BLOCK) k_sort_fill(const int *__restrict__ bin_of, int n,
                                                     const int64_t *__restrict__ start, int *__restrict__ cursor,
                                                     int *__restrict__ perm)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const int b = bin_of[i];
  perm[start[b] + atomicAdd(&cursor[b], 1)] = i;
}
__global__ void __launch_bounds__(BLOCK) k_sort_within(const int64_t *__restrict__ start, int nbins, int *__restrict__ perm)
{
  int b = blockIdx.x * BLOCK + threadIdx.x;
  if (b >= nbins) return;
  const int s = (int) start[b], e = (int) start[b + 1];
  for (int a = s + 1; a < e; a++) {
    const int v = perm[a];
    int q = a - 1;
    while (q >= s && perm[q] > v) {
      perm[q + 1] = perm[q];
      q--;
    }
    perm[q + 1] = v;
  }
}
__global__ void __launch_bounds__(BLOCK) k_permute(const int *__restrict__ perm, int n, const double4 *__restrict__ x,
                                                   const double *__restrict__ v, const int *__restrict__ type,
                                                   const int *__restrict__ tag, double4 *__restrict__ xo,
                                                   double *__restrict__ vo, int *__restrict__ to, int *__restrict__ go)
{
  int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const int j = perm[i];
  xo[i] = x[j];
  vo[3 * (size_t) i] = v[3 * (size_t) j];
  vo[3 * (size_t) i + 1] = v[3 * (size_t) j + 1];
  vo[3 * (size_t) i + 2] = v[3 * (size_t) j + 2];
  to[i] = type[j];
  go[i] = tag[j];
}

// ================================================================== host helpers
static inline int nblk(long long n) { return (int) ((n + BLOCK - 1) / BLOCK); }

#define NCCL_TRY(ctx, call)                                                                          \
  do {                                                                                               \
    ncclResult_t r__ = (call);                                                                       \
    if (r__ != ncclSuccess) {                                                                        \
      (ctx)->fail(std::string("NCCL error: ") + ncclGetErrorString(r__) + " (" #call ")");           \
      return B200MD_ERR_NCCL;                                                                        \
    }                                                                                                \
  } while (0)

static Geom make_geom(const SystemState *s)
{
  Geom g;
  for (int d = 0; d < 3; d++) g.boxlo[d] = s->boxlo[d];
  for (int k = 0; k < 6; k++) {
    g.h[k] = s->h[k];
    g.h_inv[k] = s->h_inv[k];
  }
  return g;
}

static int ensure_atoms(b200md_ctx *c, SystemState *s, size_t n)
{
  cudaStream_t st = c->stream;
  CUDA_TRY(c, c->xq.reserve(n + 64, true, st));
  CUDA_TRY(c, c->type.reserve(n + 64, true, st));
  CUDA_TRY(c, c->tag.reserve(n + 64, true, st));
  CUDA_TRY(c, c->f.reserve(3 * n + 64, true, st));
  CUDA_TRY(c, s->v.reserve(3 * n + 64, true, st));
  return B200MD_OK;
}

// ------------------------------------------------------------------ transport
// loopback: sender publishes its device buffer, receiver copies device-to-device, sender waits for the copy
static int local_sendrecv(b200md_ctx *c, SystemState *s, const double *sbuf, size_t nsend, int sendto, double *rbuf,
                          size_t nrecv, int recvfrom)
{
  LocalGroup &G = *s->local;
  const int me = s->me, R = G.nranks;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));    // my send buffer is complete
  if (nsend) {
    std::unique_lock<std::mutex> lk(G.mu);
    LocalSlot &sl = G.slot[(size_t) me * R + sendto];
    sl.ptr = sbuf;
    sl.n = nsend;
    sl.posted++;
    G.cv.notify_all();
  }
  if (nrecv) {
    const double *src = nullptr;
    size_t n = 0;
    {
      std::unique_lock<std::mutex> lk(G.mu);
      LocalSlot &sl = G.slot[(size_t) recvfrom * R + me];
      if (!G.cv.wait_for(lk, std::chrono::seconds(LOCAL_TIMEOUT_S), [&] { return sl.posted > sl.consumed; })) {
        c->fail("loopback transport: timed out waiting for a message from rank " + std::to_string(recvfrom));
        return B200MD_ERR_NCCL;
      }
      src = sl.ptr;
      n = sl.n;
    }
    ARG_CHECK(c, n == nrecv, "loopback transport: message size mismatch between ranks");
    // device-to-device cudaMemcpy does not block the host: copy on MY stream and wait, so that the copy is
    // ordered before my next kernels and finished before the sender may reuse its buffer
    CUDA_TRY(c, cudaMemcpyAsync(rbuf, src, n * sizeof(double), cudaMemcpyDefault, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    {
      std::unique_lock<std::mutex> lk(G.mu);
      G.slot[(size_t) recvfrom * R + me].consumed++;
      G.cv.notify_all();
    }
  }
  if (nsend) {
    std::unique_lock<std::mutex> lk(G.mu);
    LocalSlot &sl = G.slot[(size_t) me * R + sendto];
    if (!G.cv.wait_for(lk, std::chrono::seconds(LOCAL_TIMEOUT_S), [&] { return sl.consumed == sl.posted; })) {
      c->fail("loopback transport: rank " + std::to_string(sendto) + " never took my message");
      return B200MD_ERR_NCCL;
    }
  }
  return B200MD_OK;
}

// op: 0 = sum, 1 = max; values are doubles on the host.  Returns false on timeout (a peer died).
static bool local_allreduce_host(SystemState *s, double *v, int n, int op)
{
  LocalGroup &G = *s->local;
  std::unique_lock<std::mutex> lk(G.mu);
  const long long gen = G.red_gen;
  if (G.red_arrived == 0) G.red.assign(v, v + n);
  else
    for (int k = 0; k < n; k++) G.red[k] = op ? fmax(G.red[k], v[k]) : G.red[k] + v[k];
  if (++G.red_arrived == G.nranks) {
    G.red_out = G.red;
    G.red_arrived = 0;
    G.red_gen++;
    G.cv.notify_all();
  } else if (!G.cv.wait_for(lk, std::chrono::seconds(LOCAL_TIMEOUT_S), [&] { return G.red_gen != gen; }))
    return false;
  for (int k = 0; k < n; k++) v[k] = G.red_out[k];
  // nobody may start the next reduction before everyone has read this one: second phase
  const long long gen2 = G.red_gen;
  if (++G.red_arrived == G.nranks) {
    G.red_arrived = 0;
    G.red_gen++;
    G.cv.notify_all();
  } else if (!G.cv.wait_for(lk, std::chrono::seconds(LOCAL_TIMEOUT_S), [&] { return G.red_gen != gen2; }))
    return false;
  return true;
}

// exchange doubles with the swap partners: send to `sendto`, receive from `recvfrom` (stream-ordered with NCCL)
static int xfer_sendrecv(b200md_ctx *c, SystemState *s, const double *sbuf, size_t nsend, int sendto, double *rbuf,
                         size_t nrecv, int recvfrom)
{
  if (s->local) return local_sendrecv(c, s, sbuf, nsend, sendto, rbuf, nrecv, recvfrom);
  NCCL_TRY(c, ncclGroupStart());
  if (nsend) NCCL_TRY(c, ncclSend(sbuf, nsend, ncclDouble, sendto, s->nccl, c->stream));
  if (nrecv) NCCL_TRY(c, ncclRecv(rbuf, nrecv, ncclDouble, recvfrom, s->nccl, c->stream));
  NCCL_TRY(c, ncclGroupEnd());
  return B200MD_OK;
}

// several exchanges that do not depend on each other (the -d and +d swaps of one dimension when one layer of
// neighbors suffices): ONE NCCL group = one fused send/recv kernel instead of one per swap
struct Xfer {
  const double *sbuf;
  size_t nsend;
  int to;
  double *rbuf;
  size_t nrecv;
  int from;
};
static int xfer_multi(b200md_ctx *c, SystemState *s, const Xfer *x, int n)
{
  if (s->local) {
    for (int k = 0; k < n; k++) {
      int rc = local_sendrecv(c, s, x[k].sbuf, x[k].nsend, x[k].to, x[k].rbuf, x[k].nrecv, x[k].from);
      if (rc) return rc;
    }
    return B200MD_OK;
  }
  NCCL_TRY(c, ncclGroupStart());
  for (int k = 0; k < n; k++) {
    if (x[k].nsend) NCCL_TRY(c, ncclSend(x[k].sbuf, x[k].nsend, ncclDouble, x[k].to, s->nccl, c->stream));
    if (x[k].nrecv) NCCL_TRY(c, ncclRecv(x[k].rbuf, x[k].nrecv, ncclDouble, x[k].from, s->nccl, c->stream));
  }
  NCCL_TRY(c, ncclGroupEnd());
  return B200MD_OK;
}

// swaps [first, first + count) of dimension `dim`; `paired` = exactly the -d/+d pair, both with a remote partner
struct DimSwaps {
  int first, count;
  bool paired;
};
static DimSwaps dim_swaps(const SystemState *s, int dim)
{
  DimSwaps d = {0, 0, false};
  int is = 0;
  for (int k = 0; k < dim; k++) is += 2 * s->maxneed[k];
  d.first = is;
  d.count = 2 * s->maxneed[dim];
  d.paired = d.count == 2 && s->swaps[is].sendproc != s->me && s->swaps[is + 1].sendproc != s->me;
  return d;
}

// in-place reductions over ranks of small device arrays
static int xfer_allreduce_sum(b200md_ctx *c, SystemState *s, double *dbuf, int n)
{
  if (s->nranks == 1) return B200MD_OK;
  if (s->local) {
    std::vector<double> h(n);
    CUDA_TRY(c, cudaMemcpyAsync(h.data(), dbuf, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (!local_allreduce_host(s, h.data(), n, 0)) {
      c->fail("loopback transport: allreduce timed out (a peer rank stopped)");
      return B200MD_ERR_NCCL;
    }
    CUDA_TRY(c, cudaMemcpyAsync(dbuf, h.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return B200MD_OK;
  }
  NCCL_TRY(c, ncclAllReduce(dbuf, dbuf, n, ncclDouble, ncclSum, s->nccl, c->stream));
  return B200MD_OK;
}
static int xfer_allreduce_max_int(b200md_ctx *c, SystemState *s, int *dbuf)
{
  if (s->nranks == 1) return B200MD_OK;
  if (s->local) {
    int h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&h, dbuf, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    double v = h;
    if (!local_allreduce_host(s, &v, 1, 1)) {
      c->fail("loopback transport: allreduce timed out (a peer rank stopped)");
      return B200MD_ERR_NCCL;
    }
    h = (int) v;
    CUDA_TRY(c, cudaMemcpyAsync(dbuf, &h, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return B200MD_OK;
  }
  NCCL_TRY(c, ncclAllReduce(dbuf, dbuf, 1, ncclInt, ncclMax, s->nccl, c->stream));
  return B200MD_OK;
}

// ------------------------------------------------------------------ geometry / comm setup
static void setup_geometry(SystemState *s)
{
  const b200md_box &b = s->d.box;
  s->triclinic = b.triclinic;
  for (int d = 0; d < 3; d++) {
    s->boxlo[d] = b.boxlo[d];
    s->boxhi[d] = b.boxhi[d];
    s->prd[d] = b.boxhi[d] - b.boxlo[d];
  }
  // Domain::set_global_box
  s->h[0] = s->prd[0];
  s->h[1] = s->prd[1];
  s->h[2] = s->prd[2];
  s->h[3] = b.yz;
  s->h[4] = b.xz;
  s->h[5] = b.xy;
  s->h_inv[0] = 1.0 / s->h[0];
  s->h_inv[1] = 1.0 / s->h[1];
  s->h_inv[2] = 1.0 / s->h[2];
  s->h_inv[3] = -s->h[3] / (s->h[1] * s->h[2]);
  s->h_inv[4] = (s->h[3] * s->h[5] - s->h[1] * s->h[4]) / (s->h[0] * s->h[1] * s->h[2]);
  s->h_inv[5] = -s->h[5] / (s->h[0] * s->h[1]);
  if (!s->triclinic) s->h[3] = s->h[4] = s->h[5] = s->h_inv[3] = s->h_inv[4] = s->h_inv[5] = 0.0;

  // rank -> brick location, x fastest (same numbering as the host engine)
  const int *pg = s->d.procgrid;
  s->nranks = pg[0] * pg[1] * pg[2];
  s->me = s->d.rank;
  s->myloc[0] = s->me % pg[0];
  s->myloc[1] = (s->me / pg[0]) % pg[1];
  s->myloc[2] = s->me / (pg[0] * pg[1]);
  auto rankof = [&](int ix, int iy, int iz) {
    ix = (ix + pg[0]) % pg[0];
    iy = (iy + pg[1]) % pg[1];
    iz = (iz + pg[2]) % pg[2];
    return iz * pg[1] * pg[0] + iy * pg[0] + ix;
  };
  for (int d = 0; d < 3; d++)
    for (int dir = 0; dir < 2; dir++) {
      int l[3] = {s->myloc[0], s->myloc[1], s->myloc[2]};
      l[d] += dir ? 1 : -1;
      s->procneigh[d][dir] = rankof(l[0], l[1], l[2]);
    }
  // Domain::set_local_box (uniform layout)
  for (int d = 0; d < 3; d++) {
    if (!s->triclinic) {
      s->sublo[d] = s->boxlo[d] + s->prd[d] * (s->myloc[d] * 1.0 / pg[d]);
      if (s->myloc[d] < pg[d] - 1) s->subhi[d] = s->boxlo[d] + s->prd[d] * ((s->myloc[d] + 1) * 1.0 / pg[d]);
      else s->subhi[d] = s->boxhi[d];
    } else {
      s->sublo[d] = s->myloc[d] * 1.0 / pg[d];
      s->subhi[d] = (s->myloc[d] < pg[d] - 1) ? (s->myloc[d] + 1) * 1.0 / pg[d] : 1.0;
    }
  }
}

// CommBrick::setup
static void setup_swaps(SystemState *s)
{
  const double cut = s->cutneighmax;
  double prd[3];
  if (!s->triclinic) {
    for (int d = 0; d < 3; d++) {
      s->cutghost[d] = cut;
      prd[d] = s->prd[d];
    }
  } else {
    const double *hi = s->h_inv;
    s->cutghost[0] = cut * sqrt(hi[0] * hi[0] + hi[5] * hi[5] + hi[4] * hi[4]);
    s->cutghost[1] = cut * sqrt(hi[1] * hi[1] + hi[3] * hi[3]);
    s->cutghost[2] = cut * hi[2];
    prd[0] = prd[1] = prd[2] = 1.0;
  }
  const int *pg = s->d.procgrid;
  for (auto &sw : s->swaps) sw.sendlist.release();
  s->swaps.clear();
  for (int dim = 0; dim < 3; dim++) {
    s->maxneed[dim] = static_cast<int>(s->cutghost[dim] * pg[dim] / prd[dim]) + 1;
    for (int ineed = 0; ineed < 2 * s->maxneed[dim]; ineed++) {
      Swap sw;
      sw.dim = dim;
      if (ineed % 2 == 0) {
        sw.sendproc = s->procneigh[dim][0];
        sw.recvproc = s->procneigh[dim][1];
        sw.lo = (ineed < 2) ? -BIGSLAB : 0.5 * (s->sublo[dim] + s->subhi[dim]);
        sw.hi = s->sublo[dim] + s->cutghost[dim];
        if (s->myloc[dim] == 0) {
          sw.pbc_flag = 1;
          sw.pbc[dim] = 1;
          if (s->triclinic) {
            if (dim == 1) sw.pbc[5] = 1;
            else if (dim == 2) sw.pbc[4] = sw.pbc[3] = 1;
          }
        }
      } else {
        sw.sendproc = s->procneigh[dim][1];
        sw.recvproc = s->procneigh[dim][0];
        sw.lo = s->subhi[dim] - s->cutghost[dim];
        sw.hi = (ineed < 2) ? BIGSLAB : 0.5 * (s->sublo[dim] + s->subhi[dim]);
        if (s->myloc[dim] == pg[dim] - 1) {
          sw.pbc_flag = 1;
          sw.pbc[dim] = -1;
          if (s->triclinic) {
            if (dim == 1) sw.pbc[5] = -1;
            else if (dim == 2) sw.pbc[4] = sw.pbc[3] = -1;
          }
        }
      }
      // AtomVec::pack_border / pack_comm shifts
      const double xy = s->h[5], xz = s->h[4], yz = s->h[3];
      if (!s->triclinic) {
        for (int d = 0; d < 3; d++) sw.bshift[d] = sw.fshift[d] = sw.pbc[d] * s->prd[d];
      } else {
        for (int d = 0; d < 3; d++) sw.bshift[d] = sw.pbc[d];
        sw.fshift[0] = sw.pbc[0] * s->prd[0] + sw.pbc[5] * xy + sw.pbc[4] * xz;
        sw.fshift[1] = sw.pbc[1] * s->prd[1] + sw.pbc[3] * yz;
        sw.fshift[2] = sw.pbc[2] * s->prd[2];
      }
      s->swaps.push_back(std::move(sw));
    }
  }
}

// neighbor cutoffs per LAMMPS type pair (Pair::init -> init_one, Neighbor::init)
static int setup_cutoffs(b200md_ctx *c, SystemState *s)
{
  const int nt = s->d.ntypes;
  const double skin = s->d.skin;
  s->cutneighsq.assign((size_t) (nt + 1) * (nt + 1), 0.0);
  s->cutneighghostsq.assign((size_t) (nt + 1) * (nt + 1), 0.0);
  s->cutneighmax = 0.0;
  // Pair::init:      cutsq[i][j] = init_one(i,j)^2 for i <= j, mirrored
  // Neighbor::init:  cutneighsq = (sqrt(cutsq) + skin)^2 ; cutneighghostsq = (cutghost + skin)^2
  // (same operation order as the host code, so the squared cutoffs are bit-identical)
  if (s->d.style == 0) {
    ARG_CHECK(c, c->rebomos_ready && c->ntypes == nt, "system_create: call b200md_rebomos_init with the same ntypes first");
    const double cut3rebo = 3.0 * c->rp.rcmax[0];    // PairREBOMoS::init_one: same cutoff for every type pair
    for (int i = 1; i <= nt; i++)
      for (int j = 1; j <= nt; j++) {
        ARG_CHECK(c, c->map_h[i] >= 0 && c->map_h[j] >= 0, "system_create: NULL-mapped types are not supported");
        const double cutsq = cut3rebo * cut3rebo;
        const double cn = sqrt(cutsq) + skin;
        s->cutneighsq[(size_t) i * (nt + 1) + j] = cn * cn;
        const double cg = c->rp.rcmax[c->map_h[i] * 2 + c->map_h[j]] + skin;    // cutghost[i][j] = rcmax
        s->cutneighghostsq[(size_t) i * (nt + 1) + j] = cg * cg;
        s->cutneighmax = fmax(s->cutneighmax, cn);
      }
    s->ghost_rows = 1;
  } else {
    ARG_CHECK(c, c->aeam_ready && c->ap.nel == nt, "system_create: call b200md_aeam_init with nelements == ntypes first");
    for (int i = 1; i <= nt; i++)
      for (int j = 1; j <= nt; j++) {
        // Pair::init loops i <= j and mirrors cutsq, so the (i,j) cutoff with i <= j is used for both
        const int a = i <= j ? i : j, b = i <= j ? j : i;
        const double cut = c->ap.cut[(a - 1) * nt + (b - 1)];
        const double cutsq = cut * cut;
        const double cn = sqrt(cutsq) + skin;
        s->cutneighsq[(size_t) i * (nt + 1) + j] = cn * cn;
        s->cutneighghostsq[(size_t) i * (nt + 1) + j] = cn * cn;
        s->cutneighmax = fmax(s->cutneighmax, cn);
      }
    s->ghost_rows = 0;
  }
  return B200MD_OK;
}

// ------------------------------------------------------------------ Atom::sort
static int sort_atoms(b200md_ctx *c, SystemState *s)
{
  const int n = s->nlocal;
  s->nextsort = (s->step / s->d.sort_every) * s->d.sort_every + s->d.sort_every;
  // Atom::setup_sort_bins
  const double binsize = 0.5 * s->cutneighmax;
  if (binsize == 0.0 || n == 0) return B200MD_OK;
  const double bininv = 1.0 / binsize;
  double lo[3], hi[3];
  if (!s->triclinic) {
    for (int d = 0; d < 3; d++) {
      lo[d] = s->sublo[d];
      hi[d] = s->subhi[d];
    }
  } else {
    for (int d = 0; d < 3; d++) {
      lo[d] = 1.0e30;
      hi[d] = -1.0e30;
    }
    for (int cc = 0; cc < 8; cc++) {
      const double l0 = (cc & 1) ? s->subhi[0] : s->sublo[0], l1 = (cc & 2) ? s->subhi[1] : s->sublo[1],
                   l2 = (cc & 4) ? s->subhi[2] : s->sublo[2];
      double x[3];
      x[0] = s->h[0] * l0 + s->h[5] * l1 + s->h[4] * l2 + s->boxlo[0];
      x[1] = s->h[1] * l1 + s->h[3] * l2 + s->boxlo[1];
      x[2] = s->h[2] * l2 + s->boxlo[2];
      for (int d = 0; d < 3; d++) {
        lo[d] = fmin(lo[d], x[d]);
        hi[d] = fmax(hi[d], x[d]);
      }
    }
  }
  SortGeom g;
  long long nbins = 1;
  for (int d = 0; d < 3; d++) {
    g.nb[d] = static_cast<int>((hi[d] - lo[d]) * bininv);
    if (g.nb[d] == 0) g.nb[d] = 1;
    g.inv[d] = g.nb[d] / (hi[d] - lo[d]);
    g.lo[d] = lo[d];
    nbins *= g.nb[d];
  }
  ARG_CHECK(c, nbins < 2000000000LL, "Too many atom sorting bins");
  if (nbins == 1) return B200MD_OK;
  const Geom geo = make_geom(s);
  // for triclinic, atoms must be in box coords (not lamda) to match bbox; LAMMPS converts back after
  if (s->triclinic) {
    LaunchScope ls(c, "lamda2x");
    k_lamda2x<<<nblk(n), BLOCK, 0, c->stream>>>(geo, c->xq.p, n);
  }
  CUDA_TRY(c, s->itmp.reserve((size_t) 2 * n + nbins + 64));
  CUDA_TRY(c, s->itmp2.reserve((size_t) n + nbins + 64));
  CUDA_TRY(c, s->scan64.reserve((size_t) nbins + 8));
  int *bin_of = s->itmp.p, *count = s->itmp.p + n;
  int *perm = s->itmp2.p;
  CUDA_TRY(c, cudaMemsetAsync(count, 0, nbins * sizeof(int), c->stream));
  {
    LaunchScope ls(c, "sort_bins");
    k_sort_bins<<<nblk(n), BLOCK, 0, c->stream>>>(g, c->xq.p, n, bin_of, count);
  }
  int rc = b200md_exclusive_scan_i64(c, count, s->scan64.p, (int) nbins, 1);
  if (rc) return rc;
  CUDA_TRY(c, cudaMemsetAsync(count, 0, nbins * sizeof(int), c->stream));
  {
    LaunchScope ls(c, "sort_fill");
    k_sort_fill<<<nblk(n), BLOCK, 0, c->stream>>>(bin_of, n, s->scan64.p, count, perm);
  }
  {
    LaunchScope ls(c, "sort_within");
    k_sort_within<<<nblk(nbins), BLOCK, 0, c->stream>>>(s->scan64.p, (int) nbins, perm);
  }
  if (s->triclinic) {
    LaunchScope ls(c, "x2lamda");
    k_x2lamda<<<nblk(n), BLOCK, 0, c->stream>>>(geo, c->xq.p, n);
  }
  // gather into temporaries, then copy back
  CUDA_TRY(c, s->x4tmp.reserve((size_t) n + 8));
  CUDA_TRY(c, s->dtmp.reserve(3 * (size_t) n + 8));
  int *to = s->itmp.p, *go = s->itmp.p + n;    // bin_of/count are no longer needed
  {
    LaunchScope ls(c, "permute");
    k_permute<<<nblk(n), BLOCK, 0, c->stream>>>(perm, n, c->xq.p, s->v.p, c->type.p, c->tag.p, s->x4tmp.p, s->dtmp.p, to, go);
  }
  CUDA_TRY(c, cudaMemcpyAsync(c->xq.p, s->x4tmp.p, n * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(s->v.p, s->dtmp.p, 3 * (size_t) n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(c->type.p, to, n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(c->tag.p, go, n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

static int p2p_setup(b200md_ctx *c, SystemState *s);

// ------------------------------------------------------------------ CommBrick::borders
static int halo_borders(b200md_ctx *c, SystemState *s)
{
  s->nghost = 0;
  int nfirst = 0, nlast = 0;
  size_t is = 0;
  for (int dim = 0; dim < 3; dim++) {
    nlast = 0;
    for (int ineed = 0; ineed < 2 * s->maxneed[dim]; ineed++, is++) {
      Swap &sw = s->swaps[is];
      if (ineed % 2 == 0) {
        nfirst = nlast;
        nlast = s->nlocal + s->nghost;
      }
      const int nscan = nlast - nfirst;
      int nsend = 0;
      if (nscan > 0) {
        CUDA_TRY(c, s->itmp.reserve((size_t) nscan + 64));
        CUDA_TRY(c, s->scan64.reserve((size_t) nscan + 8));
        {
          LaunchScope ls(c, "slab_flags");
          k_slab_flags<<<nblk(nscan), BLOCK, 0, c->stream>>>(c->xq.p, nfirst, nlast, sw.dim, sw.lo, sw.hi, s->itmp.p);
        }
        int rc = b200md_exclusive_scan_i64(c, s->itmp.p, s->scan64.p, nscan, 1);
        if (rc) return rc;
        int64_t tot = 0;
        CUDA_TRY(c, cudaMemcpyAsync(&tot, s->scan64.p + nscan, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        nsend = (int) tot;
        CUDA_TRY(c, sw.sendlist.reserve((size_t) nsend + 8));
        if (nsend) {
          LaunchScope ls(c, "scatter_list");
          k_scatter_list<<<nblk(nscan), BLOCK, 0, c->stream>>>(s->itmp.p, s->scan64.p, nfirst, nscan, sw.sendlist.p);
        }
      }
      sw.nsend = nsend;
      const int first = s->nlocal + s->nghost;
      int nrecv = 0;
      if (sw.sendproc == s->me) {
        nrecv = nsend;
        int rc = ensure_atoms(c, s, (size_t) first + nrecv);
        if (rc) return rc;
        if (nrecv) {
          LaunchScope ls(c, "border_copy");
          k_border_copy<<<nblk(nrecv), BLOCK, 0, c->stream>>>(c->xq.p, c->type.p, c->tag.p, sw.sendlist.p, nrecv, first,
                                                            sw.bshift[0], sw.bshift[1], sw.bshift[2], sw.pbc_flag);
        }
      } else {
        // counts first (host round trip is fine at rebuild time), then the atoms
        CUDA_TRY(c, s->sendbuf.reserve(6 * (size_t) nsend + 8));
        double cnt_h = (double) nsend, cnt_r = 0.0;
        double *dcnt = c->scal.p + 32;
        CUDA_TRY(c, cudaMemcpyAsync(dcnt, &cnt_h, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        int rc = xfer_sendrecv(c, s, dcnt, 1, sw.sendproc, dcnt + 1, 1, sw.recvproc);
        if (rc) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(&cnt_r, dcnt + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        nrecv = (int) cnt_r;
        if ((rc = ensure_atoms(c, s, (size_t) first + nrecv))) return rc;
        CUDA_TRY(c, s->recvbuf.reserve(6 * (size_t) nrecv + 8));
        if (nsend) {
          LaunchScope ls(c, "border_pack");
          k_border_pack<<<nblk(nsend), BLOCK, 0, c->stream>>>(c->xq.p, c->type.p, c->tag.p, sw.sendlist.p, nsend,
                                                            sw.bshift[0], sw.bshift[1], sw.bshift[2], sw.pbc_flag,
                                                            s->sendbuf.p);
        }
        if ((rc = xfer_sendrecv(c, s, s->sendbuf.p, 6 * (size_t) nsend, sw.sendproc, s->recvbuf.p, 6 * (size_t) nrecv,
                                sw.recvproc)))
          return rc;
        if (nrecv) {
          LaunchScope ls(c, "border_unpack");
          k_border_unpack<<<nblk(nrecv), BLOCK, 0, c->stream>>>(c->xq.p, c->type.p, c->tag.p, first, nrecv, s->recvbuf.p);
        }
      }
      sw.nrecv = nrecv;
      sw.firstrecv = first;
      s->nghost += nrecv;
    }
  }
  CUDA_TRY(c, cudaGetLastError());
  // one rank, one layer of images per dimension: set the flat halo up (the swaps in order: an image of an image finds
  // its source's entry already written)
  s->flat_ok = false;
  if (c->flat_halo && s->nranks == 1 && s->swaps.size() == 6 && s->maxneed[0] == 1 && s->maxneed[1] == 1 && s->maxneed[2] == 1 &&
      s->nghost > 0) {
    bool all_self = true;
    for (const Swap &sw : s->swaps) all_self = all_self && sw.sendproc == s->me && sw.nrecv == sw.nsend;
    if (all_self) {
      CUDA_TRY(c, s->flat_src.reserve((size_t) s->nghost + 8));
      CUDA_TRY(c, s->flat_code.reserve((size_t) s->nghost + 8));
      for (size_t k = 0; k < s->swaps.size(); k++) {
        const Swap &sw = s->swaps[k];
        if (!sw.nrecv) continue;
        LaunchScope ls(c, "border_copy");
        k_flat_build<<<nblk(sw.nrecv), BLOCK, 0, c->stream>>>(sw.sendlist.p, sw.nrecv, sw.firstrecv, s->nlocal, (int) k, sw.dim,
                                                             s->flat_src.p, s->flat_code.p);
      }
      CUDA_TRY(c, cudaGetLastError());
      s->flat_ok = true;
    }
  }
  return p2p_setup(c, s);
}

// ------------------------------------------------------------------ peer-memory halo: windows + IPC mapping
static void p2p_release(SystemState *s)
{
  SystemState::PeerHalo &P = s->p2p;
  for (size_t r = 0; r < P.pwin.size(); r++) {
    if (P.pwin[r]) cudaIpcCloseMemHandle(P.pwin[r]);
    if (P.pflag[r]) cudaIpcCloseMemHandle(P.pflag[r]);
  }
  P.pwin.clear();
  P.pflag.clear();
  P.vote_peers.release();
  if (P.win) cudaFree(P.win);
  if (P.flag) cudaFree(P.flag);
  if (P.done) cudaFree(P.done);
  P.win = nullptr;
  P.flag = P.done = nullptr;
  P.slot_cap = 0;
  P.ok = false;
}

// Collective (called by every rank at every list rebuild, after borders): make sure each rank owns a window whose
// slots hold the largest halo message of any rank, and that every rank has every other rank's window mapped.
// Any failure (IPC not permitted, no peer access) switches the peer path off on ALL ranks; NCCL send/recv stays.
static int p2p_setup(b200md_ctx *c, SystemState *s)
{
  SystemState::PeerHalo &P = s->p2p;
  P.want = c->p2p_halo;
  if (!s->nccl || s->local || s->nranks == 1 || !P.want) {
    P.ok = false;
    return B200MD_OK;
  }
  double need = 0.0;
  for (const Swap &sw : s->swaps) need = fmax(need, (double) (sw.nsend > sw.nrecv ? sw.nsend : sw.nrecv));
  double *dn = c->scal.p + 44;
  CUDA_TRY(c, cudaMemcpyAsync(dn, &need, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, ncclAllReduce(dn, dn, 1, ncclDouble, ncclMax, s->nccl, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(&need, dn, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  const size_t want_cap = 3 * (size_t) need + 64;
  if (P.ok && want_cap <= P.slot_cap) return B200MD_OK;
  if (P.slot_cap == (size_t) -1) return B200MD_OK;    // tried before and failed: stay on NCCL
  p2p_release(s);
  const size_t cap = (want_cap + want_cap / 4 + 1) & ~(size_t) 1;    // even: every slot starts on a 16-byte boundary
  double okv = 1.0;
  if (cudaMalloc((void **) &P.win, (size_t) P2P_SLOTS * 2 * cap * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void **) &P.flag, (P2P_SLOTS * 2 + 2 * VOTE_MAXR) * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **) &P.done, 4 * sizeof(int)) != cudaSuccess)
    okv = 0.0;
  struct Handles {
    cudaIpcMemHandle_t win, flag;
  } mine;
  memset(&mine, 0, sizeof(mine));
  if (okv > 0.0) {
    cudaMemsetAsync(P.flag, 0, (P2P_SLOTS * 2 + 2 * VOTE_MAXR) * sizeof(int), c->stream);
    cudaMemsetAsync(P.done, 0, 4 * sizeof(int), c->stream);
    if (cudaIpcGetMemHandle(&mine.win, P.win) != cudaSuccess || cudaIpcGetMemHandle(&mine.flag, P.flag) != cudaSuccess)
      okv = 0.0;
  }
  cudaGetLastError();
  // everybody's handles to everybody (tiny all-gather on the library's own communicator)
  DevBuf<char> hb;
  CUDA_TRY(c, hb.reserve(sizeof(Handles) * ((size_t) s->nranks + 1)));
  CUDA_TRY(c, cudaMemcpyAsync(hb.p, &mine, sizeof(Handles), cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, ncclAllGather(hb.p, hb.p + sizeof(Handles), sizeof(Handles), ncclChar, s->nccl, c->stream));
  std::vector<Handles> all(s->nranks);
  CUDA_TRY(c, cudaMemcpyAsync(all.data(), hb.p + sizeof(Handles), sizeof(Handles) * s->nranks, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  hb.release();
  P.pwin.assign(s->nranks, nullptr);
  P.pflag.assign(s->nranks, nullptr);
  if (okv > 0.0) {
    for (int r = 0; r < s->nranks && okv > 0.0; r++) {
      if (r == s->me) continue;
      void *pw = nullptr, *pf = nullptr;
      if (cudaIpcOpenMemHandle(&pw, all[r].win, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pf, all[r].flag, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        okv = 0.0;
      P.pwin[r] = (double *) pw;
      P.pflag[r] = (int *) pf;
    }
  }
  cudaGetLastError();
  CUDA_TRY(c, cudaMemcpyAsync(dn, &okv, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, ncclAllReduce(dn, dn, 1, ncclDouble, ncclMin, s->nccl, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(&okv, dn, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (okv > 0.0 && s->nranks <= VOTE_MAXR) {
    std::vector<int *> vp(s->nranks);
    for (int r = 0; r < s->nranks; r++) vp[r] = (r == s->me ? P.flag : P.pflag[r]) + P2P_SLOTS * 2;
    if (P.vote_peers.reserve((size_t) s->nranks + 1) != cudaSuccess ||
        cudaMemcpyAsync(P.vote_peers.p, vp.data(), s->nranks * sizeof(int *), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
      P.vote_peers.release();
  }
  if (okv > 0.0) {
    P.ok = true;
    P.slot_cap = cap;
    P.epoch[0] = P.epoch[1] = P.epoch[2] = 0;
  } else {
    p2p_release(s);
    P.slot_cap = (size_t) -1;
  }
  return B200MD_OK;
}

// blocks of a push kernel: enough for one item per thread, at most two per SM (one fence per block)
static inline int push_grid(const b200md_ctx *c, long long items)
{
  long long nb = (items + BLOCK - 1) / BLOCK;
  if (nb < 1) nb = 1;
  if (nb > 2LL * c->num_sms) nb = 2LL * c->num_sms;
  return (int) nb;
}

// slot of message kind (0 forward x, 1 forward rho/fp, 2 reverse f) for swap (dim, dir) and epoch parity
static inline size_t p2p_slot(int kind, int dim, int dir, int epoch) { return (size_t) ((kind * 6 + dim * 2 + dir) * 2 + (epoch & 1)); }

// one swap of the forward position halo, self or remote, on its own
static int forward_x_one(b200md_ctx *c, SystemState *s, Swap &sw)
{
  if (sw.sendproc == s->me) {
    if (sw.nsend) {
      LaunchScope ls(c, "forward_x");
      k_forward_x<<<nblk(sw.nsend), BLOCK, 0, c->stream>>>(c->xq.p, sw.sendlist.p, sw.nsend, sw.firstrecv, sw.fshift[0],
                                                         sw.fshift[1], sw.fshift[2], sw.pbc_flag);
    }
    return B200MD_OK;
  }
  CUDA_TRY(c, s->sendbuf.reserve(3 * (size_t) sw.nsend + 8));
  CUDA_TRY(c, s->recvbuf.reserve(3 * (size_t) sw.nrecv + 8));
  if (sw.nsend) {
    LaunchScope ls(c, "forward_x_pack");
    k_forward_x_pack<<<nblk(sw.nsend), BLOCK, 0, c->stream>>>(c->xq.p, sw.sendlist.p, sw.nsend, sw.fshift[0], sw.fshift[1],
                                                            sw.fshift[2], sw.pbc_flag, s->sendbuf.p);
  }
  int rc = xfer_sendrecv(c, s, s->sendbuf.p, 3 * (size_t) sw.nsend, sw.sendproc, s->recvbuf.p, 3 * (size_t) sw.nrecv,
                         sw.recvproc);
  if (rc) return rc;
  if (sw.nrecv) {
    LaunchScope ls(c, "forward_x_unpack");
    k_forward_x_unpack<<<nblk(sw.nrecv), BLOCK, 0, c->stream>>>(c->xq.p, sw.firstrecv, sw.nrecv, s->recvbuf.p);
  }
  return B200MD_OK;
}

static FlatShifts flat_shifts(const SystemState *s)
{
  FlatShifts sh;
  for (int k = 0; k < 6; k++) {
    for (int d = 0; d < 3; d++) sh.s[k][d] = s->swaps[k].fshift[d];
    sh.pbc[k] = s->swaps[k].pbc_flag;
  }
  return sh;
}

static int halo_forward_x(b200md_ctx *c, SystemState *s)
{
  if (s->flat_ok) {
    LaunchScope ls(c, "forward_x");
    k_flat_forward_x<<<nblk(s->nghost), BLOCK, 0, c->stream>>>(c->xq.p, s->flat_src.p, s->flat_code.p, s->nghost, s->nlocal,
                                                              flat_shifts(s));
    CUDA_TRY(c, cudaGetLastError());
    return B200MD_OK;
  }
  s->p2p.epoch[0]++;    // one epoch per halo call, whichever dimensions take the peer path
  for (int dim = 0; dim < 3; dim++) {
    const DimSwaps d = dim_swaps(s, dim);
    if (!d.paired) {
      if (d.count == 2 && s->swaps[d.first].sendproc == s->me && s->swaps[d.first + 1].sendproc == s->me) {
        const Swap &a = s->swaps[d.first], &b = s->swaps[d.first + 1];
        if (a.nsend + b.nsend) {
          LaunchScope ls(c, "forward_x");
          k_forward_x2<<<nblk(a.nsend + b.nsend), BLOCK, 0, c->stream>>>(
              c->xq.p, a.sendlist.p, a.nsend, a.firstrecv, a.fshift[0], a.fshift[1], a.fshift[2], a.pbc_flag, b.sendlist.p,
              b.nsend, b.firstrecv, b.fshift[0], b.fshift[1], b.fshift[2], b.pbc_flag);
        }
        continue;
      }
      for (int k = 0; k < d.count; k++) {
        int rc = forward_x_one(c, s, s->swaps[d.first + k]);
        if (rc) return rc;
      }
      continue;
    }
    // -d and +d swaps scan the same atoms (CommBrick::borders updates its window every second swap): independent
    Swap &a = s->swaps[d.first], &b = s->swaps[d.first + 1];
    if (s->p2p.ok) {
      SystemState::PeerHalo &P = s->p2p;
      const int ep = P.epoch[0];
      PushDesc pd;
      const Swap *sw2[2] = {&a, &b};
      for (int w = 0; w < 2; w++) {
        const size_t slot = p2p_slot(0, dim, w, ep);
        pd.list[w] = sw2[w]->sendlist.p;
        pd.n[w] = sw2[w]->nsend;
        pd.first[w] = 0;
        pd.dst[w] = P.pwin[sw2[w]->sendproc] + slot * P.slot_cap;
        pd.flag[w] = P.pflag[sw2[w]->sendproc] + slot;
        for (int k = 0; k < 3; k++) pd.shift[w][k] = sw2[w]->fshift[k];
        pd.pbc[w] = sw2[w]->pbc_flag;
      }
      {
        LaunchScope ls(c, "p2p_push_x");
        k_p2p_push_x<<<push_grid(c, max(a.nsend, b.nsend)), BLOCK, 0, c->stream>>>(c->xq.p, pd, ep, P.done);
      }
      {
        LaunchScope ls(c, "p2p_unpack_x");
        const size_t s0 = p2p_slot(0, dim, 0, ep), s1 = p2p_slot(0, dim, 1, ep);
        k_p2p_wait<<<1, 32, 0, c->stream>>>(P.flag + s0, P.flag + s1, ep);
        k_p2p_unpack_x<<<max(1, nblk(a.nrecv + b.nrecv)), BLOCK, 0, c->stream>>>(
            c->xq.p, P.win + s0 * P.slot_cap, a.firstrecv, a.nrecv, nullptr, P.win + s1 * P.slot_cap, b.firstrecv,
            b.nrecv, nullptr, ep);
      }
      c->n_p2p++;
      continue;
    }
    const size_t sa = 3 * (size_t) a.nsend, sb = 3 * (size_t) b.nsend, ra = 3 * (size_t) a.nrecv, rb = 3 * (size_t) b.nrecv;
    CUDA_TRY(c, s->sendbuf.reserve(sa + sb + 16));
    CUDA_TRY(c, s->recvbuf.reserve(ra + rb + 16));
    if (a.nsend) {
      LaunchScope ls(c, "forward_x_pack");
      k_forward_x_pack<<<nblk(a.nsend), BLOCK, 0, c->stream>>>(c->xq.p, a.sendlist.p, a.nsend, a.fshift[0], a.fshift[1],
                                                             a.fshift[2], a.pbc_flag, s->sendbuf.p);
    }
    if (b.nsend) {
      LaunchScope ls(c, "forward_x_pack");
      k_forward_x_pack<<<nblk(b.nsend), BLOCK, 0, c->stream>>>(c->xq.p, b.sendlist.p, b.nsend, b.fshift[0], b.fshift[1],
                                                             b.fshift[2], b.pbc_flag, s->sendbuf.p + sa);
    }
    const Xfer x[2] = {{s->sendbuf.p, sa, a.sendproc, s->recvbuf.p, ra, a.recvproc},
                       {s->sendbuf.p + sa, sb, b.sendproc, s->recvbuf.p + ra, rb, b.recvproc}};
    int rc = xfer_multi(c, s, x, 2);
    if (rc) return rc;
    if (a.nrecv) {
      LaunchScope ls(c, "forward_x_unpack");
      k_forward_x_unpack<<<nblk(a.nrecv), BLOCK, 0, c->stream>>>(c->xq.p, a.firstrecv, a.nrecv, s->recvbuf.p);
    }
    if (b.nrecv) {
      LaunchScope ls(c, "forward_x_unpack");
      k_forward_x_unpack<<<nblk(b.nrecv), BLOCK, 0, c->stream>>>(c->xq.p, b.firstrecv, b.nrecv, s->recvbuf.p + ra);
    }
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

static int forward_rho_fp_one(b200md_ctx *c, SystemState *s, Swap &sw)
{
  if (sw.sendproc == s->me) {
    if (sw.nsend) {
      LaunchScope ls(c, "forward_fp");
      k_forward_s2<<<nblk(sw.nsend), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, sw.sendlist.p, sw.nsend, sw.firstrecv);
    }
    return B200MD_OK;
  }
  CUDA_TRY(c, s->sendbuf.reserve(2 * (size_t) sw.nsend + 8));
  CUDA_TRY(c, s->recvbuf.reserve(2 * (size_t) sw.nrecv + 8));
  if (sw.nsend) {
    LaunchScope ls(c, "forward_fp_pack");
    k_forward_s2_pack<<<nblk(sw.nsend), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, sw.sendlist.p, sw.nsend, s->sendbuf.p);
  }
  int rc = xfer_sendrecv(c, s, s->sendbuf.p, 2 * (size_t) sw.nsend, sw.sendproc, s->recvbuf.p, 2 * (size_t) sw.nrecv,
                         sw.recvproc);
  if (rc) return rc;
  if (sw.nrecv) {
    LaunchScope ls(c, "forward_fp_unpack");
    k_forward_s2_unpack<<<nblk(sw.nrecv), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, sw.firstrecv, sw.nrecv, s->recvbuf.p);
  }
  return B200MD_OK;
}

static int halo_forward_rho_fp(b200md_ctx *c, SystemState *s)
{
  if (s->flat_ok) {
    LaunchScope ls(c, "forward_fp");
    k_flat_forward_s2<<<nblk(s->nghost), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, s->flat_src.p, s->nghost, s->nlocal);
    CUDA_TRY(c, cudaGetLastError());
    return B200MD_OK;
  }
  s->p2p.epoch[1]++;
  for (int dim = 0; dim < 3; dim++) {
    const DimSwaps d = dim_swaps(s, dim);
    if (!d.paired) {
      for (int k = 0; k < d.count; k++) {
        int rc = forward_rho_fp_one(c, s, s->swaps[d.first + k]);
        if (rc) return rc;
      }
      continue;
    }
    Swap &a = s->swaps[d.first], &b = s->swaps[d.first + 1];
    if (s->p2p.ok) {
      SystemState::PeerHalo &P = s->p2p;
      const int ep = P.epoch[1];
      PushDesc pd;
      const Swap *sw2[2] = {&a, &b};
      for (int w = 0; w < 2; w++) {
        const size_t slot = p2p_slot(1, dim, w, ep);
        pd.list[w] = sw2[w]->sendlist.p;
        pd.n[w] = sw2[w]->nsend;
        pd.first[w] = 0;
        pd.dst[w] = P.pwin[sw2[w]->sendproc] + slot * P.slot_cap;
        pd.flag[w] = P.pflag[sw2[w]->sendproc] + slot;
        for (int k = 0; k < 3; k++) pd.shift[w][k] = 0.0;
        pd.pbc[w] = 0;
      }
      {
        LaunchScope ls(c, "p2p_push_fp");
        k_p2p_push_s2<<<push_grid(c, max(a.nsend, b.nsend)), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, pd, ep, P.done + 1);
      }
      {
        LaunchScope ls(c, "p2p_unpack_fp");
        const size_t s0 = p2p_slot(1, dim, 0, ep), s1 = p2p_slot(1, dim, 1, ep);
        k_p2p_wait<<<1, 32, 0, c->stream>>>(P.flag + s0, P.flag + s1, ep);
        k_p2p_unpack_s2<<<max(1, nblk(a.nrecv + b.nrecv)), BLOCK, 0, c->stream>>>(
            c->rho.p, c->fp.p, P.win + s0 * P.slot_cap, a.firstrecv, a.nrecv, nullptr, P.win + s1 * P.slot_cap,
            b.firstrecv, b.nrecv, nullptr, ep);
      }
      continue;
    }
    const size_t sa = 2 * (size_t) a.nsend, sb = 2 * (size_t) b.nsend, ra = 2 * (size_t) a.nrecv, rb = 2 * (size_t) b.nrecv;
    CUDA_TRY(c, s->sendbuf.reserve(sa + sb + 16));
    CUDA_TRY(c, s->recvbuf.reserve(ra + rb + 16));
    if (a.nsend) {
      LaunchScope ls(c, "forward_fp_pack");
      k_forward_s2_pack<<<nblk(a.nsend), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, a.sendlist.p, a.nsend, s->sendbuf.p);
    }
    if (b.nsend) {
      LaunchScope ls(c, "forward_fp_pack");
      k_forward_s2_pack<<<nblk(b.nsend), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, b.sendlist.p, b.nsend, s->sendbuf.p + sa);
    }
    const Xfer x[2] = {{s->sendbuf.p, sa, a.sendproc, s->recvbuf.p, ra, a.recvproc},
                       {s->sendbuf.p + sa, sb, b.sendproc, s->recvbuf.p + ra, rb, b.recvproc}};
    int rc = xfer_multi(c, s, x, 2);
    if (rc) return rc;
    if (a.nrecv) {
      LaunchScope ls(c, "forward_fp_unpack");
      k_forward_s2_unpack<<<nblk(a.nrecv), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, a.firstrecv, a.nrecv, s->recvbuf.p);
    }
    if (b.nrecv) {
      LaunchScope ls(c, "forward_fp_unpack");
      k_forward_s2_unpack<<<nblk(b.nrecv), BLOCK, 0, c->stream>>>(c->rho.p, c->fp.p, b.firstrecv, b.nrecv, s->recvbuf.p + ra);
    }
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

static int reverse_f_one(b200md_ctx *c, SystemState *s, Swap &sw)
{
  if (sw.sendproc == s->me) {
    if (sw.nsend) {
      LaunchScope ls(c, "reverse_f");
      k_reverse_f<<<nblk(sw.nsend), BLOCK, 0, c->stream>>>(c->f.p, sw.sendlist.p, sw.nsend, sw.firstrecv, s->fold_atomic);
    }
    return B200MD_OK;
  }
  // ghost forces are contiguous at f[3*firstrecv ...]: send them back, add what my send-list atoms receive
  CUDA_TRY(c, s->recvbuf.reserve(3 * (size_t) sw.nsend + 8));
  int rc = xfer_sendrecv(c, s, c->f.p + 3 * (size_t) sw.firstrecv, 3 * (size_t) sw.nrecv, sw.recvproc, s->recvbuf.p,
                         3 * (size_t) sw.nsend, sw.sendproc);
  if (rc) return rc;
  if (sw.nsend) {
    LaunchScope ls(c, "reverse_f_unpack");
    k_reverse_f_unpack<<<nblk(sw.nsend), BLOCK, 0, c->stream>>>(c->f.p, sw.sendlist.p, sw.nsend, s->recvbuf.p, s->fold_atomic);
  }
  return B200MD_OK;
}

static int halo_reverse_f(b200md_ctx *c, SystemState *s)
{
  if (s->flat_ok && !c->deterministic) {    // every image's force straight to its owned atom (atomic folds, as the staged ones)
    LaunchScope ls(c, "reverse_f");
    k_flat_reverse_f<<<nblk(s->nghost), BLOCK, 0, c->stream>>>(c->f.p, s->flat_src.p, s->nghost, s->nlocal);
    CUDA_TRY(c, cudaGetLastError());
    return B200MD_OK;
  }
  s->p2p.epoch[2]++;
  for (int dim = 2; dim >= 0; dim--) {
    const DimSwaps d = dim_swaps(s, dim);
    if (!d.paired) {
      if (d.count == 2 && !c->deterministic && s->swaps[d.first].sendproc == s->me &&
          s->swaps[d.first + 1].sendproc == s->me) {
        const Swap &a = s->swaps[d.first], &b = s->swaps[d.first + 1];
        if (a.nsend + b.nsend) {
          LaunchScope ls(c, "reverse_f");
          k_reverse_f2<<<nblk(a.nsend + b.nsend), BLOCK, 0, c->stream>>>(c->f.p, a.sendlist.p, a.nsend, a.firstrecv,
                                                                         b.sendlist.p, b.nsend, b.firstrecv);
        }
        continue;
      }
      for (int k = d.count - 1; k >= 0; k--) {
        int rc = reverse_f_one(c, s, s->swaps[d.first + k]);
        if (rc) return rc;
      }
      continue;
    }
    // the +d swap is folded first, then the -d swap, as in the sequential order; their ghost ranges are disjoint and
    // neither send list contains the other's ghosts, so both transfers travel in one group
    Swap &a = s->swaps[d.first], &b = s->swaps[d.first + 1];
    if (s->p2p.ok) {
      SystemState::PeerHalo &P = s->p2p;
      const int ep = P.epoch[2];
      PushDesc pd;
      const Swap *sw2[2] = {&a, &b};
      for (int w = 0; w < 2; w++) {
        // my ghosts of swap (dim, w) came from recvproc; their forces go back into ITS slot (dim, w)
        const size_t slot = p2p_slot(2, dim, w, ep);
        pd.list[w] = nullptr;
        pd.n[w] = sw2[w]->nrecv;
        pd.first[w] = sw2[w]->firstrecv;
        pd.dst[w] = P.pwin[sw2[w]->recvproc] + slot * P.slot_cap;
        pd.flag[w] = P.pflag[sw2[w]->recvproc] + slot;
        for (int k = 0; k < 3; k++) pd.shift[w][k] = 0.0;
        pd.pbc[w] = 0;
      }
      {
        LaunchScope ls(c, "p2p_push_f");
        k_p2p_push_f<<<push_grid(c, (3 * (long long) max(a.nrecv, b.nrecv) + 1) / 2), BLOCK, 0, c->stream>>>(c->f.p, pd, ep, P.done + 2);
      }
      // fold +d first, then -d, in two launches: an atom may sit in both send lists
      {
        LaunchScope ls(c, "p2p_unpack_f");
        k_p2p_wait<<<1, 32, 0, c->stream>>>(P.flag + p2p_slot(2, dim, 0, ep), P.flag + p2p_slot(2, dim, 1, ep), ep);
      }
      for (int w = 1; w >= 0; w--) {
        LaunchScope ls(c, "p2p_unpack_f");
        const size_t slot = p2p_slot(2, dim, w, ep);
        k_p2p_unpack_f<<<max(1, nblk(sw2[w]->nsend)), BLOCK, 0, c->stream>>>(c->f.p, sw2[w]->sendlist.p, sw2[w]->nsend,
                                                                           P.win + slot * P.slot_cap, nullptr, ep,
                                                                           s->fold_atomic);
      }
      continue;
    }
    const size_t sa = 3 * (size_t) a.nsend, sb = 3 * (size_t) b.nsend;
    CUDA_TRY(c, s->recvbuf.reserve(sa + sb + 16));
    const Xfer x[2] = {{c->f.p + 3 * (size_t) b.firstrecv, 3 * (size_t) b.nrecv, b.recvproc, s->recvbuf.p + sa, sb, b.sendproc},
                       {c->f.p + 3 * (size_t) a.firstrecv, 3 * (size_t) a.nrecv, a.recvproc, s->recvbuf.p, sa, a.sendproc}};
    int rc = xfer_multi(c, s, x, 2);
    if (rc) return rc;
    if (b.nsend) {
      LaunchScope ls(c, "reverse_f_unpack");
      k_reverse_f_unpack<<<nblk(b.nsend), BLOCK, 0, c->stream>>>(c->f.p, b.sendlist.p, b.nsend, s->recvbuf.p + sa, s->fold_atomic);
    }
    if (a.nsend) {
      LaunchScope ls(c, "reverse_f_unpack");
      k_reverse_f_unpack<<<nblk(a.nsend), BLOCK, 0, c->stream>>>(c->f.p, a.sendlist.p, a.nsend, s->recvbuf.p, s->fold_atomic);
    }
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// Replay of CommBrick::exchange's compaction loop on indices alone (host, no CUDA):
//   while (i < nlocal) { if (leaves(i)) { pack(i); copy(nlocal-1 -> i); nlocal--; } else i++; }
// `leavers` ascending.  order[k] = k-th atom packed; moves = (dst, src) pairs of the copies that survive;
// every leaver is packed exactly once, so order has nleave entries.
extern "C" int b200md_exchange_plan(int n, const int *leavers, int nleave, int *order, int *moves, int *nmoves,
                                    int *nlocal_out)
{
  if (n < 0 || nleave < 0 || nleave > n || (nleave && (!leavers || !order || !moves)) || !nmoves || !nlocal_out)
    return B200MD_ERR_ARG;
  std::unordered_set<int> L(leavers, leavers + nleave);
  int nl = n, no = 0, nm = 0, li = 0;
  while (li < nleave && leavers[li] < nl) {
    const int i = leavers[li++];
    order[no++] = i;
    for (;;) {
      nl--;
      if (nl == i) break;
      if (L.count(nl)) order[no++] = nl;
      else {
        moves[2 * nm] = i;
        moves[2 * nm + 1] = nl;
        nm++;
        break;
      }
    }
  }
  *nmoves = nm;
  *nlocal_out = nl;
  return no == nleave ? B200MD_OK : B200MD_ERR_ARG;
}

// ------------------------------------------------------------------ CommBrick::exchange
// Owned atoms that left the sub-box move to the neighbor rank, dimension by dimension.  The reference loop
// (comm_brick.cpp exchange(): "when atom is deleted, fill it in with last atom") fixes BOTH the order of the
// atoms that stay and the order in which leavers are packed; the order decides where migrated atoms land in
// the receiver's arrays and therefore neighbor-row order.  Only the (few) leaver indices go to the host,
// where that loop is replayed on indices alone; packing, hole filling and unpacking run on the device.
static int migrate(b200md_ctx *c, SystemState *s)
{
  const int *pg = s->d.procgrid;
  for (int dim = 0; dim < 3; dim++) {
    if (pg[dim] == 1) continue;    // Domain::pbc already wrapped this dimension: nobody leaves
    const int n = s->nlocal;
    const double lo = s->sublo[dim], hi = s->subhi[dim];
    std::vector<int> leavers;
    if (n) {
      CUDA_TRY(c, s->itmp.reserve((size_t) n + 64));
      CUDA_TRY(c, s->itmp2.reserve((size_t) n + 64));
      CUDA_TRY(c, s->scan64.reserve((size_t) n + 8));
      {
        LaunchScope ls(c, "leave_flags");
        k_leave_flags<<<nblk(n), BLOCK, 0, c->stream>>>(c->xq.p, n, dim, lo, hi, s->itmp.p);
      }
      int rc = b200md_exclusive_scan_i64(c, s->itmp.p, s->scan64.p, n, 1);
      if (rc) return rc;
      int64_t tot = 0;
      CUDA_TRY(c, cudaMemcpyAsync(&tot, s->scan64.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      if (tot) {
        LaunchScope ls(c, "scatter_list");
        k_scatter_list<<<nblk(n), BLOCK, 0, c->stream>>>(s->itmp.p, s->scan64.p, 0, n, s->itmp2.p);
        leavers.resize((size_t) tot);
        CUDA_TRY(c, cudaMemcpyAsync(leavers.data(), s->itmp2.p, tot * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      }
    }
    // replay of the reference loop on indices: emission order + (hole <- tail stayer) moves
    std::vector<int> order(leavers.size()), moves(2 * leavers.size() + 2);
    int nl = n, nmoves = 0;
    b200md_exchange_plan(n, leavers.data(), (int) leavers.size(), order.data(), moves.data(), &nmoves, &nl);
    moves.resize(2 * (size_t) nmoves);
    const int nsend = (int) order.size();
    CUDA_TRY(c, s->xbuf.reserve(8 * (size_t) nsend + 8));
    if (nsend) {
      CUDA_TRY(c, s->itmp.reserve((size_t) nsend + 2 * moves.size() + 64));
      CUDA_TRY(c, cudaMemcpyAsync(s->itmp.p, order.data(), nsend * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      {
        LaunchScope ls(c, "exchange_pack");
        k_exchange_pack<<<nblk(nsend), BLOCK, 0, c->stream>>>(c->xq.p, s->v.p, c->type.p, c->tag.p, s->itmp.p, nsend, s->xbuf.p);
      }
      if (!moves.empty()) {
        const int nm = (int) moves.size() / 2;
        CUDA_TRY(c, cudaMemcpyAsync(s->itmp.p + nsend, moves.data(), moves.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        LaunchScope ls(c, "exchange_fill");
        k_exchange_fill<<<nblk(nm), BLOCK, 0, c->stream>>>(c->xq.p, s->v.p, c->type.p, c->tag.p, s->itmp.p + nsend, nm);
      }
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));    // order/moves are host vectors
    }
    s->nlocal = nl;
    s->nmigrated += nsend;
    // counts, then atoms: first what the +dim neighbor sent, then (if more than 2 ranks) the -dim neighbor
    const int npass = pg[dim] > 2 ? 2 : 1;
    int nrecv[2] = {0, 0};
    double *dcnt = c->scal.p + 32;
    for (int pass = 0; pass < npass; pass++) {
      const int to = s->procneigh[dim][pass == 0 ? 0 : 1], from = s->procneigh[dim][pass == 0 ? 1 : 0];
      double cnt_h = (double) nsend, cnt_r = 0.0;
      CUDA_TRY(c, cudaMemcpyAsync(dcnt, &cnt_h, sizeof(double), cudaMemcpyHostToDevice, c->stream));
      int rc = xfer_sendrecv(c, s, dcnt, 1, to, dcnt + 1, 1, from);
      if (rc) return rc;
      CUDA_TRY(c, cudaMemcpyAsync(&cnt_r, dcnt + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      nrecv[pass] = (int) cnt_r;
    }
    const int nrtot = nrecv[0] + nrecv[1];
    CUDA_TRY(c, s->recvbuf.reserve(8 * (size_t) nrtot + 8));
    for (int pass = 0; pass < npass; pass++) {
      const int to = s->procneigh[dim][pass == 0 ? 0 : 1], from = s->procneigh[dim][pass == 0 ? 1 : 0];
      int rc = xfer_sendrecv(c, s, s->xbuf.p, 8 * (size_t) nsend, to, s->recvbuf.p + 8 * (size_t) (pass ? nrecv[0] : 0),
                             8 * (size_t) nrecv[pass], from);
      if (rc) return rc;
    }
    if (nrtot) {
      int rc = ensure_atoms(c, s, (size_t) s->nlocal + nrtot + 64);
      if (rc) return rc;
      {
        LaunchScope ls(c, "exchange_unpack");
        k_exchange_unpack<<<1, 32, 0, c->stream>>>(s->recvbuf.p, nrtot, dim, lo, hi, c->xq.p, s->v.p, c->type.p, c->tag.p,
                                                   s->nlocal, c->flags.p + 11);
      }
      int nnew = 0;
      CUDA_TRY(c, cudaMemcpyAsync(&nnew, c->flags.p + 11, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      s->nlocal = nnew;
    }
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// ------------------------------------------------------------------ halo overlap: interior / boundary centers
// An owned atom farther than R = max(rcLJmax) + skin from every face of the sub-domain has no ghost among the candidates
// of its inner rows: a candidate lies within rcLJmax + margin (margin <= skin/2) of the center when the rows are
// derived, and a ghost sits outside the sub-domain at the last master rebuild and has moved less than skin/2 since.
static void set_split_geometry(b200md_ctx *c, SystemState *s)
{
  SplitGeom &g = c->split;
  g.on = 0;
  // one rank: the self halos are two ~25 us kernels, less than the split costs (option overlap_halo = 2 forces it)
  if (s->d.style != 0 || !c->overlap_halo || c->deterministic || !c->lj_pairs) return;
  if (s->nranks == 1 && c->overlap_halo < 2) return;
  double rmax = 0.0;
  for (int k = 0; k < 4; k++) rmax = fmax(rmax, c->rp.rcLJmax[k]);
  const double R = rmax + s->d.skin;
  g.triclinic = s->triclinic;
  for (int d = 0; d < 3; d++) {
    g.boxlo[d] = s->boxlo[d];
    g.lo[d] = s->sublo[d];
    g.hi[d] = s->subhi[d];
    g.r[d] = R * s->cutghost[d] / s->cutneighmax;    // lamda units for a triclinic box (CommBrick::setup's scaling)
  }
  for (int k = 0; k < 6; k++) g.h_inv[k] = s->h_inv[k];
  g.on = 1;
}

// ------------------------------------------------------------------ reneighbor + forces
static int reneighbor(b200md_ctx *c, SystemState *s, bool first)
{
  const Geom geo = make_geom(s);
  int n = s->nlocal;
  if (s->triclinic && n) {
    LaunchScope ls(c, "x2lamda");
    k_x2lamda<<<nblk(n), BLOCK, 0, c->stream>>>(geo, c->xq.p, n);
  }
  if (n) {
    LaunchScope ls(c, "pbc");
    if (s->triclinic) k_pbc<<<nblk(n), BLOCK, 0, c->stream>>>(c->xq.p, n, 0, 0, 0, 1, 1, 1, 1, 1, 1);
    else
      k_pbc<<<nblk(n), BLOCK, 0, c->stream>>>(c->xq.p, n, s->boxlo[0], s->boxlo[1], s->boxlo[2], s->boxhi[0], s->boxhi[1],
                                             s->boxhi[2], s->prd[0], s->prd[1], s->prd[2]);
  }
  int rc;
  if (s->nranks > 1) {
    if ((rc = migrate(c, s))) return rc;
    n = s->nlocal;
  }
  if (s->d.sort_every > 0 && (first || s->step >= s->nextsort))
    if ((rc = sort_atoms(c, s))) return rc;
  if ((rc = halo_borders(c, s))) return rc;
  const int nall = s->nlocal + s->nghost;
  if (s->triclinic && nall) {
    LaunchScope ls(c, "lamda2x");
    k_lamda2x<<<nblk(nall), BLOCK, 0, c->stream>>>(geo, c->xq.p, nall);
  }
  c->nlocal = s->nlocal;
  c->nghost = s->nghost;
  c->nall = nall;
  // element code in w for the force kernels (ghosts included)
  {
    LaunchScope ls(c, "set_w");
    k_set_w<<<nblk(nall), BLOCK, 0, c->stream>>>(c->xq.p, c->type.p, c->map_d.p, s->d.style == 0 ? 1 : 0, nall);
  }
  // Neighbor::build: hold positions of owned atoms, then bins + rows
  CUDA_TRY(c, s->xhold.reserve((size_t) s->nlocal + 8));
  CUDA_TRY(c, cudaMemcpyAsync(s->xhold.p, c->xq.p, s->nlocal * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(c, s->xt.reserve((size_t) nall + 8));
  {
    LaunchScope ls(c, "make_xt");
    k_make_xt<<<nblk(nall), BLOCK, 0, c->stream>>>(c->xq.p, c->type.p, nall, s->xt.p);
  }
  b200md_box box = s->d.box;
  for (int d = 0; d < 3; d++) {
    box.sublo[d] = s->sublo[d];
    box.subhi[d] = s->subhi[d];
    box.cutghost[d] = s->cutghost[d];
  }
  box.cutneighmax = s->cutneighmax;
  if (first) c->list_stride = 0;    // a new system: the sticky stride of the one-pass builds starts over
  if ((rc = b200md_neigh_build_device(c, box, s->d.ntypes, s->cutneighsq.data(), s->cutneighghostsq.data(), s->nlocal,
                                      s->nghost, s->xt.p, s->ghost_rows, s->d.skin, !first && c->one_pass_neigh)))
    return rc;
  set_split_geometry(c, s);
  c->sys_owns_tight = true;
  rc = (s->d.style == 0) ? b200md_rebomos_build_inner(c) : b200md_aeam_build_inner(c);
  if (rc) return rc;
  if (s->d.style == 0 && (rc = b200md_rebomos_derive_tight(c))) return rc;
  s->ago = 0;
  if (!first) s->nbuild++;
  return B200MD_OK;
}

// the halo stream (highest priority: its small kernels take the SM slots that compute CTAs free) and its events
int b200md_ensure_halo_stream(b200md_ctx *c)
{
  if (c->halo_stream) return B200MD_OK;
  int lo = 0, hi = 0;
  CUDA_TRY(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUDA_TRY(c, cudaStreamCreateWithPriority(&c->halo_stream, cudaStreamNonBlocking, hi));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_fwd, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_reb, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_rev, cudaEventDisableTiming));
  return B200MD_OK;
}

static int compute_forces(b200md_ctx *c, SystemState *s, int eflag, int vflag)
{
  ARG_CHECK(c, !c->deterministic || c->nall <= 32 * (B200MD_DET_BLOCKS - 64),
            "deterministic mode holds per-block partial sums for at most 8.3 M atoms per GPU");
  const size_t n3 = 3 * (size_t) c->nall;
  CUDA_TRY(c, cudaMemsetAsync(c->f.p, 0, (n3 + 8) * sizeof(double), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), c->stream));
  int rc;
  if (s->d.style == 0) {
    if ((rc = b200md_rebomos_forces(c, eflag, vflag))) return rc;
  } else {
    if ((rc = b200md_aeam_density(c))) return rc;
    if ((rc = halo_forward_rho_fp(c, s))) return rc;
    // force-only steps on more than one rank: the reverse halo runs on its own stream beside the pair kernel (only the
    // angular kernel writes ghost forces)
    if (!eflag && !vflag && s->nranks > 1 && c->overlap_halo && !c->deterministic && !c->sync_timing && c->aeam_cluster == 2 &&
        !c->ap.asym_dr) {
      if ((rc = b200md_ensure_halo_stream(c))) return rc;
      cudaStream_t main = c->stream, halo = c->halo_stream;
      if ((rc = b200md_aeam_forces(c, 0, 0, 1))) return rc;
      CUDA_TRY(c, cudaEventRecord(c->ev_reb, main));
      CUDA_TRY(c, cudaStreamWaitEvent(halo, c->ev_reb, 0));
      c->stream = halo;
      s->fold_atomic = true;
      rc = halo_reverse_f(c, s);
      s->fold_atomic = false;
      c->stream = main;
      if (rc) return rc;
      CUDA_TRY(c, cudaEventRecord(c->ev_rev, halo));
      if ((rc = b200md_aeam_forces(c, 0, 0, 2))) return rc;
      CUDA_TRY(c, cudaStreamWaitEvent(main, c->ev_rev, 0));
      s->noverlap++;
      return B200MD_OK;
    }
    if ((rc = b200md_aeam_forces(c, eflag, vflag, 0))) return rc;
  }
  return halo_reverse_f(c, s);
}

// One force evaluation of the resident loop with the halos on their own stream (rebomos, split center lists):
//   halo stream   : forward x halo ...... | ................................ wait(bond order) reverse f halo (atomic folds)
//   compute stream: interior LJ ......... | wait(forward) bond order, all centers | boundary LJ ........ | wait(reverse)
// Interior centers read no ghost, so their LJ launches hide the forward halo; only the bond-order kernels write ghost
// forces (LJ is gather-form), so the reverse halo starts as soon as they are done and runs beside the boundary LJ
// launches -- its folds use atomics because those launches still add to the same owned atoms.  The bond-order launches
// stay whole and the center lists keep their ascending order; only the LJ pair ROWS are stored interior-first (pairs
// keep their two centers).  Reordering the center lists cost 0.06 ms per step in locality, selecting rows by a flag
// 0.10 ms in idle lanes -- both more than the halo they hid.
// derive: the tight rows are re-derived first (that needs the ghosts, so only the reverse halo is hidden).
static int forces_overlapped(b200md_ctx *c, SystemState *s, bool derive)
{
  int rc = b200md_ensure_halo_stream(c);
  if (rc) return rc;
  cudaStream_t main = c->stream, halo = c->halo_stream;
  // B200MD_OVERLAP_TRACE=1: device timestamps of one step's phases on both streams (diagnostic)
  static const bool trace_on = getenv("B200MD_OVERLAP_TRACE") != nullptr;
  const bool trace = trace_on && s->noverlap == 60;
  cudaEvent_t tv[8];
  if (trace)
    for (auto &e : tv) cudaEventCreate(&e);
  if (trace) cudaEventRecord(tv[0], main);
  CUDA_TRY(c, cudaEventRecord(c->ev_ready, main));
  CUDA_TRY(c, cudaStreamWaitEvent(halo, c->ev_ready, 0));
  c->stream = halo;
  rc = halo_forward_x(c, s);
  c->stream = main;
  if (rc) return rc;
  CUDA_TRY(c, cudaEventRecord(c->ev_fwd, halo));
  if (trace) cudaEventRecord(tv[1], halo);
  const size_t n3 = 3 * (size_t) c->nall;
  CUDA_TRY(c, cudaMemsetAsync(c->f.p, 0, (n3 + 8) * sizeof(double), main));
  CUDA_TRY(c, cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), main));
  if (derive) {
    CUDA_TRY(c, cudaStreamWaitEvent(main, c->ev_fwd, 0));
    if ((rc = b200md_rebomos_derive_tight(c))) return rc;
  }
  if ((rc = b200md_rebomos_forces_part(c, 0, 1))) return rc;
  if (trace) cudaEventRecord(tv[2], main);
  if (!derive) CUDA_TRY(c, cudaStreamWaitEvent(main, c->ev_fwd, 0));
  if ((rc = b200md_rebomos_forces_part(c, 0, 0))) return rc;
  CUDA_TRY(c, cudaEventRecord(c->ev_reb, main));
  if (trace) cudaEventRecord(tv[3], main);
  if ((rc = b200md_rebomos_forces_part(c, 1, 1))) return rc;
  if (trace) cudaEventRecord(tv[4], main);
  CUDA_TRY(c, cudaStreamWaitEvent(halo, c->ev_reb, 0));
  c->stream = halo;
  s->fold_atomic = true;
  rc = halo_reverse_f(c, s);
  s->fold_atomic = false;
  c->stream = main;
  if (rc) return rc;
  CUDA_TRY(c, cudaEventRecord(c->ev_rev, halo));
  if (trace) cudaEventRecord(tv[5], halo);
  CUDA_TRY(c, cudaStreamWaitEvent(main, c->ev_rev, 0));
  if (trace) {
    cudaEventRecord(tv[6], main);
    cudaEventSynchronize(tv[6]);
    float t[7];
    for (int k = 1; k <= 6; k++) cudaEventElapsedTime(&t[k], tv[0], tv[k]);
    fprintf(stderr, "[overlap trace rank %d] forward halo done %.3f | interior LJ done %.3f | bond order done %.3f | boundary LJ done "
                    "%.3f | reverse halo done %.3f | joined %.3f ms\n", s->me, t[1], t[2], t[3], t[4], t[5], t[6]);
    for (auto &e : tv) cudaEventDestroy(e);
  }
  s->noverlap++;
  return B200MD_OK;
}

// compute temp / pressure / pe (global sums over ranks)
static int thermo(b200md_ctx *c, SystemState *s)
{
  {
    LaunchScope ls(c, "ke");
    k_ke<<<c->num_sms * 2, BLOCK, 0, c->stream>>>(s->v.p, c->type.p, s->dmass.p, s->nlocal, b200md_scal_arg(c));
  }
  {
    int rc = b200md_det_fold(c);
    if (rc) return rc;
  }
  {
    int rc = xfer_allreduce_sum(c, s, c->scal.p, 9);
    if (rc) return rc;
  }
  double h[16];
  int fl[16];
  CUDA_TRY(c, cudaMemcpyAsync(h, c->scal.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(fl, c->flags.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (fl[4] || fl[5]) {    // inner-list statistics of the latest (re)build
    c->n_short_entries = fl[4];
    c->n_lj_entries = fl[5];
  }
  if (fl[0]) {
    c->fail("per-atom row overflow in the force kernels (flag " + std::to_string(fl[0]) + ")");
    return B200MD_ERR_OVERFLOW;
  }
  const double dof = fmax(3.0 * (double) s->natoms - 3.0, 0.0);
  const double tfactor = dof > 0.0 ? s->d.mvv2e / (dof * s->d.boltz) : 0.0;
  const double temp = h[8] * tfactor;
  const double vol = s->prd[0] * s->prd[1] * s->prd[2];
  const double press = (dof * s->d.boltz * temp + h[1] + h[2] + h[3]) / 3.0 / vol * s->d.nktv2p;
  std::vector<double> row(12);
  row[0] = (double) s->step;
  row[1] = temp;
  row[2] = press;
  row[3] = h[0];
  row[4] = temp * 0.5 * dof * s->d.boltz;
  row[5] = vol;
  for (int k = 0; k < 6; k++) row[6 + k] = h[1 + k];
  s->rows.push_back(row);
  return B200MD_OK;
}

// ================================================================== C ABI
void b200md_system_free(b200md_ctx *c)
{
  b200md_neigh_forget(c);
  b200md_aeam_forget(c);
  SystemState *s = c->sys;
  if (!s) return;
  for (auto &sw : s->swaps) sw.sendlist.release();
  s->v.release(); s->xhold.release(); s->xt.release(); s->dmass.release(); s->itmp.release(); s->itmp2.release();
  s->x4tmp.release(); s->dtmp.release(); s->scan64.release(); s->sendbuf.release(); s->recvbuf.release();
  s->xbuf.release();
  s->flat_src.release();
  s->flat_code.release();
  p2p_release(s);
  if (s->vote_host) cudaFreeHost(s->vote_host);
  if (s->nccl) ncclCommDestroy(s->nccl);
  delete s;
  c->sys = nullptr;
}

extern "C" int b200md_nccl_unique_id(void *id128)
{
  if (!id128) return B200MD_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return B200MD_ERR_NCCL;
  memcpy(id128, &id, 128);
  return B200MD_OK;
}

extern "C" int b200md_system_create(b200md_ctx *c, const b200md_system_desc *d, int nlocal, const double *x,
                                    const double *v, const int *type, const int *tag)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, d && nlocal >= 0 && (nlocal == 0 || (x && type && tag)), "system_create: NULL arguments");
  ARG_CHECK(c, d->ntypes >= 1 && d->ntypes <= B200MD_MAX_TYPES && d->mass, "system_create: bad ntypes/mass");
  ARG_CHECK(c, d->procgrid[0] >= 1 && d->procgrid[1] >= 1 && d->procgrid[2] >= 1, "system_create: bad procgrid");
  ARG_CHECK(c, d->rank >= 0 && d->rank < d->procgrid[0] * d->procgrid[1] * d->procgrid[2], "system_create: bad rank");
  ARG_CHECK(c, d->style == 0 || d->style == 1, "system_create: style must be 0 (rebomos) or 1 (aeam)");
  CUDA_TRY(c, cudaSetDevice(c->device));
  ncclComm_t keep = nullptr;
  std::shared_ptr<LocalGroup> keep_local;
  if (c->sys) {
    keep = c->sys->nccl;
    keep_local = c->sys->local;
    c->sys->nccl = nullptr;
    SystemState *old = c->sys;
    for (auto &sw : old->swaps) sw.sendlist.release();
    old->v.release(); old->xhold.release(); old->xt.release(); old->dmass.release(); old->itmp.release();
    old->itmp2.release(); old->x4tmp.release(); old->dtmp.release(); old->scan64.release(); old->sendbuf.release();
    old->recvbuf.release();
    old->xbuf.release();
    old->flat_src.release();
    old->flat_code.release();
    p2p_release(old);
    if (old->vote_host) cudaFreeHost(old->vote_host);
    delete old;
    c->sys = nullptr;
  }
  SystemState *s = new SystemState();
  c->sys = s;
  if (cudaHostAlloc((void **) &s->vote_host, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) *s->vote_host = 0;
  else {
    s->vote_host = nullptr;    // the vote then goes through the copy + synchronise path
    cudaGetLastError();
  }
  s->nccl = keep;
  s->local = keep_local;
  s->d = *d;
  s->mass.assign(d->mass, d->mass + d->ntypes + 1);
  s->d.mass = s->mass.data();
  int rc = setup_cutoffs(c, s);
  if (rc) return rc;
  setup_geometry(s);
  setup_swaps(s);
  ARG_CHECK(c, s->nranks == 1 || s->nccl || s->local,
            "system_create: call b200md_system_comm_init (or _comm_init_local) before creating a multi-rank system");
  ARG_CHECK(c, !s->local || s->local->nranks == s->nranks, "system_create: procgrid does not match the loopback group size");
  s->nlocal = nlocal;
  if ((rc = ensure_atoms(c, s, (size_t) nlocal + nlocal / 2 + 1024))) return rc;
  CUDA_TRY(c, s->dmass.reserve(B200MD_MAX_TYPES + 2));
  CUDA_TRY(c, cudaMemcpyAsync(s->dmass.p, s->mass.data(), (d->ntypes + 1) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (nlocal) {
    CUDA_TRY(c, c->x_aos.reserve(3 * (size_t) nlocal + 8));
    CUDA_TRY(c, cudaMemcpyAsync(c->x_aos.p, x, 3 * (size_t) nlocal * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    {
      LaunchScope ls(c, "upload_x");
      k_upload_x<<<nblk(nlocal), BLOCK, 0, c->stream>>>(c->x_aos.p, nlocal, c->xq.p);
    }
    if (v) CUDA_TRY(c, cudaMemcpyAsync(s->v.p, v, 3 * (size_t) nlocal * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    else CUDA_TRY(c, cudaMemsetAsync(s->v.p, 0, 3 * (size_t) nlocal * sizeof(double), c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->type.p, type, nlocal * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->tag.p, tag, nlocal * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  // global atom count
  s->natoms = nlocal;
  if (s->nranks > 1) {
    double cnt = (double) nlocal;
    CUDA_TRY(c, cudaMemcpyAsync(c->scal.p + 40, &cnt, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = xfer_allreduce_sum(c, s, c->scal.p + 40, 1))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(&cnt, c->scal.p + 40, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    s->natoms = (long long) cnt;
  }
  // Verlet::setup: ghosts, lists, forces with energy + virial, thermo row 0
  s->step = 0;
  s->nbuild = s->ndanger = 0;
  if ((rc = reneighbor(c, s, true))) return rc;
  if ((rc = compute_forces(c, s, 1, 2))) return rc;
  if ((rc = thermo(c, s))) return rc;
  return B200MD_OK;
}

extern "C" int b200md_system_comm_init(b200md_ctx *c, const void *id128, int nranks, int rank)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, id128 && nranks >= 1 && rank >= 0 && rank < nranks, "system_comm_init: bad arguments");
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (!c->sys) c->sys = new SystemState();
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NCCL_TRY(c, ncclCommInitRank(&c->sys->nccl, nranks, id, rank));
  return B200MD_OK;
}

extern "C" int b200md_local_group_create(int nranks)
{
  if (nranks < 1) return B200MD_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_groups_mu);
  auto g = std::make_shared<LocalGroup>();
  g->nranks = nranks;
  g->slot.resize((size_t) nranks * nranks);
  const int id = g_next_group++;
  g_groups[id] = g;
  return id;
}

extern "C" int b200md_system_comm_init_local(b200md_ctx *c, int group, int nranks, int rank)
{
  if (!c) return B200MD_ERR_ARG;
  std::shared_ptr<LocalGroup> g;
  {
    std::lock_guard<std::mutex> lk(g_groups_mu);
    auto it = g_groups.find(group);
    if (it != g_groups.end()) g = it->second;
  }
  ARG_CHECK(c, g && g->nranks == nranks && rank >= 0 && rank < nranks, "system_comm_init_local: unknown group or bad rank");
  if (!c->sys) c->sys = new SystemState();
  c->sys->local = g;
  return B200MD_OK;
}

// ---- fix nvt on the device loop.  The chain variables live on the host (a handful of doubles); per step the loop
// needs the kinetic energy once (one reduction kernel, read back with the step's other host sync) and scales the
// velocities twice (FixNH::initial_integrate / final_integrate, [LAMMPS-core] src/fix_nh.cpp).
static int nh_temperature(b200md_ctx *c, SystemState *s, double *t_out)
{
  CUDA_TRY(c, cudaMemsetAsync(c->scal.p + 10, 0, sizeof(double), c->stream));
  {
    LaunchScope ls(c, "ke");
    k_ke<<<c->num_sms * 2, BLOCK, 0, c->stream>>>(s->v.p, c->type.p, s->dmass.p, s->nlocal, b200md_scal_arg(c) + 2);    // -> scal[10]
  }
  {
    int rcf = b200md_det_fold(c);
    if (rcf) return rcf;
  }
  int rc = xfer_allreduce_sum(c, s, c->scal.p + 10, 1);
  if (rc) return rc;
  double h = 0.0;
  CUDA_TRY(c, cudaMemcpyAsync(&h, c->scal.p + 10, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  const double dof = fmax(3.0 * (double) s->natoms - 3.0, 0.0);
  *t_out = dof > 0.0 ? h * s->d.mvv2e / (dof * s->d.boltz) : 0.0;
  return B200MD_OK;
}

static void nh_temp_target(SystemState *s)
{
  SystemState::NoseHoover &n = s->nh;
  double delta = (double) (s->step - n.begin);
  if (delta != 0.0) delta /= (double) (n.end - n.begin);
  n.t_target = n.t_start + delta * (n.t_stop - n.t_start);
  n.ke_target = n.tdof * s->d.boltz * n.t_target;
}

// FixNH::nhc_temp_integrate (nc_tchain = 1, tdrag_factor = 1, eta_mass_flag = 1); returns the velocity factor
static double nh_chain_half_step(SystemState *s)
{
  SystemState::NoseHoover &n = s->nh;
  const double boltz = s->d.boltz, dt = s->d.dt;
  const double dthalf = 0.5 * dt, dt4 = 0.25 * dt, dt8 = 0.125 * dt;
  double expfac;
  double kecurrent = n.tdof * boltz * n.t_current;
  n.eta_mass[0] = n.tdof * boltz * n.t_target / (n.t_freq * n.t_freq);
  for (int ich = 1; ich < 3; ich++) n.eta_mass[ich] = boltz * n.t_target / (n.t_freq * n.t_freq);
  n.eta_dotdot[0] = n.eta_mass[0] > 0.0 ? (kecurrent - n.ke_target) / n.eta_mass[0] : 0.0;
  for (int ich = 2; ich > 0; ich--) {
    expfac = exp(-dt8 * n.eta_dot[ich + 1]);
    n.eta_dot[ich] *= expfac;
    n.eta_dot[ich] += n.eta_dotdot[ich] * dt4;
    n.eta_dot[ich] *= expfac;
  }
  expfac = exp(-dt8 * n.eta_dot[1]);
  n.eta_dot[0] *= expfac;
  n.eta_dot[0] += n.eta_dotdot[0] * dt4;
  n.eta_dot[0] *= expfac;
  const double factor = exp(-dthalf * n.eta_dot[0]);
  n.t_current *= factor * factor;
  kecurrent = n.tdof * boltz * n.t_current;
  n.eta_dotdot[0] = n.eta_mass[0] > 0.0 ? (kecurrent - n.ke_target) / n.eta_mass[0] : 0.0;
  for (int ich = 0; ich < 3; ich++) n.eta[ich] += dthalf * n.eta_dot[ich];
  n.eta_dot[0] *= expfac;
  n.eta_dot[0] += n.eta_dotdot[0] * dt4;
  n.eta_dot[0] *= expfac;
  for (int ich = 1; ich < 3; ich++) {
    expfac = exp(-dt8 * n.eta_dot[ich + 1]);
    n.eta_dot[ich] *= expfac;
    n.eta_dotdot[ich] = (n.eta_mass[ich - 1] * n.eta_dot[ich - 1] * n.eta_dot[ich - 1] - boltz * n.t_target) / n.eta_mass[ich];
    n.eta_dot[ich] += n.eta_dotdot[ich] * dt4;
    n.eta_dot[ich] *= expfac;
  }
  return factor;
}

static int nh_scale_velocities(b200md_ctx *c, SystemState *s, double factor)
{
  const size_t n3 = 3 * (size_t) s->nlocal;
  if (n3) {
    LaunchScope ls(c, "nh_scale_v");
    k_scale_v<<<(unsigned) ((n3 + BLOCK - 1) / BLOCK), BLOCK, 0, c->stream>>>(s->v.p, n3, factor);
  }
  return B200MD_OK;
}

extern "C" int b200md_system_set_nvt(b200md_ctx *c, double t_start, double t_stop, double t_period)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->sys, "system_set_nvt: call b200md_system_create first");
  SystemState::NoseHoover &n = c->sys->nh;
  if (t_period <= 0.0) {    // back to plain NVE
    n.on = false;
    return B200MD_OK;
  }
  ARG_CHECK(c, t_start > 0.0 && t_stop > 0.0, "Target temperature for fix nvt cannot be 0.0");
  n = SystemState::NoseHoover();
  n.on = true;
  n.t_start = t_start;
  n.t_stop = t_stop;
  n.t_period = t_period;
  return B200MD_OK;
}

// thermostat part of the conserved quantity (FixNH::compute_scalar for a pure thermostat)
extern "C" double b200md_system_nh_energy(b200md_ctx *c)
{
  if (!c || !c->sys || !c->sys->nh.on) return 0.0;
  const SystemState::NoseHoover &n = c->sys->nh;
  const double kt = c->sys->d.boltz * n.t_target;
  double e = n.ke_target * n.eta[0] + 0.5 * n.eta_mass[0] * n.eta_dot[0] * n.eta_dot[0];
  for (int ich = 1; ich < 3; ich++) e += kt * n.eta[ich] + 0.5 * n.eta_mass[ich] * n.eta_dot[ich] * n.eta_dot[ich];
  return e;
}

extern "C" int b200md_system_run(b200md_ctx *c, int nsteps, int thermo_every)
{
  if (!c) return B200MD_ERR_ARG;
  SystemState *s = c->sys;
  ARG_CHECK(c, s && s->swaps.size(), "system_run: call b200md_system_create first");
  ARG_CHECK(c, nsteps >= 0 && thermo_every >= 0, "system_run: bad arguments");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const double dtv = s->d.dt, dtf = 0.5 * s->d.dt * s->d.ftm2v;
  const double triggersq = 0.25 * s->d.skin * s->d.skin;
  const long long last = s->step + nsteps;
  int rc;
  if (s->nh.on && nsteps > 0) {    // FixNH::setup at the start of every run
    SystemState::NoseHoover &n = s->nh;
    n.begin = s->step;
    n.end = last;
    n.t_freq = 1.0 / n.t_period;
    n.tdof = fmax(3.0 * (double) s->natoms - 3.0, 0.0);
    if ((rc = nh_temperature(c, s, &n.t_current))) return rc;
    nh_temp_target(s);
    n.eta_mass[0] = n.tdof * s->d.boltz * n.t_target / (n.t_freq * n.t_freq);
    for (int ich = 1; ich < 3; ich++) n.eta_mass[ich] = s->d.boltz * n.t_target / (n.t_freq * n.t_freq);
    for (int ich = 1; ich < 3; ich++)
      n.eta_dotdot[ich] = (n.eta_mass[ich - 1] * n.eta_dot[ich - 1] * n.eta_dot[ich - 1] - s->d.boltz * n.t_target) / n.eta_mass[ich];
  }
  bool final_pending = false;
  // B200MD_STEP_TRACE=1: device timestamps at the start of a step, after the reneighbor vote has reached the host, and at
  // the end of the step, for steps 60-67 of a run (diagnostic: what the vote costs, what the halos leave exposed)
  static const bool step_trace_on = getenv("B200MD_STEP_TRACE") != nullptr;
  static const bool step_trace_slow = step_trace_on && !strcmp(getenv("B200MD_STEP_TRACE"), "slow");    // every step; slow ones printed
  std::vector<cudaEvent_t> tev;
  std::vector<int> tflag;
  for (int it = 0; it < nsteps; it++) {
    s->step++;
    const bool tr = step_trace_on && (step_trace_slow || (it >= 60 && it < 68));
    if (tr) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->stream);
      tev.push_back(e);
    }
    if (s->nh.on) {    // FixNH::initial_integrate: thermostat half step before the kick
      nh_temp_target(s);
      if ((rc = nh_scale_velocities(c, s, nh_chain_half_step(s)))) return rc;
    }
    const bool thermo_step = (s->step == last) || (thermo_every > 0 && s->step % thermo_every == 0);
    const int n = s->nlocal;
    if (n) {
      LaunchScope ls(c, "initial_integrate");
      const bool two_level = c->inner_valid && c->margin < s->d.skin;
      const double innersq = 0.25 * c->margin * c->margin;
      const bool three_level = c->inner_valid && c->tight_valid;
      const double tightsq = 0.25 * c->margin_t * c->margin_t;
      k_initial_integrate<<<nblk(n), BLOCK, 0, c->stream>>>(c->xq.p, s->v.p, c->f.p, c->type.p, s->dmass.p, n, dtf, dtv,
                                                          s->xhold.p, triggersq,
                                                          two_level ? (const double4 *) c->xhold.p : nullptr, innersq,
                                                          three_level ? (const double4 *) c->xhold_t.p : nullptr, tightsq,
                                                          c->flags.p, final_pending ? 1 : 0);
      final_pending = false;
    }
    // Neighbor::decide (every 1, delay 0, check yes): rebuild if any owned atom moved more than skin/2
    s->ago++;
    int flag = 0;
    s->vote_epoch++;
    const bool peer_vote = c->peer_vote && s->vote_host && !c->sync_timing &&
        (s->nranks == 1 || (s->p2p.ok && s->p2p.vote_peers.p && !s->local));
    if (peer_vote) {
      {
        LaunchScope ls(c, "vote");
        k_vote<<<1, VOTE_MAXR, 0, c->stream>>>(c->flags.p + 9, s->nranks > 1 ? s->p2p.vote_peers.p : nullptr, s->me, s->nranks,
                                              s->vote_epoch, s->vote_host);
      }
      volatile int *hw = s->vote_host;
      int v;
      long long spins = 0;
      const auto t0 = std::chrono::steady_clock::now();
      while (((v = *hw) >> 3) != s->vote_epoch) {
        if ((++spins & 0xfffff) == 0) {    // now and then: has the stream died, or has a peer stopped?
          const cudaError_t q = cudaStreamQuery(c->stream);
          if (q != cudaSuccess && q != cudaErrorNotReady) CUDA_TRY(c, q);
          if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0) {
            c->fail("reneighbor vote: no report from the device within 60 s");
            return B200MD_ERR_NCCL;
          }
        }
      }
      flag = v & 7;
      if (flag == 7) {
        c->fail("reneighbor vote: a peer rank's vote did not arrive (peer stopped?)");
        return B200MD_ERR_NCCL;
      }
    } else {
      if ((rc = xfer_allreduce_max_int(c, s, c->flags.p + 9))) return rc;
      CUDA_TRY(c, cudaMemcpyAsync(&flag, c->flags.p + 9, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      if (flag) CUDA_TRY(c, cudaMemsetAsync(c->flags.p + 9, 0, sizeof(int), c->stream));
    }
    if (tr) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->stream);
      tev.push_back(e);
      tflag.push_back(flag);
    }
    if (c->force_rebuild) {    // option "force_rebuild": this step takes the reneighboring path (measurement; collective)
      c->force_rebuild = 0;
      flag = 3;
    }
    bool forces_done = false;
    if (flag >= 3) {
      if (s->ago == 1) s->ndanger++;
      if ((rc = reneighbor(c, s, false))) return rc;
    } else if (flag <= 1 && !thermo_step && s->d.style == 0 && c->split_valid && !c->sync_timing && c->overlap_halo && !c->deterministic) {
      if ((rc = forces_overlapped(c, s, flag == 1))) return rc;
      forces_done = true;
    } else {
      if ((rc = halo_forward_x(c, s))) return rc;
      if (flag == 2) {    // master list still valid: re-derive the inner lists from it at the current positions
        rc = (s->d.style == 0) ? b200md_rebomos_build_inner(c) : b200md_aeam_build_inner(c);
        if (rc) return rc;
        if (s->d.style == 0 && (rc = b200md_rebomos_derive_tight(c))) return rc;
        s->ninner++;
      } else if (flag == 1 && s->d.style == 0) {    // inner lists still valid: only the tight rows are re-derived
        if ((rc = b200md_rebomos_derive_tight(c))) return rc;
      }
    }
    if (!forces_done && (rc = compute_forces(c, s, thermo_step ? 1 : 0, thermo_step ? 2 : 0))) return rc;
    // the second half kick moves into the next step's initial_integrate launch when nothing reads v in between
    if (s->nlocal && it + 1 < nsteps && !thermo_step && !s->nh.on && c->fuse_integrate) final_pending = true;
    else if (s->nlocal) {    // not `n`: migration at a reneighboring step changes the owned count
      LaunchScope ls(c, "final_integrate");
      k_final_integrate<<<nblk(s->nlocal), BLOCK, 0, c->stream>>>(s->v.p, c->f.p, c->type.p, s->dmass.p, s->nlocal, dtf);
    }
    if (s->nh.on) {    // FixNH::final_integrate: temperature of the kicked velocities, thermostat half step
      if ((rc = nh_temperature(c, s, &s->nh.t_current))) return rc;
      if ((rc = nh_scale_velocities(c, s, nh_chain_half_step(s)))) return rc;
    }
    if (thermo_step)
      if ((rc = thermo(c, s))) return rc;
  }
  if (!tev.empty()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    tev.push_back(e);
    cudaStreamSynchronize(c->stream);
    std::vector<float> dur;
    for (size_t k = 0; k + 2 < tev.size(); k += 2) {
      float b = 0;
      cudaEventElapsedTime(&b, tev[k], tev[k + 2]);
      dur.push_back(b);
    }
    std::vector<float> sorted = dur;
    std::sort(sorted.begin(), sorted.end());
    const float med = sorted.empty() ? 0.f : sorted[sorted.size() / 2];
    if (step_trace_slow) fprintf(stderr, "[step trace rank %d] %zu steps, median %.3f ms\n", s->me, dur.size(), med);
    for (size_t k = 0; k + 2 < tev.size(); k += 2) {
      float a = 0;
      cudaEventElapsedTime(&a, tev[k], tev[k + 1]);
      if (!step_trace_slow || dur[k / 2] > 1.5f * med)
        fprintf(stderr, "[step trace rank %d] step %zu (vote %d): integrate + vote + host %.3f ms | step %.3f ms\n", s->me,
                k / 2, tflag[k / 2], a, dur[k / 2]);
    }
    for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaGetLastError());
  b200md_collect_timers(c);
  return B200MD_OK;
}

extern "C" int b200md_system_thermo_count(b200md_ctx *c)
{
  if (!c || !c->sys) return 0;
  return (int) c->sys->rows.size();
}
extern "C" int b200md_system_thermo_row(b200md_ctx *c, int i, double *out)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->sys && out && i >= 0 && i < (int) c->sys->rows.size(), "system_thermo_row: bad index");
  memcpy(out, c->sys->rows[i].data(), 12 * sizeof(double));
  return B200MD_OK;
}
extern "C" int b200md_system_thermo(b200md_ctx *c, double *out)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->sys && c->sys->rows.size(), "system_thermo: nothing evaluated yet");
  return b200md_system_thermo_row(c, (int) c->sys->rows.size() - 1, out);
}
extern "C" int b200md_system_sizes(b200md_ctx *c, long long *out)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->sys && out, "system_sizes: no system");
  out[0] = c->sys->nlocal;
  out[1] = c->sys->nghost;
  out[2] = c->sys->nbuild;
  out[3] = c->sys->ndanger;
  out[4] = c->sys->nmigrated;
  out[5] = c->sys->natoms;
  out[6] = c->sys->ninner;
  out[7] = c->sys->noverlap;
  return B200MD_OK;
}
extern "C" int b200md_system_download(b200md_ctx *c, double *x, double *v, double *f, int *type, int *tag)
{
  if (!c) return B200MD_ERR_ARG;
  SystemState *s = c->sys;
  ARG_CHECK(c, s, "system_download: no system");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int nall = s->nlocal + s->nghost;
  if (x && nall) {
    CUDA_TRY(c, c->x_aos.reserve(3 * (size_t) nall + 8));
    {
      LaunchScope ls(c, "download_x");
      k_download_x<<<nblk(nall), BLOCK, 0, c->stream>>>(c->xq.p, nall, c->x_aos.p);
    }
    CUDA_TRY(c, cudaMemcpyAsync(x, c->x_aos.p, 3 * (size_t) nall * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (v && s->nlocal) CUDA_TRY(c, cudaMemcpyAsync(v, s->v.p, 3 * (size_t) s->nlocal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (f && nall) CUDA_TRY(c, cudaMemcpyAsync(f, c->f.p, 3 * (size_t) nall * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (type && nall) CUDA_TRY(c, cudaMemcpyAsync(type, c->type.p, nall * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (tag && nall) CUDA_TRY(c, cudaMemcpyAsync(tag, c->tag.p, nall * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200MD_OK;
}
