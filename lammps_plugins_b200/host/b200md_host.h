/* -*- c++ -*- ----------------------------------------------------------
   Shared glue between the LAMMPS-facing pair classes and the b200md C ABI:
   one device context per Pair instance, neighbor-list hand-over (device build
   from LAMMPS' box/cutoffs, or upload of LAMMPS' own list), error mapping
   (status != 0  ->  error->one with the library's message; there is no CPU
   fallback to fall back to).
------------------------------------------------------------------------- */
#ifndef B200MD_HOST_H
#define B200MD_HOST_H

#include "b200md.h"

#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "neigh_list.h"
#include "neighbor.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace B200MDHost {

// which GPU this rank drives: B200MD_DEVICE if set, else the NODE-LOCAL rank (as the MPI launcher exports it; the
// global rank if it exports nothing) modulo the number of visible sm_100 devices (B200MD_DEVICES_PER_NODE overrides
// the count).  Returns -1 when no usable device is visible: b200md_create then reports why.
inline int pick_device(int me)
{
  const char *env = getenv("B200MD_DEVICE");
  if (env) return atoi(env);
  int local = me;
  static const char *const local_rank_vars[] = {"OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID",
                                                "PMI_LOCAL_RANK", "SLURM_LOCALID", "LOCAL_RANK"};
  for (const char *v : local_rank_vars) {
    const char *e = getenv(v);
    if (e && *e) {
      local = atoi(e);
      break;
    }
  }
  const char *vis = getenv("B200MD_DEVICES_PER_NODE");
  int per = vis ? atoi(vis) : b200md_device_count();
  if (per <= 0) return vis ? 0 : -1;
  return local % per;
}

// B200MD_STATS_FILE=<path>: when the pair style is destroyed, append one JSON line with the library's counters
// (what the benchmark's drop-in leg reads: pipelined calls / recomputes, list hand-overs, bytes moved)
inline void write_stats(b200md_ctx *ctx, const char *style, int me)
{
  const char *path = getenv("B200MD_STATS_FILE");
  if (!ctx || !path || !*path) return;
  FILE *fp = fopen(path, "a");
  if (!fp) return;
  static const char *const names[] = {"kernel_launches", "list_uploads", "inner_rebuilds", "h2d_bytes", "d2h_bytes",
                                      "pipelined_calls", "pipelined_redos", "compute_calls"};
  fprintf(fp, "{\"style\": \"%s\", \"rank\": %d", style, me);
  for (const char *n : names) fprintf(fp, ", \"%s\": %lld", n, b200md_get_counter(ctx, n));
  fprintf(fp, "}\n");
  fclose(fp);
}

// true -> rebuild the list on the device from LAMMPS' positions (default);
// false -> upload LAMMPS' own NeighList (B200MD_NEIGH=host)
inline bool device_neighbor_build()
{
  const char *env = getenv("B200MD_NEIGH");
  return !(env && strcmp(env, "host") == 0);
}

template <class ErrorT> inline void check(ErrorT *error, b200md_ctx *ctx, int rc, const char *what)
{
  if (rc != B200MD_OK)
    error->one(__FILE__, __LINE__, "B200 {} failed ({}): {}", what, rc, b200md_last_error(ctx));
}

// hand the current neighbor list to the device (called when LAMMPS has just rebuilt it: neighbor->ago == 0)
inline int sync_neighbor_list(b200md_ctx *ctx, LAMMPS_NS::Atom *atom, LAMMPS_NS::Neighbor *neighbor,
                              LAMMPS_NS::Comm *comm, LAMMPS_NS::Domain *domain, LAMMPS_NS::NeighList *list,
                              int ghost_rows)
{
  using namespace LAMMPS_NS;
  if (device_neighbor_build()) {
    b200md_box box;
    memset(&box, 0, sizeof(box));
    box.triclinic = domain->triclinic;
    for (int d = 0; d < 3; d++) {
      box.boxlo[d] = domain->boxlo[d];
      box.boxhi[d] = domain->boxhi[d];
      box.sublo[d] = domain->triclinic ? domain->sublo_lamda[d] : domain->sublo[d];
      box.subhi[d] = domain->triclinic ? domain->subhi_lamda[d] : domain->subhi[d];
      box.cutghost[d] = comm->cutghost[d];
    }
    box.xy = domain->xy;
    box.xz = domain->xz;
    box.yz = domain->yz;
    box.cutneighmax = neighbor->cutneighmax;
    return b200md_neigh_build(ctx, &box, atom->ntypes, &neighbor->cutneighsq[0][0],
                              &neighbor->cutneighghostsq[0][0], atom->nlocal, atom->nghost, &atom->x[0][0],
                              atom->type, ghost_rows, neighbor->skin);
  }
  // LAMMPS indexes numneigh[] / firstneigh[] by ATOM index and names the atoms that have a row in ilist[]
  return b200md_set_neighbor_list_ilist(ctx, list->inum, ghost_rows ? list->gnum : 0, list->ilist, list->numneigh,
                                        list->firstneigh, neighbor->skin);
}

// Is atom->f known to be all zero when Pair::compute() is entered?  Verlet::force_clear() zeroes it and only fixes with
// a pre_force() method run between that and the pair style; a sub-style of pair hybrid may find the forces of the
// styles before it.  When it is zero the library WRITES the forces (DMA straight into atom->f) instead of adding them
// on the host -- at 1 M atoms the host-side add (96 MB of memory traffic on one core) costs several times the GPU
// step.  B200MD_F_OVERWRITE=0 forces the accumulate path.
template <class PairT, class ForceT, class ModifyT> inline bool forces_zero_on_entry(PairT *self, ForceT *force, ModifyT *modify)
{
  const char *env = getenv("B200MD_F_OVERWRITE");
  if (env && atoi(env) == 0) return false;
  if ((void *) force->pair != (void *) self) return false;    // pair hybrid and friends
#ifdef LMPSHIM_H
  (void) modify;    // the API shim's host application has no pre_force fixes (its Modify type is opaque here)
  return true;
#else
  return modify->n_pre_force == 0;
#endif
}

// Page-locks LAMMPS' per-atom x and f blocks so that the library's piecewise upload / ranged download run as DMA beside
// its kernels.  OPT-IN (B200MD_PIN_HOST=1): LAMMPS reallocates these blocks when nmax grows (Atom::avec->grow, at a
// reneighboring step, i.e. between two compute() calls), and CUDA wants a block unregistered BEFORE it is freed -- the
// pair style only sees the new pointer afterwards.  The registration is refreshed whenever a base pointer or nmax
// changes, which is what current drivers tolerate, but that order could not be exercised against a real LAMMPS here.
// Without pinning everything works the same, the driver stages the transfers.
struct PinnedAtomArrays {
  void *seen_x = nullptr, *seen_f = nullptr;    // the blocks last looked at ...
  int seen_nmax = 0;
  void *reg_x = nullptr, *reg_f = nullptr;      // ... and what is actually registered (a refusal is not retried)
  void release()
  {
    if (reg_x) b200md_host_unregister(reg_x);
    if (reg_f) b200md_host_unregister(reg_f);
    reg_x = reg_f = seen_x = seen_f = nullptr;
    seen_nmax = 0;
  }
  void refresh(LAMMPS_NS::Atom *atom)
  {
    const char *env = getenv("B200MD_PIN_HOST");
    if (!env || atoi(env) == 0) return;
    if (atom->nmax <= 0 || !atom->x || !atom->f) return;
    void *px = &atom->x[0][0], *pf = &atom->f[0][0];
    if (px == seen_x && pf == seen_f && atom->nmax == seen_nmax) return;
    release();
    const size_t bytes = 3 * (size_t) atom->nmax * sizeof(double);
    if (b200md_host_register(px, bytes) == B200MD_OK) reg_x = px;
    if (b200md_host_register(pf, bytes) == B200MD_OK) reg_f = pf;
    seen_x = px;
    seen_f = pf;
    seen_nmax = atom->nmax;
  }
  ~PinnedAtomArrays() { release(); }
};

}    // namespace B200MDHost
#endif
