/* -*- c++ -*- ----------------------------------------------------------
   pair_style aeam -- B200-native implementation.

   Same style name, pair_coeff grammar and potential-file format as the
   reference USER-AEAM package (lammps/lammps-plugins USER-AEAM/pair_aeam.h:27-83,
   pair_aeam.cpp:513-746); density, embedding and force passes run on the GPU
   through the b200md C ABI.  The fp forward exchange declared by the reference
   (comm_forward = 1) is what carries F'(rho) to ghost atoms between the two
   device phases.
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS
// clang-format off
PairStyle(aeam,PairAEAM);
// clang-format on
#else

#ifndef LMP_PAIR_AEAM_B200_H
#define LMP_PAIR_AEAM_B200_H

#include "b200md.h"
#include "pair.h"
#include "b200md_host.h"

#include <string>
#include <vector>

namespace LAMMPS_NS {

class PairAEAM : public Pair {
 public:
  explicit PairAEAM(class LAMMPS *lmp);
  ~PairAEAM() override;

  // ---- what LAMMPS drives: the virtuals the reference overrides (USER-AEAM/pair_aeam.h:30-41), same meaning
  void settings(int narg, char **arg) override;      // pair_style aeam   (takes no arguments)
  void coeff(int narg, char **arg) override;         // pair_coeff * * <setfl file> <element per atom type>
  void init_style() override;                        // newton on, full list; tables go to the device here
  double init_one(int itype, int jtype) override;    // largest cutoff of the file
  void compute(int eflag, int vflag) override;       // density phase | fp halo by LAMMPS' Comm | force phase
  double memory_usage() override;

  // F'(rho) of ghost atoms between the two device phases (forward); rho (reverse) kept for interface parity
  int pack_forward_comm(int n, int *list, double *buf, int pbc_flag, int *pbc) override;
  void unpack_forward_comm(int n, int first, double *buf) override;
  int pack_reverse_comm(int n, int first, double *buf) override;
  void unpack_reverse_comm(int n, int *list, double *buf) override;

 protected:
  int nmax;             // allocated size of per-atom arrays
  double *rho, *fp;     // host mirrors: rho[i] of owned atoms, fp[] of owned + ghost atoms

  // the potential file, as read
  struct Setfl {
    int nelements = 0, nnonangular = 0, nangular = 0;
    std::vector<std::string> elements;
    std::vector<int> nrho;
    std::vector<double> drho, mass;
    std::vector<int> nr;              // [nelements*nelements]
    std::vector<double> dr, cut;      // [nelements*nelements]
    std::vector<std::vector<double>> frho, rhor, z2r;
  };
  Setfl *setfl;

  b200md_ctx *ctx;
  int uploaded_nlocal, uploaded_nghost;
  int f_overwrite = -1;                   // what the library was last told (b200md_set_option "f_overwrite")
  B200MDHost::PinnedAtomArrays pinned;    // atom->x / atom->f page-locked for DMA beside the kernels

  void allocate();
  virtual void read_file(char *);
};
}    // namespace LAMMPS_NS

#endif
#endif
