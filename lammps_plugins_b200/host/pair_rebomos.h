/* -*- c++ -*- ----------------------------------------------------------
   pair_style rebomos -- B200-native implementation.

   Same style name, pair_coeff grammar and potential-file format as the
   reference USER-REBOMOS package (lammps/lammps-plugins
   USER-REBOMOS/pair_rebomos.h:30-60, pair_rebomos.cpp:144-274, 857-1107);
   the force computation itself runs on the GPU through the b200md C ABI
   (include/b200md.h).  This class owns nothing numerical beyond the parsed
   parameter block.
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS
// clang-format off
PairStyle(rebomos,PairREBOMoS);
// clang-format on
#else

#ifndef LMP_PAIR_REBOMOS_B200_H
#define LMP_PAIR_REBOMOS_B200_H

#include "b200md.h"
#include "pair.h"
#include "b200md_host.h"

namespace LAMMPS_NS {

class PairREBOMoS : public Pair {
 public:
  PairREBOMoS(class LAMMPS *);
  ~PairREBOMoS() override;
  void compute(int, int) override;
  void settings(int, char **) override;
  void coeff(int, char **) override;
  void init_style() override;
  double init_one(int, int) override;
  double memory_usage() override;

 protected:
  b200md_rebomos_params params;    // as read from the file (+ mixed LJ terms)
  b200md_ctx *ctx;                 // device context, created in init_style()
  double cut3rebo;
  bigint last_list_step;           // timestep stamp of the list now on the device
  int uploaded_nlocal, uploaded_nghost;
  int f_overwrite = -1;                   // what the library was last told (b200md_set_option "f_overwrite")
  B200MDHost::PinnedAtomArrays pinned;    // atom->x / atom->f page-locked for DMA beside the kernels

  void read_file(char *);
  void allocate();
};
}    // namespace LAMMPS_NS

#endif
#endif
