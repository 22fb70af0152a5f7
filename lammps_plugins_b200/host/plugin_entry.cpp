/* ----------------------------------------------------------------------
   LAMMPS plugin entry point of the B200 pair styles.  Built twice:
     -DB200MD_PLUGIN_REBOMOS -> rebomosplugin.so registers pair style "rebomos"
     -DB200MD_PLUGIN_AEAM    -> aeamplugin.so    registers pair style "aeam"
   Same exported symbol, style names and registration protocol as the reference
   loaders (USER-REBOMOS/rebomosplugin.cpp:14-28, USER-AEAM/aeamplugin.cpp:14-28):
   LAMMPS dlopen()s the file, calls lammpsplugin_init(lmp, handle, regfunc), and the
   registered factory overrides any built-in style of the same name.
------------------------------------------------------------------------- */

#include "lammpsplugin.h"
#include "version.h"

#if defined(B200MD_PLUGIN_REBOMOS)
#include "pair_rebomos.h"
#define STYLE_NAME "rebomos"
#define STYLE_CLASS PairREBOMoS
#define STYLE_INFO "REBOMoS pair style, B200 (sm_100a) implementation v1.0"
#elif defined(B200MD_PLUGIN_AEAM)
#include "pair_aeam.h"
#define STYLE_NAME "aeam"
#define STYLE_CLASS PairAEAM
#define STYLE_INFO "AEAM pair style, B200 (sm_100a) implementation v1.0"
#else
#error "define B200MD_PLUGIN_REBOMOS or B200MD_PLUGIN_AEAM"
#endif

using namespace LAMMPS_NS;

static Pair *b200_pair_factory(LAMMPS *lmp)
{
  return new STYLE_CLASS(lmp);
}

extern "C" __attribute__((visibility("default"))) void lammpsplugin_init(void *lmp, void *handle, void *regfunc)
{
  lammpsplugin_t plugin;
  plugin.version = LAMMPS_VERSION;
  plugin.style = "pair";
  plugin.name = STYLE_NAME;
  plugin.info = STYLE_INFO;
  plugin.author = "b200md";
  plugin.creator.v1 = (lammpsplugin_factory1 *) &b200_pair_factory;
  plugin.handle = handle;
  ((lammpsplugin_regfunc) regfunc)(&plugin, lmp);
}
