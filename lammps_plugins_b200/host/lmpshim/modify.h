// API-surface shim: see lmpshim.h.  In LAMMPS modify.h declares class Modify (the pair classes read n_pre_force);
// here the type stays opaque and the host classes do not look inside it (b200md_host.h, forces_zero_on_entry).
#include "lmpshim.h"
