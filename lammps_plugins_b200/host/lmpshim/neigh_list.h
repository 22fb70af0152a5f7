// lmpshim forwarding header: stands in for LAMMPS src/neigh_list.h (see lmpshim.h)
#include "lmpshim.h"
