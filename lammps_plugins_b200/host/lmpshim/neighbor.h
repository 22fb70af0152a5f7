// lmpshim forwarding header: stands in for LAMMPS src/neighbor.h (see lmpshim.h)
#include "lmpshim.h"
