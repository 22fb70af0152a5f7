// lmpshim forwarding header: stands in for LAMMPS src/lammps.h (see lmpshim.h)
#include "lmpshim.h"
