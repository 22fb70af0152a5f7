// lmpshim forwarding header: stands in for LAMMPS src/force.h (see lmpshim.h)
#include "lmpshim.h"
