// lmpshim forwarding header: stands in for LAMMPS src/utils.h (see lmpshim.h)
#include "lmpshim.h"
