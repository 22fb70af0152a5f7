// lmpshim forwarding header: stands in for LAMMPS src/domain.h (see lmpshim.h)
#include "lmpshim.h"
