// lmpshim forwarding header: stands in for LAMMPS src/version.h (see lmpshim.h)
#include "lmpshim.h"
