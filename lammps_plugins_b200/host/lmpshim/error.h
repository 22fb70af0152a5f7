// lmpshim forwarding header: stands in for LAMMPS src/error.h (see lmpshim.h)
#include "lmpshim.h"
