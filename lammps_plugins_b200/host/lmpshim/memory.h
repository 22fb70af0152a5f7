// lmpshim forwarding header: stands in for LAMMPS src/memory.h (see lmpshim.h)
#include "lmpshim.h"
