// lmpshim forwarding header: stands in for LAMMPS src/pointers.h (see lmpshim.h)
#include "lmpshim.h"
