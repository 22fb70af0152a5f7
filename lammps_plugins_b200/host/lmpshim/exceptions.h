// lmpshim forwarding header: stands in for LAMMPS src/exceptions.h (see lmpshim.h)
#include "lmpshim.h"
