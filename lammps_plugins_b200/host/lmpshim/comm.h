// lmpshim forwarding header: stands in for LAMMPS src/comm.h (see lmpshim.h)
#include "lmpshim.h"
