// lmpshim forwarding header: stands in for LAMMPS src/mpi.h (see lmpshim.h)
#include "lmpshim.h"
