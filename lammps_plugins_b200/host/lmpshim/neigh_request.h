// lmpshim forwarding header: stands in for LAMMPS src/neigh_request.h (see lmpshim.h)
#include "lmpshim.h"
