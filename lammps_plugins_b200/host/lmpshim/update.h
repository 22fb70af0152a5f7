// lmpshim forwarding header: stands in for LAMMPS src/update.h (see lmpshim.h)
#include "lmpshim.h"
