// lmpshim forwarding header: stands in for LAMMPS src/pair.h (see lmpshim.h)
#include "lmpshim.h"
