// lmpshim forwarding header: stands in for LAMMPS src/atom.h (see lmpshim.h)
#include "lmpshim.h"
