// lmpshim forwarding header: stands in for LAMMPS src/lmptype.h (see lmpshim.h)
#include "lmpshim.h"
