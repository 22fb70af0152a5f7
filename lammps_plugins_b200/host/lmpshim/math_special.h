// lmpshim forwarding header: stands in for LAMMPS src/math_special.h (see lmpshim.h)
#include "lmpshim.h"
