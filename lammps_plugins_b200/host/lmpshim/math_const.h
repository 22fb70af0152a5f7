// lmpshim forwarding header: stands in for LAMMPS src/math_const.h (see lmpshim.h)
#include "lmpshim.h"
