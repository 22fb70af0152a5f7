// lmpshim forwarding header: stands in for LAMMPS src/lammpsplugin.h (see lmpshim.h)
#include "lmpshim.h"
