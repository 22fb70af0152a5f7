// lmpshim forwarding header: stands in for LAMMPS src/my_page.h (see lmpshim.h)
#include "lmpshim.h"
