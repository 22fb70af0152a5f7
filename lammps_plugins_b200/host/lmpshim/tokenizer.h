// lmpshim forwarding header: stands in for LAMMPS src/tokenizer.h (see lmpshim.h)
#include "lmpshim.h"
