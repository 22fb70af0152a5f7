// lmpshim forwarding header: stands in for LAMMPS src/text_file_reader.h (see lmpshim.h)
#include "lmpshim.h"
