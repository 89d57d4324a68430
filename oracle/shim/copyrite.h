/* Stand-in for the Gromacs 4.5 header of the same name; everything lives in gmxshim.h. */
#include "gmxshim.h"
