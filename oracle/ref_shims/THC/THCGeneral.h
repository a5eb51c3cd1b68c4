/* oracle/ref_shims/THC/THCGeneral.h -- TEST INFRASTRUCTURE ONLY: see THC.h in this directory. */
#include "THC.h"
