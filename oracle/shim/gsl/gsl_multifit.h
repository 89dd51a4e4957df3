/* oracle shim (test infrastructure): forwards to gsl_shim_all.h */
#include "gsl_shim_all.h"
