#include "gsl_shim.h"
