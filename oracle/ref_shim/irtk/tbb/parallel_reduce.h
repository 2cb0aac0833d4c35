#include "tbb_shim.h"
