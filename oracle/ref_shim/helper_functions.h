/* TEST INFRASTRUCTURE: stands in for the CUDA-samples header of the same name (see helper_cuda.h here). */
#pragma once
#include "helper_cuda.h"
