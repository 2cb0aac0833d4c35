#include "irtk_stub.h"
