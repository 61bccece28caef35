// Drop-in stub for libelas/src/descriptor.cpp (listed in stereomapper.pro:26): intentionally empty,
// see descriptor.h.
#include "descriptor.h"
