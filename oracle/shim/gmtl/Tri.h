/* forwarding header: the whole GMTL stand-in lives in gmtl.h */
#include "gmtl.h"
