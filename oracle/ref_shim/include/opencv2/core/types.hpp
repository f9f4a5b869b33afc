// stand-in: see oracle/ref_shim/ref_cv.hpp (force-included)
#include "ref_cv.hpp"
