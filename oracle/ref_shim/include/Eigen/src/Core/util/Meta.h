// stand-in: the matrix types come from oracle/ref_shim/ref_eigen.hpp (force-included)
#include "ref_eigen.hpp"
