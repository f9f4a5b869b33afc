// stand-in: cv::Mat_ and the few calls the CAPE sources make come from oracle/ref_shim/ref_cv.hpp (force-included)
#include "ref_cv.hpp"
