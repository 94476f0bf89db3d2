// ORACLE BUILD SHIM (test infrastructure): see orbslam_stubs.h
#include "orbslam_stubs.h"
