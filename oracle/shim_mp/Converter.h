// ORACLE BUILD SHIM (test infrastructure): see mp_stubs.h
#include "mp_stubs.h"
