// ORACLE BUILD SHIM (test infrastructure)
#include <boost/serialization/serialization.hpp>
