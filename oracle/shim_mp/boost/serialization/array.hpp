// ORACLE BUILD SHIM (test infrastructure)
#include <boost/serialization/serialization.hpp>
namespace boost { namespace serialization { template <class T> int make_array(T *, unsigned long) { return 0; } } }
