// ORACLE BUILD SHIM (test infrastructure): the reference's DBoW2 headers only need these
// names to exist; serialization is never exercised by the oracle.
#ifndef ORB_ORACLE_SHIM_BOOST_SER_HPP
#define ORB_ORACLE_SHIM_BOOST_SER_HPP
namespace boost
{
    namespace serialization
    {
        class access
        {
        };
        template <class Base, class Derived>
        Base &base_object(Derived &d) { return static_cast<Base &>(d); }
    }
}
#endif
