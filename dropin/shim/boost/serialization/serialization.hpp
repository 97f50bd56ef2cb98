// Minimal stand-in for boost::serialization: the reference's headers only declare `friend class access` and member templates
// `serialize(Archive&, unsigned)` that are never instantiated on the ORB hot path.  See dropin/cvmin/cvmin.h for why stand-ins exist.
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
template <class T> T* make_array(T* p, unsigned long) { return p; }
}}
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#define BOOST_CLASS_EXPORT_KEY(x)
#define BOOST_CLASS_EXPORT_IMPLEMENT(x)
#define BOOST_SERIALIZATION_ASSUME_ABSTRACT(x)
