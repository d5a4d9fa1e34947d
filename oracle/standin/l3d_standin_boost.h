// oracle/standin/l3d_standin_boost.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Stand-ins for the Boost names the reference's Line3D++ headers mention (mutexes around its OpenMP loops,
// the serialization hooks of its result types, filesystem calls of the segment cache), so that the sources
// compile unmodified into oracle/_ref.  The cache and the archives are unused.
#pragma once
#include <fstream>
#include <mutex>
#include <string>
#include <sys/stat.h>

namespace boost {
// a real lock: the _omp build (timing only) runs the reference's OpenMP loops, which rely on these mutexes
class mutex {
  public:
    void lock() { m_.lock(); }
    void unlock() { m_.unlock(); }
  private:
    std::mutex m_;
};
namespace filesystem {
class path {
  public:
    path() {}
    path(const std::string& s) : s_(s) {}
    path(const char* s) : s_(s) {}
    const std::string& string() const { return s_; }
  private:
    std::string s_;
};
inline bool create_directory(const path&) { return true; }   // the reference only uses it for its segment cache
inline bool exists(const path& p)
{
    struct stat st;
    return ::stat(p.string().c_str(), &st) == 0;
}
}  // namespace filesystem
namespace serialization {
class access {};
template <class T>
struct nvp_t {
    T& v;
};
template <class T>
inline nvp_t<T> make_nvp(const char*, T& v) { return nvp_t<T>{v}; }
template <class T>
inline nvp_t<const T> make_nvp(const char*, const T& v) { return nvp_t<const T>{v}; }
template <class T>
inline int make_array(T*, size_t) { return 0; }
}  // namespace serialization
namespace archive {
// archives that swallow everything: serialization is off in the reference configuration (loadAndStore = false)
class binary_oarchive {
  public:
    explicit binary_oarchive(std::ostream&) {}
    template <class T>
    binary_oarchive& operator&(const T&) { return *this; }
};
class binary_iarchive {
  public:
    explicit binary_iarchive(std::istream&) {}
    template <class T>
    binary_iarchive& operator&(const T&) { return *this; }
};
}  // namespace archive
}  // namespace boost
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#define BOOST_CLASS_VERSION(T, N)
