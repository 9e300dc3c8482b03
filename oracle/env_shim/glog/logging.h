// oracle/env_shim/glog — stand-in for glog: LOG(severity) swallows everything streamed into it.  TEST INFRASTRUCTURE ONLY.
#pragma once
namespace rvs_glog_shim {
struct NullStream {
    template <class T> NullStream &operator<<(const T &) { return *this; }
    NullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
}  // namespace rvs_glog_shim
#define LOG(severity) ::rvs_glog_shim::NullStream()
