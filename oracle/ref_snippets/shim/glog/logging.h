// Stand-in for <glog/logging.h> (glog 0.5.0 is fetched by nvblox's CMake and is absent offline).
// TEST INFRASTRUCTURE ONLY: lets the reference's headers compile; every log / check is a no-op that
// still evaluates (and type-checks) nothing at run time.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
namespace shim_glog {
struct NullStream {
  template <typename T>
  __host__ __device__ NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct Voidify {
  __host__ __device__ void operator&(NullStream&) {}
};
}  // namespace shim_glog
#define SHIM_GLOG_STREAM() ::shim_glog::NullStream()
#define LOG(sev) SHIM_GLOG_STREAM()
#define VLOG(n) SHIM_GLOG_STREAM()
#define DLOG(sev) SHIM_GLOG_STREAM()
#define LOG_IF(sev, c) SHIM_GLOG_STREAM()
#define LOG_EVERY_N(sev, n) SHIM_GLOG_STREAM()
#define LOG_FIRST_N(sev, n) SHIM_GLOG_STREAM()
#define CHECK(c) (void)(c), SHIM_GLOG_STREAM()
#define DCHECK(c) SHIM_GLOG_STREAM()
#define CHECK_NOTNULL(p) (p)
#define SHIM_CHECK_OP(a, b) (void)(a), (void)(b), SHIM_GLOG_STREAM()
#define CHECK_EQ(a, b) SHIM_CHECK_OP(a, b)
#define CHECK_NE(a, b) SHIM_CHECK_OP(a, b)
#define CHECK_LT(a, b) SHIM_CHECK_OP(a, b)
#define CHECK_LE(a, b) SHIM_CHECK_OP(a, b)
#define CHECK_GT(a, b) SHIM_CHECK_OP(a, b)
#define CHECK_GE(a, b) SHIM_CHECK_OP(a, b)
#define CHECK_NEAR(a, b, c) SHIM_CHECK_OP(a, b)
#define DCHECK_EQ(a, b) SHIM_GLOG_STREAM()
#define DCHECK_NE(a, b) SHIM_GLOG_STREAM()
#define DCHECK_LT(a, b) SHIM_GLOG_STREAM()
#define DCHECK_LE(a, b) SHIM_GLOG_STREAM()
#define DCHECK_GT(a, b) SHIM_GLOG_STREAM()
#define DCHECK_GE(a, b) SHIM_GLOG_STREAM()
