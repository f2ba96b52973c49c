// Test infrastructure only (see oracle/README.md).
// Minimal stand-in for <glog/logging.h> so that the reference's own sources
// (compiled where they lie under /root/reference by oracle/Makefile.ref) build
// without running the reference's cmake / glog configure step.  The reference
// includes glog from src/core/pbrt.h:61; only the macros below are used on the
// path-tracing hot path.  CHECKs abort like glog's; LOG/VLOG are swallowed.
#ifndef PB2_ORACLE_GLOG_SHIM_H
#define PB2_ORACLE_GLOG_SHIM_H
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace pb2_glog_shim {
struct NullStream {
    template <typename T> NullStream &operator<<(const T &) { return *this; }
    NullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
struct FatalStream {
    std::ostringstream ss;
    const char *file; int line;
    FatalStream(const char *f, int l, const char *what) : file(f), line(l) { ss << "Check failed: " << what << " "; }
    template <typename T> FatalStream &operator<<(const T &v) { ss << v; return *this; }
    FatalStream &operator<<(std::ostream &(*f)(std::ostream &)) { ss << f; return *this; }
    [[noreturn]] ~FatalStream() {
        std::fprintf(stderr, "[oracle/_ref] %s:%d %s\n", file, line, ss.str().c_str());
        std::abort();
    }
};
struct Voidify { void operator&(const NullStream &) {} void operator&(const FatalStream &) {} };
}  // namespace pb2_glog_shim

#define PB2_NULL_STREAM() true ? (void)0 : ::pb2_glog_shim::Voidify() & ::pb2_glog_shim::NullStream()
#define PB2_CHECK_IMPL(cond, text) \
    (cond) ? (void)0 : ::pb2_glog_shim::Voidify() & ::pb2_glog_shim::FatalStream(__FILE__, __LINE__, text)

#define LOG_INFO_STREAM PB2_NULL_STREAM()
#define LOG(severity) PB2_LOG_##severity
#define PB2_LOG_INFO PB2_NULL_STREAM()
#define PB2_LOG_WARNING PB2_NULL_STREAM()
#define PB2_LOG_ERROR PB2_NULL_STREAM()
#define PB2_LOG_FATAL PB2_CHECK_IMPL(false, "LOG(FATAL)")
#define VLOG(n) PB2_NULL_STREAM()
#define VLOG_IS_ON(n) false

#define CHECK(c) PB2_CHECK_IMPL((c), #c)
#define CHECK_EQ(a, b) PB2_CHECK_IMPL((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) PB2_CHECK_IMPL((a) != (b), #a " != " #b)
#define CHECK_LT(a, b) PB2_CHECK_IMPL((a) < (b), #a " < " #b)
#define CHECK_LE(a, b) PB2_CHECK_IMPL((a) <= (b), #a " <= " #b)
#define CHECK_GT(a, b) PB2_CHECK_IMPL((a) > (b), #a " > " #b)
#define CHECK_GE(a, b) PB2_CHECK_IMPL((a) >= (b), #a " >= " #b)
#define CHECK_NOTNULL(p) (p)
#define CHECK_NEAR(a, b, eps) PB2_CHECK_IMPL(std::abs((a) - (b)) <= (eps), #a " near " #b)
// NDEBUG build of the reference (CMakeLists.txt:64): DCHECKs compile away.
#define DCHECK(c) PB2_NULL_STREAM()
#define DCHECK_EQ(a, b) PB2_NULL_STREAM()
#define DCHECK_NE(a, b) PB2_NULL_STREAM()
#define DCHECK_LT(a, b) PB2_NULL_STREAM()
#define DCHECK_LE(a, b) PB2_NULL_STREAM()
#define DCHECK_GT(a, b) PB2_NULL_STREAM()
#define DCHECK_GE(a, b) PB2_NULL_STREAM()

namespace google {
inline void InitGoogleLogging(const char *) {}
inline void FlushLogFiles(int) {}
}
#endif
