// TEST INFRASTRUCTURE (oracle/_ref build only). No-op stand-in for glog, which is absent in this image.
// The reference's include/logger.h:3 pulls <glog/logging.h>; the hot path only uses LOG(x) << ... for diagnostics.
#pragma once
#include <ostream>
struct ts_null_log_stream {
    template <class T> ts_null_log_stream& operator<<(const T&) { return *this; }
    ts_null_log_stream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
#define LOG(severity) ts_null_log_stream()
#define LOG_IF(severity, cond) ts_null_log_stream()
#define VLOG(level) ts_null_log_stream()
