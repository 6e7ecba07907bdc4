// error.cpp -- see error.hpp.
#include "error.hpp"

namespace {
thread_local std::string g_last_error;
}

namespace pecs {
void set_last_error(const std::string& text) { g_last_error = text; }
} // namespace pecs

extern "C" const char* pecs_last_error(void) { return g_last_error.c_str(); }
