// error.hpp -- status-carrying exception + per-thread last-error text behind pecs_last_error().
#pragma once
#include <stdexcept>
#include <string>

#include "../../include/pecs_b200.h"

namespace pecs {

struct StatusError : std::runtime_error {
  pecs_status status;
  StatusError(pecs_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};

void set_last_error(const std::string& text);

} // namespace pecs
