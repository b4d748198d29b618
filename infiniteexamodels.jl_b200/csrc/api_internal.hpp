// api_internal.hpp — the opaque iexa_plan of include/iexa.h (shared with tests/hostcheck).
#pragma once
#include <memory>

#include "engine.hpp"
#include "plan.hpp"

struct iexa_plan {
  iexa::Plan plan;
  std::unique_ptr<iexa::Engine> engine;
  uint32_t flags = 0;
  int64_t bytes_cache[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
};
