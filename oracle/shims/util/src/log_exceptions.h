// oracle/shims: graph.cc uses none of the throw-macros; empty on purpose.
#pragma once
#include <stdexcept>
