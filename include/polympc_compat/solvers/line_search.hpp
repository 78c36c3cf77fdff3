// Path-compatible forwarding header: with -I include/polympc_compat, `#include "solvers/line_search.hpp"` (reference src/solvers/line_search.hpp)
// resolves to the B200 source-compatibility layer.  See ../polympc_compat.hpp for what is provided.
#pragma once
#include "../polympc_compat.hpp"
