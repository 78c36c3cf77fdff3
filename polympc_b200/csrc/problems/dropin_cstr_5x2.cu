// problems/dropin_cstr_5x2.cu — kernels of a reference-style problem class (Eigen functors, ContinuousOCP<> CRTP) compiled
// through the source-compatibility layer include/polympc_compat/ (see pmb_registry.hpp and polympc_compat.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../../../examples/dropin/cstr_ocp.hpp"
PMB_DEFINE_COMPAT_PROBLEM(dropin_cstr_5x2, dropin::CstrOCP)
