// problems/dropin_robot_5x3.cu — kernels of a reference-style problem class (Eigen functors, ContinuousOCP<> CRTP) compiled
// through the source-compatibility layer include/polympc_compat/ (see pmb_registry.hpp and polympc_compat.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../../../examples/dropin/robot_ocp.hpp"
PMB_DEFINE_COMPAT_PROBLEM(dropin_robot_5x3, dropin::RobotOCP)
