// problems/mobile_robot_5x3.cu — kernels of Ocp<MobileRobot, 5, 3> (see pmb_registry.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../pmb_registry.hpp"
PMB_DEFINE_PROBLEM(mobile_robot_5x3, MobileRobot, 5, 3)
