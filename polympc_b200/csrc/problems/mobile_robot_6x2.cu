// problems/mobile_robot_6x2.cu — kernels of Ocp<MobileRobot, 6, 2> (see pmb_registry.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../pmb_registry.hpp"
PMB_DEFINE_PROBLEM(mobile_robot_6x2, MobileRobot, 6, 2)
