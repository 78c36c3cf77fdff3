// problems/robot_obstacle_5x2.cu — kernels of Ocp<MobileRobotObstacle, 5, 2> (see pmb_registry.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../pmb_registry.hpp"
PMB_DEFINE_PROBLEM(robot_obstacle_5x2, MobileRobotObstacle, 5, 2)
