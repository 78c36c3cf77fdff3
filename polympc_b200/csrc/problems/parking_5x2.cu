// problems/parking_5x2.cu — kernels of Ocp<Parking, 5, 2> (NP = 1: free final time; see pmb_registry.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../pmb_registry.hpp"
PMB_DEFINE_PROBLEM(parking_5x2, Parking, 5, 2)
