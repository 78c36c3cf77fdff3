// problems/kite_12x1.cu — kernels of Ocp<Kite, 12, 1> (see pmb_registry.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../pmb_registry.hpp"
PMB_DEFINE_PROBLEM(kite_12x1, Kite, 12, 1)
