// problems/kite_4x2.cu — kernels of Ocp<Kite, 4, 2> (see pmb_registry.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../pmb_registry.hpp"
PMB_DEFINE_PROBLEM(kite_4x2, Kite, 4, 2)
