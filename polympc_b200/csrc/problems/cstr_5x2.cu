// problems/cstr_5x2.cu — kernels of Ocp<Cstr, 5, 2> (see pmb_registry.hpp)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../pmb_registry.hpp"
PMB_DEFINE_PROBLEM(cstr_5x2, Cstr, 5, 2)
