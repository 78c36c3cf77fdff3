// tests/dropin/dropin_plugin.cu — TEST INFRASTRUCTURE.  A *user translation unit*: two reference-style problem classes (Eigen
// functors over ContinuousOCP<>, robot_ocp.hpp / cstr_ocp.hpp in this directory) compiled through include/polympc_compat/
// and registered with the engine when this shared object is loaded (pmb_register_problem), the way any out-of-tree problem
// class reaches the engine.  Built twice by the Makefile next to it: nvcc + libpolympc_b200.so (GPU suite) and g++ -DPMB_EMU +
// libpmb_emu.so (CPU suite).  Nothing of this is part of the product library.
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "robot_ocp.hpp"
#include "cstr_ocp.hpp"
#include <cstdio>
#include <cstdlib>

namespace {
struct Registrar {
    Registrar()
    {
        const int a = pmb_register_problem("dropin_robot_5x3", (void* (*)())(&pmb::compat::make_problem<dropin::RobotOCP>));
        const int b = pmb_register_problem("dropin_cstr_5x2", (void* (*)())(&pmb::compat::make_problem<dropin::CstrOCP>));
        if (a != PMB_OK || b != PMB_OK) { std::fprintf(stderr, "dropin_plugin: registration failed: %s\n", pmb_last_error()); std::abort(); }
    }
} g_registrar;
}
