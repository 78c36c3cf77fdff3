// examples/dropin/cstr_ocp.hpp — continuous stirred-tank reactor stabilisation, written against the header-only PolyMPC API
// with Eigen-style functors (functor concept: reference src/control/continuous_ocp.hpp:191-288; model, weights and steady
// state as in the reference's tests/control/cstr_control_test.cpp:34-113).  Compiles unchanged against
// include/polympc_compat/; polympc_b200/csrc/problems/dropin_cstr_5x2.cu registers it as "dropin_cstr_5x2".
#pragma once
#include "polynomials/ebyshev.hpp"
#include "polynomials/splines.hpp"
#include "control/continuous_ocp.hpp"

namespace dropin {
using CstrPolynomial = polympc::Chebyshev<5, polympc::GAUSS_LOBATTO, double>;
using CstrApproximation = polympc::Spline<CstrPolynomial, 2>;
class CstrOCP;
}
template <> struct polympc_traits<dropin::CstrOCP> { using Scalar = double; enum { NX = 4, NU = 2, NP = 0, ND = 0, NG = 0 }; };

namespace dropin {

class CstrOCP : public ContinuousOCP<CstrOCP, CstrApproximation, DENSE>
{
public:
    Eigen::Matrix<scalar_t, 4, 4> Q, P;
    Eigen::Matrix<scalar_t, 2, 2> R;
    Eigen::Matrix<scalar_t, 4, 1> xs;
    Eigen::Matrix<scalar_t, 2, 1> us;

    CstrOCP()
    {
        Q.setZero();
        R.setZero();
        Q.diagonal() << 0.2, 1.0, 0.5, 0.2;
        R.diagonal() << 0.5, 5.0 * 1.0e-7;
        P << 1.4646778374584373, 0.6676889516721198, 0.35446715117028615, 0.10324422005086348,
             0.6676889516721198, 1.407812935783267, 0.17788030743777067, 0.050059833257226405,
             0.3544671511702861, 0.1778803074377706, 0.6336052592712396, 0.01110329497282364,
             0.1032442200508634, 0.05005983325722643, 0.011103294972823655, 0.229412393739723;
        xs << 2.1402105301746182e00, 1.0903043613077321e00, 1.1419108442079495e02, 1.1290659291045561e02;
        us << 14.19, -1113.50;
        set_time_limits(0, 100);   // slow process: 100 s horizon
    }

    /** x = [c_A, c_B, reactor temperature, jacket temperature], u = [feed rate, cooling power] */
    template <typename T>
    inline void dynamics_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                              const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t>& d,
                              const T& t, Eigen::Ref<state_t<T>> xdot) const noexcept
    {
        // plant data (Klatt-Engell reactor, the benchmark the reference test uses); every constant enters as a T so that the
        // operation tree — and with it every rounding — is the one of the engine's built-in "cstr_5x2" problem
        const T feed_conc = (T)5.1, feed_temp = (T)104.9;                       // c_A0 [mol/l], inflow temperature [C]
        const T kw = (T)4032.0, area = (T)0.215, volume = (T)10.0;              // jacket heat transfer, surface, reactor volume
        const T density = (T)0.9342, cp = (T)3.01;                              // of the mixture
        const T dh_ab = (T)4.2, dh_bc = (T)-11.0, dh_ad = (T)-41.85;            // reaction enthalpies
        const T coolant_mass = (T)5.0, cp_coolant = (T)2.0;
        const T k0_ab = (T)1.287e12, k0_bc = (T)1.287e12, k0_ad = (T)9.043e09;  // Arrhenius: k = k0 exp(E / (273.15 + temp))
        const T e_ab = (T)-9758.3, e_bc = (T)-9758.3, e_ad = (T)-8560.0;
        const T r_ab = k0_ab * exp(e_ab / (273.15 + x(2)));
        const T r_bc = k0_bc * exp(e_bc / (273.15 + x(2)));
        const T r_ad = k0_ad * exp(e_ad / (273.15 + x(2)));
        const T hour = (T)3600.0;                                               // the model is written per hour, the OCP per second
        const T per_second = 1 / hour;

        xdot(0) = per_second * (u(0) * (feed_conc - x(0)) - r_ab * x(0) - r_ad * x(0) * x(0));
        xdot(1) = per_second * (-u(0) * x(1) + r_ab * x(0) - r_bc * x(1));
        xdot(2) = per_second * (u(0) * (feed_temp - x(2)) + (kw * area / (density * cp * volume)) * (x(3) - x(2))
                                - (1 / (density * cp)) * (r_ab * x(0) * dh_ab + r_bc * x(1) * dh_bc + r_ad * x(0) * x(1) * dh_ad));
        xdot(3) = per_second * ((1 / (coolant_mass * cp_coolant)) * (u(1) + kw * area * (x(2) - x(3))));
    }

    template <typename T>
    inline void lagrange_term_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                                   const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t> d,
                                   const scalar_t& t, T& lagrange) noexcept
    {
        lagrange = (x - xs).dot(Q * (x - xs)) + (u - us).dot(R * (u - us));
    }

    template <typename T>
    inline void mayer_term_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                                const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t> d,
                                const scalar_t& t, T& mayer) noexcept
    {
        mayer = (x - xs).dot(P * (x - xs));
    }
};

} // namespace dropin
