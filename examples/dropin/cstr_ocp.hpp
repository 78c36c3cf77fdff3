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
        T c_AO = (T)5.1;
        T v_0 = (T)104.9;
        T k_w = (T)4032.0;
        T A_R = (T)0.215;
        T rho = (T)0.9342;
        T C_P = (T)3.01;
        T V_R = (T)10.0;
        T H_1 = (T)4.2;
        T H_2 = (T)-11.0;
        T H_3 = (T)-41.85;
        T m_K = (T)5.0;
        T C_PK = (T)2.0;
        T k10 = (T)1.287e12;
        T k20 = (T)1.287e12;
        T k30 = (T)9.043e09;
        T E1 = (T)-9758.3;
        T E2 = (T)-9758.3;
        T E3 = (T)-8560.0;
        T k_1 = k10 * exp(E1 / (273.15 + x(2)));
        T k_2 = k20 * exp(E2 / (273.15 + x(2)));
        T k_3 = k30 * exp(E3 / (273.15 + x(2)));
        T TIMEUNITS_PER_HOUR = (T)3600.0;

        xdot(0) = (1 / TIMEUNITS_PER_HOUR) * (u(0) * (c_AO - x(0)) - k_1 * x(0) - k_3 * x(0) * x(0));
        xdot(1) = (1 / TIMEUNITS_PER_HOUR) * (-u(0) * x(1) + k_1 * x(0) - k_2 * x(1));
        xdot(2) = (1 / TIMEUNITS_PER_HOUR) * (u(0) * (v_0 - x(2)) + (k_w * A_R / (rho * C_P * V_R)) *
                                              (x(3) - x(2)) - (1 / (rho * C_P)) * (k_1 * x(0) * H_1 + k_2 * x(1) * H_2 + k_3 * x(0) * x(1) * H_3));
        xdot(3) = (1 / TIMEUNITS_PER_HOUR) * ((1 / (m_K * C_PK)) * (u(1) + k_w * A_R * (x(2) - x(3))));
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
