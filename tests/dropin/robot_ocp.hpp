// examples/dropin/robot_ocp.hpp — a mobile-robot OCP written the way a PolyMPC user writes it: against the header-only
// ContinuousOCP<> / Chebyshev<> / Spline<> API with Eigen-style functors (the functor concept of reference
// src/control/continuous_ocp.hpp:191-288; same model and cost as the reference's tests/control/mpc_wrapper_test.cpp:38-80:
// unicycle with wheel base d(0), quadratic stage cost and terminal cost).  Nothing in this file knows about CUDA: it
// compiles against include/polympc_compat/ with nvcc (kernels) and g++ (warp emulator) alike, and
// polympc_b200/csrc/problems/dropin_robot_5x3.cu turns it into the engine problem "dropin_robot_5x3".
#pragma once
#include "polynomials/ebyshev.hpp"
#include "polynomials/splines.hpp"
#include "control/continuous_ocp.hpp"

#ifndef DROPIN_ROBOT_SEGMENTS
#define DROPIN_ROBOT_SEGMENTS 3     // mpc_wrapper_test.cpp uses 3 segments of order 5; the CasADi fixture grid is 5 x 2
#endif

namespace dropin {
using RobotPolynomial = polympc::Chebyshev<5, polympc::GAUSS_LOBATTO, double>;
using RobotApproximation = polympc::Spline<RobotPolynomial, DROPIN_ROBOT_SEGMENTS>;
class RobotOCP;
}
template <> struct polympc_traits<dropin::RobotOCP> { using Scalar = double; enum { NX = 3, NU = 2, NP = 0, ND = 1, NG = 0 }; };

namespace dropin {

class RobotOCP : public ContinuousOCP<RobotOCP, RobotApproximation, DENSE>
{
public:
    Eigen::DiagonalMatrix<scalar_t, 3> Q{1, 1, 1};
    Eigen::DiagonalMatrix<scalar_t, 2> R{1, 1};
    Eigen::DiagonalMatrix<scalar_t, 3> QN{1, 1, 1};

    /** unicycle: position (x0, x1), heading x2; speed u0, steering angle u1; wheel base d0 */
    template <typename T>
    inline void dynamics_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                              const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t>& d,
                              const T& t, Eigen::Ref<state_t<T>> xdot) const noexcept
    {
        xdot(0) = u(0) * cos(x(2)) * cos(u(1));
        xdot(1) = u(0) * sin(x(2)) * cos(u(1));
        xdot(2) = u(0) * sin(u(1)) / d(0);
    }

    template <typename T>
    inline void lagrange_term_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                                   const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t> d,
                                   const scalar_t& t, T& lagrange) noexcept
    {
        const Eigen::Matrix<T, 3, 3> Qm = Q.toDenseMatrix().template cast<T>();
        const Eigen::Matrix<T, 2, 2> Rm = R.toDenseMatrix().template cast<T>();
        lagrange = x.dot(Qm * x) + u.dot(Rm * u);
    }

    template <typename T>
    inline void mayer_term_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                                const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t> d,
                                const scalar_t& t, T& mayer) noexcept
    {
        const Eigen::Matrix<T, 3, 3> Qm = Q.toDenseMatrix().template cast<T>();
        mayer = x.dot(Qm * x);
    }

    void set_Q_coeff(const scalar_t& coeff) { Q.diagonal() << coeff, coeff, coeff; }
};

} // namespace dropin
