// pmb_problems.hpp — problem classes (dynamics / Lagrange / Mayer / inequality functors) instantiated by the GPU engine.
//
// Each class is the B200-side twin of a reference problem class; functors are templates over the scalar type T
// (double, Dual<double,n>, Dual<Dual<double,n>,n>) exactly like the reference's dynamics_impl / lagrange_term_impl /
// mayer_term_impl / inequality_constraints_impl (reference src/control/continuous_ocp.hpp:191-288), but take plain
// pointers instead of Eigen::Ref so they can be inlined into kernels.  Data members (Q, R, xs ...) are plain doubles and
// the object is trivially copyable: it is passed to kernels by value.
//   MobileRobot : reference tests/control/mpc_wrapper_test.cpp:38-80     (NX=3, NU=2, ND=1)
//   Cstr        : reference tests/control/cstr_control_test.cpp:34-113   (NX=4, NU=2)
//   Kite        : NOT in the reference tree — our 13-state rigid-body kite, BASELINE config 4.
// Small fixed-size products (x.dot(Q*x)) use Eigen's unrolled halving order, spelled out term by term.
#pragma once
#include "pmb_dual.hpp"

namespace pmb {

struct MobileRobot {
    static constexpr int NX = 3, NU = 2, NP = 0, ND = 1, NG = 0, NPARAM = 8;
    double Q[3], R[2], QN[3];
    PMB_HD void defaults() { Q[0] = Q[1] = Q[2] = 1; R[0] = R[1] = 1; QN[0] = QN[1] = QN[2] = 1; }
    PMB_HD void set_params(const double* v) { Q[0] = v[0]; Q[1] = v[1]; Q[2] = v[2]; R[0] = v[3]; R[1] = v[4]; QN[0] = v[5]; QN[1] = v[6]; QN[2] = v[7]; }
    PMB_HD void get_params(double* v) const { v[0] = Q[0]; v[1] = Q[1]; v[2] = Q[2]; v[3] = R[0]; v[4] = R[1]; v[5] = QN[0]; v[6] = QN[1]; v[7] = QN[2]; }

    /** mpc_wrapper_test.cpp:46-55 */
    template <class T>
    PMB_HD void dynamics(const T* x, const T* u, const T*, const double* d, const T&, T* xdot) const
    {
        xdot[0] = u[0] * cos(x[2]) * cos(u[1]);
        xdot[1] = u[0] * sin(x[2]) * cos(u[1]);
        xdot[2] = u[0] * sin(u[1]) / d[0];
    }
    /** y = diag(q).toDenseMatrix().cast<T>() * x for a 3-vector: row i = (m_i0*x0) + ((m_i1*x1) + (m_i2*x2)) */
    template <class T>
    PMB_HD static void diag3_times(const double* q, const T* x, T* y)
    {
        y[0] = (T(q[0]) * x[0]) + ((T(0.0) * x[1]) + (T(0.0) * x[2]));
        y[1] = (T(0.0) * x[0]) + ((T(q[1]) * x[1]) + (T(0.0) * x[2]));
        y[2] = (T(0.0) * x[0]) + ((T(0.0) * x[1]) + (T(q[2]) * x[2]));
    }
    /** mpc_wrapper_test.cpp:57-66:  x.dot(Qm * x) + u.dot(Rm * u) */
    template <class T>
    PMB_HD void lagrange(const T* x, const T* u, const T*, const double*, double, T& L) const
    {
        T Qx[3], Ru[2];
        diag3_times(Q, x, Qx);
        Ru[0] = (T(R[0]) * u[0]) + (T(0.0) * u[1]);
        Ru[1] = (T(0.0) * u[0]) + (T(R[1]) * u[1]);
        L = ((x[0] * Qx[0]) + ((x[1] * Qx[1]) + (x[2] * Qx[2]))) + ((u[0] * Ru[0]) + (u[1] * Ru[1]));
    }
    /** mpc_wrapper_test.cpp:68-74:  x.dot(Qm * x)  (the reference uses Q, not QN) */
    template <class T>
    PMB_HD void mayer(const T* x, const T*, const T*, const double*, double, T& M) const
    {
        T Qx[3];
        diag3_times(Q, x, Qx);
        M = (x[0] * Qx[0]) + ((x[1] * Qx[1]) + (x[2] * Qx[2]));
    }
    template <class T> PMB_HD void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

/** mobile robot with one generic inequality constraint per node (NG = 1): squared distance to a disc-shaped obstacle.
 *  Exercises the NG > 0 paths (reference continuous_ocp.hpp:546-575, 769-782, 2150-2157; sqp_base.hpp:425-443, 457-465). */
struct MobileRobotObstacle : MobileRobot {
    static constexpr int NG = 1;
    template <class T>
    PMB_HD void ineq(const T* x, const T*, const T*, const double*, double, T* g) const
    {
        g[0] = (x[0] - 0.25) * (x[0] - 0.25) + (x[1] - 0.25) * (x[1] - 0.25);
    }
};

/** free-final-time valet parking: the reference's ParkingOCP (tests/control/minimal_time_test.cpp:32-63,
 *  tests/control/dense_sparse_compare.cpp:22-55): mobile-robot kinematics scaled by the optimised parameter p(0) = final time
 *  on the normalised horizon [0, 1]; cost = Mayer term p(0).  NP = 1 exercises the parameter columns / blocks
 *  (continuous_ocp.hpp:860-872, 1314-1366, 2161-2172). */
struct Parking {
    static constexpr int NX = 3, NU = 2, NP = 1, ND = 1, NG = 0, NPARAM = 1;
    double unused;
    PMB_HD void defaults() { unused = 0; }
    PMB_HD void set_params(const double* v) { unused = v[0]; }
    PMB_HD void get_params(double* v) const { v[0] = unused; }
    template <class T>
    PMB_HD void dynamics(const T* x, const T* u, const T* p, const double* d, const T&, T* xdot) const
    {
        xdot[0] = p[0] * u[0] * cos(x[2]) * cos(u[1]);
        xdot[1] = p[0] * u[0] * sin(x[2]) * cos(u[1]);
        xdot[2] = p[0] * u[0] * sin(u[1]) / d[0];
    }
    template <class T> PMB_HD void lagrange(const T*, const T*, const T*, const double*, double, T& L) const { L = T(0.0); }
    template <class T> PMB_HD void mayer(const T*, const T*, const T* p, const double*, double, T& M) const { M = p[0]; }
    template <class T> PMB_HD void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

struct Cstr {
    static constexpr int NX = 4, NU = 2, NP = 0, ND = 0, NG = 0, NPARAM = 16 + 4 + 16 + 4 + 2;
    double Q[16], R[4], P[16], xs[4], us[2];  // column-major dense, cstr_control_test.cpp:40-50
    void defaults()
    {
        for (int i = 0; i < 16; ++i) { Q[i] = 0; P[i] = 0; }
        for (int i = 0; i < 4; ++i) R[i] = 0;
        Q[0] = 0.2; Q[5] = 1.0; Q[10] = 0.5; Q[15] = 0.2;
        R[0] = 0.5; R[3] = 5.0 * 1.0e-7;
        const double p[16] = {1.4646778374584373, 0.6676889516721198, 0.35446715117028615, 0.10324422005086348,
                              0.6676889516721198, 1.407812935783267, 0.17788030743777067, 0.050059833257226405,
                              0.3544671511702861, 0.1778803074377706, 0.6336052592712396, 0.01110329497282364,
                              0.1032442200508634, 0.05005983325722643, 0.011103294972823655, 0.229412393739723};
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) P[r + 4 * c] = p[4 * r + c];  // comma initialiser is row by row
        xs[0] = 2.1402105301746182e00; xs[1] = 1.0903043613077321e00; xs[2] = 1.1419108442079495e02; xs[3] = 1.1290659291045561e02;
        us[0] = 14.19; us[1] = -1113.50;
    }
    void set_params(const double* v)
    { int k = 0; for (int i = 0; i < 16; ++i) Q[i] = v[k++]; for (int i = 0; i < 4; ++i) R[i] = v[k++]; for (int i = 0; i < 16; ++i) P[i] = v[k++];
      for (int i = 0; i < 4; ++i) xs[i] = v[k++]; for (int i = 0; i < 2; ++i) us[i] = v[k++]; }
    void get_params(double* v) const
    { int k = 0; for (int i = 0; i < 16; ++i) v[k++] = Q[i]; for (int i = 0; i < 4; ++i) v[k++] = R[i]; for (int i = 0; i < 16; ++i) v[k++] = P[i];
      for (int i = 0; i < 4; ++i) v[k++] = xs[i]; for (int i = 0; i < 2; ++i) v[k++] = us[i]; }

    /** cstr_control_test.cpp:63-100 */
    template <class T>
    PMB_HD void dynamics(const T* x, const T* u, const T*, const double*, const T&, T* xdot) const
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
        T k_1 = k10 * exp(E1 / (273.15 + x[2]));
        T k_2 = k20 * exp(E2 / (273.15 + x[2]));
        T k_3 = k30 * exp(E3 / (273.15 + x[2]));
        T TIMEUNITS_PER_HOUR = (T)3600.0;

        xdot[0] = (1 / TIMEUNITS_PER_HOUR) * (u[0] * (c_AO - x[0]) - k_1 * x[0] - k_3 * x[0] * x[0]);
        xdot[1] = (1 / TIMEUNITS_PER_HOUR) * (-u[0] * x[1] + k_1 * x[0] - k_2 * x[1]);
        xdot[2] = (1 / TIMEUNITS_PER_HOUR) * (u[0] * (v_0 - x[2]) + (k_w * A_R / (rho * C_P * V_R)) *
                                              (x[3] - x[2]) - (1 / (rho * C_P)) * (k_1 * x[0] * H_1 + k_2 * x[1] * H_2 + k_3 * x[0] * x[1] * H_3));
        xdot[3] = (1 / TIMEUNITS_PER_HOUR) * ((1 / (m_K * C_PK)) * (u[1] + k_w * A_R * (x[2] - x[3])));
    }
    /** cstr_control_test.cpp:102-107: (x - xs).dot(Q * (x - xs)) + (u - us).dot(R * (u - us)) */
    template <class T>
    PMB_HD void lagrange(const T* x, const T* u, const T*, const double*, double, T& L) const
    {
        T dx[4], du[2], Qdx[4], Rdu[2];
        for (int i = 0; i < 4; ++i) dx[i] = x[i] - xs[i];
        for (int i = 0; i < 2; ++i) du[i] = u[i] - us[i];
        for (int i = 0; i < 4; ++i) Qdx[i] = ((Q[i] * dx[0]) + (Q[i + 4] * dx[1])) + ((Q[i + 8] * dx[2]) + (Q[i + 12] * dx[3]));
        for (int i = 0; i < 2; ++i) Rdu[i] = (R[i] * du[0]) + (R[i + 2] * du[1]);
        L = (((dx[0] * Qdx[0]) + (dx[1] * Qdx[1])) + ((dx[2] * Qdx[2]) + (dx[3] * Qdx[3]))) + ((du[0] * Rdu[0]) + (du[1] * Rdu[1]));
    }
    /** cstr_control_test.cpp:109-113 */
    template <class T>
    PMB_HD void mayer(const T* x, const T*, const T*, const double*, double, T& M) const
    {
        T dx[4], Pdx[4];
        for (int i = 0; i < 4; ++i) dx[i] = x[i] - xs[i];
        for (int i = 0; i < 4; ++i) Pdx[i] = ((P[i] * dx[0]) + (P[i + 4] * dx[1])) + ((P[i + 8] * dx[2]) + (P[i + 12] * dx[3]));
        M = ((dx[0] * Pdx[0]) + (dx[1] * Pdx[1])) + ((dx[2] * Pdx[2]) + (dx[3] * Pdx[3]));
    }
    template <class T> PMB_HD void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

/** 13-state rigid-body kite: x = [v_b(3) | w_b(3) | r_ned(3) | q(4, scalar first)], u = [thrust, elevator, rudder],
 *  d = [wind speed along NED x].  Quadratic tracking cost.  Our own model (the reference ships none). */
struct Kite {
    static constexpr int NX = 13, NU = 3, NP = 0, ND = 1, NG = 0, NPARAM = 13 + 3 + 13 + 13 + 3;
    double Q[13], R[3], QN[13], xref[13], uref[3];
    void defaults()
    {
        const double q[13] = {0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.01, 0.01, 0.1, 0.5, 0.5, 0.5, 0.5};
        for (int i = 0; i < 13; ++i) { Q[i] = q[i]; QN[i] = 10.0 * q[i]; }
        R[0] = 0.01; R[1] = 1.0; R[2] = 1.0;
        const double xr[13] = {12.0, 0.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, -50.0, 1.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < 13; ++i) xref[i] = xr[i];
        uref[0] = 1.5; uref[1] = 0.0; uref[2] = 0.0;
    }
    void set_params(const double* v)
    { int k = 0; for (int i = 0; i < 13; ++i) Q[i] = v[k++]; for (int i = 0; i < 3; ++i) R[i] = v[k++]; for (int i = 0; i < 13; ++i) QN[i] = v[k++];
      for (int i = 0; i < 13; ++i) xref[i] = v[k++]; for (int i = 0; i < 3; ++i) uref[i] = v[k++]; }
    void get_params(double* v) const
    { int k = 0; for (int i = 0; i < 13; ++i) v[k++] = Q[i]; for (int i = 0; i < 3; ++i) v[k++] = R[i]; for (int i = 0; i < 13; ++i) v[k++] = QN[i];
      for (int i = 0; i < 13; ++i) v[k++] = xref[i]; for (int i = 0; i < 3; ++i) v[k++] = uref[i]; }

    template <class T>
    PMB_HD void dynamics(const T* x, const T* u, const T*, const double* d, const T&, T* xdot) const
    {
        const double mass = 2.5, Ixx = 0.25, Iyy = 0.12, Izz = 0.32, g = 9.81;
        const double rho_air = 1.2, Sref = 0.45, bref = 2.0, cref = 0.23;
        const double CL0 = 0.3, CLa = 4.5, CLde = 0.4, CD0 = 0.03, Kind = 0.05, CYb = -0.3;
        const double Cm0 = 0.02, Cma = -0.6, Cmq = -8.0, Cmde = -0.9;
        const double Clb = -0.06, Clp = -0.5, Cnb = 0.06, Cnr = -0.1, Cndr = -0.05, Cldr = 0.005;
        const double lam_q = 1.0;

        T vx = x[0], vy = x[1], vz = x[2];
        T wx = x[3], wy = x[4], wz = x[5];
        T q0 = x[9], q1 = x[10], q2 = x[11], q3 = x[12];

        // rotation body -> NED
        T r00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3);
        T r01 = 2.0 * (q1 * q2 - q0 * q3);
        T r02 = 2.0 * (q1 * q3 + q0 * q2);
        T r10 = 2.0 * (q1 * q2 + q0 * q3);
        T r11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3);
        T r12 = 2.0 * (q2 * q3 - q0 * q1);
        T r20 = 2.0 * (q1 * q3 - q0 * q2);
        T r21 = 2.0 * (q2 * q3 + q0 * q1);
        T r22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);

        // apparent wind in the body frame
        T vax = vx - r00 * d[0];
        T vay = vy - r01 * d[0];
        T vaz = vz - r02 * d[0];
        T Va2 = vax * vax + vay * vay + vaz * vaz + 1.0e-4;
        T Va = sqrt(Va2);
        T alpha = atan2(vaz, vax);
        T beta = vay / Va;
        T qbarS = (0.5 * rho_air * Sref) * Va2;

        T CL = CL0 + CLa * alpha + CLde * u[1];
        T CD = CD0 + Kind * CL * CL;
        T CY = CYb * beta;
        T ca = cos(alpha), sa = sin(alpha);

        T Fx = qbarS * (CL * sa - CD * ca) + u[0];
        T Fy = qbarS * CY;
        T Fz = qbarS * (-CD * sa - CL * ca);

        xdot[0] = Fx / mass + r20 * g - (wy * vz - wz * vy);
        xdot[1] = Fy / mass + r21 * g - (wz * vx - wx * vz);
        xdot[2] = Fz / mass + r22 * g - (wx * vy - wy * vx);

        T ph = (0.5 * bref) * wx / Va, qh = (0.5 * cref) * wy / Va, rh = (0.5 * bref) * wz / Va;
        T Lm = (qbarS * bref) * (Clb * beta + Clp * ph + Cldr * u[2]);
        T Mm = (qbarS * cref) * (Cm0 + Cma * alpha + Cmq * qh + Cmde * u[1]);
        T Nm = (qbarS * bref) * (Cnb * beta + Cnr * rh + Cndr * u[2]);
        xdot[3] = (Lm - (Izz - Iyy) * wy * wz) / Ixx;
        xdot[4] = (Mm - (Ixx - Izz) * wz * wx) / Iyy;
        xdot[5] = (Nm - (Iyy - Ixx) * wx * wy) / Izz;

        xdot[6] = r00 * vx + r01 * vy + r02 * vz;
        xdot[7] = r10 * vx + r11 * vy + r12 * vz;
        xdot[8] = r20 * vx + r21 * vy + r22 * vz;

        T nq = lam_q * (1.0 - (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3));
        xdot[9]  = 0.5 * (-q1 * wx - q2 * wy - q3 * wz) + nq * q0;
        xdot[10] = 0.5 * (q0 * wx + q2 * wz - q3 * wy) + nq * q1;
        xdot[11] = 0.5 * (q0 * wy - q1 * wz + q3 * wx) + nq * q2;
        xdot[12] = 0.5 * (q0 * wz + q1 * wy - q2 * wx) + nq * q3;
    }
    template <class T>
    PMB_HD void lagrange(const T* x, const T* u, const T*, const double*, double, T& L) const
    {
        T acc = Q[0] * ((x[0] - xref[0]) * (x[0] - xref[0]));
        for (int i = 1; i < 13; ++i) acc = acc + Q[i] * ((x[i] - xref[i]) * (x[i] - xref[i]));
        for (int i = 0; i < 3; ++i) acc = acc + R[i] * ((u[i] - uref[i]) * (u[i] - uref[i]));
        L = acc;
    }
    template <class T>
    PMB_HD void mayer(const T* x, const T*, const T*, const double*, double, T& M) const
    {
        T acc = QN[0] * ((x[0] - xref[0]) * (x[0] - xref[0]));
        for (int i = 1; i < 13; ++i) acc = acc + QN[i] * ((x[i] - xref[i]) * (x[i] - xref[i]));
        M = acc;
    }
    template <class T> PMB_HD void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

} // namespace pmb
