// oracle/models.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// Problem functors of the three BASELINE model families, restated with plain pointers.
//   RobotModel : reference tests/control/mpc_wrapper_test.cpp:38-80 (mobile robot, NX=3 NU=2 ND=1)
//   CstrModel  : reference tests/control/cstr_control_test.cpp:34-113 (CSTR, NX=4 NU=2)
//   KiteModel  : NOT in the reference; our own 13-state rigid-body kite, defined identically in
//                polympc_b200/csrc/problems/kite.hpp.  Parity for it is oracle-vs-kernel only ("parity unpinned").
// Fixed-size products/dots use Eigen's unrolled halving order (canon.hpp), matching the product's Eigen shim.
#pragma once
#include "ad.hpp"

namespace orc {

template <class T, int n, class F> inline T hsum(F&& term) { return sum_halving<T>(0, n, term); }

/** x.dot(y) for fixed-size vectors: sum_i x_i*y_i (halving order) */
template <class T, int n, class A, class B> inline T dotn(const A& a, const B& b)
{ return hsum<T, n>([&](int i) { return a[i] * b[i]; }); }

// =====================================================================================================================
struct RobotModel {
    static constexpr int NX = 3, NU = 2, NP = 0, ND = 1, NG = 0;
    static constexpr int NPARAM = 8;
    double Q[3] = {1, 1, 1};   // DiagonalMatrix Q{1,1,1}
    double R[2] = {1, 1};
    double QN[3] = {1, 1, 1};  // present in the reference class but unused by its mayer term
    void set_params(const double* v) { for (int i = 0; i < 3; ++i) Q[i] = v[i]; R[0] = v[3]; R[1] = v[4]; for (int i = 0; i < 3; ++i) QN[i] = v[5 + i]; }
    void get_params(double* v) const { for (int i = 0; i < 3; ++i) v[i] = Q[i]; v[3] = R[0]; v[4] = R[1]; for (int i = 0; i < 3; ++i) v[5 + i] = QN[i]; }

    template <class T>
    void dynamics(const T* x, const T* u, const T*, const double* d, const T&, T* xdot) const
    {
        xdot[0] = u[0] * cos(x[2]) * cos(u[1]);
        xdot[1] = u[0] * sin(x[2]) * cos(u[1]);
        xdot[2] = u[0] * sin(u[1]) / d[0];
    }
    /** lagrange = x.dot(Qm * x) + u.dot(Rm * u), Qm = Q.toDenseMatrix().cast<T>() (dense 3x3 with explicit zeros) */
    template <class T>
    void lagrange(const T* x, const T* u, const T*, const double*, double, T& L) const
    {
        T Qx[3], Ru[2];
        for (int i = 0; i < 3; ++i) Qx[i] = hsum<T, 3>([&](int j) { return T(i == j ? Q[i] : 0.0) * x[j]; });
        for (int i = 0; i < 2; ++i) Ru[i] = hsum<T, 2>([&](int j) { return T(i == j ? R[i] : 0.0) * u[j]; });
        L = dotn<T, 3>(x, Qx) + dotn<T, 2>(u, Ru);
    }
    template <class T>
    void mayer(const T* x, const T*, const T*, const double*, double, T& M) const
    {
        T Qx[3];
        for (int i = 0; i < 3; ++i) Qx[i] = hsum<T, 3>([&](int j) { return T(i == j ? Q[i] : 0.0) * x[j]; });
        M = dotn<T, 3>(x, Qx);
    }
    template <class T> void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

// =====================================================================================================================
/** mobile robot with one generic inequality constraint per node (NG = 1): squared distance to a disc-shaped obstacle,
 *  lbg <= g(x) <= ubg.  Not a reference test case; it exercises the NG > 0 paths of the transcription
 *  (continuous_ocp.hpp:546-575, 769-782, 2150-2157) and of the SQP (sqp_base.hpp:425-443, 457-465, 588-593). */
struct RobotObstacleModel : RobotModel {
    static constexpr int NG = 1;
    template <class T>
    void ineq(const T* x, const T*, const T*, const double*, double, T* g) const
    {
        g[0] = (x[0] - 0.25) * (x[0] - 0.25) + (x[1] - 0.25) * (x[1] - 0.25);
    }
};

// =====================================================================================================================
/** reference tests/control/minimal_time_test.cpp:32-63 (ParkingOCP): NX=3 NU=2 NP=1 ND=1, free final time p(0) */
struct ParkingModel {
    static constexpr int NX = 3, NU = 2, NP = 1, ND = 1, NG = 0;
    static constexpr int NPARAM = 1;
    double unused = 0;
    void set_params(const double* v) { unused = v[0]; }
    void get_params(double* v) const { v[0] = unused; }
    template <class T>
    void dynamics(const T* x, const T* u, const T* p, const double* d, const T&, T* xdot) const
    {
        xdot[0] = p[0] * u[0] * cos(x[2]) * cos(u[1]);
        xdot[1] = p[0] * u[0] * sin(x[2]) * cos(u[1]);
        xdot[2] = p[0] * u[0] * sin(u[1]) / d[0];
    }
    template <class T> void lagrange(const T*, const T*, const T*, const double*, double, T& L) const { L = T(0.0); }
    template <class T> void mayer(const T*, const T*, const T* p, const double*, double, T& M) const { M = p[0]; }
    template <class T> void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

// =====================================================================================================================
struct CstrModel {
    static constexpr int NX = 4, NU = 2, NP = 0, ND = 0, NG = 0;
    static constexpr int NPARAM = 16 + 4 + 16 + 4 + 2;
    double Q[16], R[4], P[16], xs[4], us[2];  // column-major dense
    CstrModel()
    {
        for (double& v : Q) v = 0; for (double& v : R) v = 0;
        const double qd[4] = {0.2, 1.0, 0.5, 0.2};
        for (int i = 0; i < 4; ++i) Q[i + 4 * i] = qd[i];
        R[0] = 0.5; R[3] = 5.0 * 1.0e-7;
        const double p[16] = {1.4646778374584373, 0.6676889516721198, 0.35446715117028615, 0.10324422005086348,
                              0.6676889516721198, 1.407812935783267, 0.17788030743777067, 0.050059833257226405,
                              0.3544671511702861, 0.1778803074377706, 0.6336052592712396, 0.01110329497282364,
                              0.1032442200508634, 0.05005983325722643, 0.011103294972823655, 0.229412393739723};
        // the reference fills P row by row with the comma initialiser: P(r,c) = p[4*r + c]
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) P[r + 4 * c] = p[4 * r + c];
        xs[0] = 2.1402105301746182e00; xs[1] = 1.0903043613077321e00; xs[2] = 1.1419108442079495e02; xs[3] = 1.1290659291045561e02;
        us[0] = 14.19; us[1] = -1113.50;
    }
    void set_params(const double* v)
    { int k = 0; for (double& a : Q) a = v[k++]; for (double& a : R) a = v[k++]; for (double& a : P) a = v[k++]; for (double& a : xs) a = v[k++]; for (double& a : us) a = v[k++]; }
    void get_params(double* v) const
    { int k = 0; for (double a : Q) v[k++] = a; for (double a : R) v[k++] = a; for (double a : P) v[k++] = a; for (double a : xs) v[k++] = a; for (double a : us) v[k++] = a; }

    template <class T>
    void dynamics(const T* x, const T* u, const T*, const double*, const T&, T* xdot) const
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
    /** (x - xs).dot(Q * (x - xs)) + (u - us).dot(R * (u - us)) with double matrices and T vectors */
    template <class T>
    void lagrange(const T* x, const T* u, const T*, const double*, double, T& L) const
    {
        T dx[4], du[2], Qdx[4], Rdu[2];
        for (int i = 0; i < 4; ++i) dx[i] = x[i] - xs[i];
        for (int i = 0; i < 2; ++i) du[i] = u[i] - us[i];
        for (int i = 0; i < 4; ++i) Qdx[i] = hsum<T, 4>([&](int j) { return Q[i + 4 * j] * dx[j]; });
        for (int i = 0; i < 2; ++i) Rdu[i] = hsum<T, 2>([&](int j) { return R[i + 2 * j] * du[j]; });
        L = dotn<T, 4>(dx, Qdx) + dotn<T, 2>(du, Rdu);
    }
    template <class T>
    void mayer(const T* x, const T*, const T*, const double*, double, T& M) const
    {
        T dx[4], Pdx[4];
        for (int i = 0; i < 4; ++i) dx[i] = x[i] - xs[i];
        for (int i = 0; i < 4; ++i) Pdx[i] = hsum<T, 4>([&](int j) { return P[i + 4 * j] * dx[j]; });
        M = dotn<T, 4>(dx, Pdx);
    }
    template <class T> void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

// =====================================================================================================================
/** 13-state rigid-body kite (our model; see polympc_b200/csrc/problems/kite.hpp for the product-side definition).
 *  x = [v_b(3) | w_b(3) | r_ned(3) | q(4, scalar first)],  u = [thrust, elevator, rudder],  d = [wind speed along NED x] */
struct KiteModel {
    static constexpr int NX = 13, NU = 3, NP = 0, ND = 1, NG = 0;
    static constexpr int NPARAM = 13 + 3 + 13 + 13 + 3;
    double Q[13], R[3], QN[13], xref[13], uref[3];
    KiteModel()
    {
        const double q[13] = {0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.01, 0.01, 0.1, 0.5, 0.5, 0.5, 0.5};
        for (int i = 0; i < 13; ++i) { Q[i] = q[i]; QN[i] = 10.0 * q[i]; }
        R[0] = 0.01; R[1] = 1.0; R[2] = 1.0;
        const double xr[13] = {12.0, 0.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, -50.0, 1.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < 13; ++i) xref[i] = xr[i];
        uref[0] = 1.5; uref[1] = 0.0; uref[2] = 0.0;
    }
    void set_params(const double* v)
    { int k = 0; for (double& a : Q) a = v[k++]; for (double& a : R) a = v[k++]; for (double& a : QN) a = v[k++]; for (double& a : xref) a = v[k++]; for (double& a : uref) a = v[k++]; }
    void get_params(double* v) const
    { int k = 0; for (double a : Q) v[k++] = a; for (double a : R) v[k++] = a; for (double a : QN) v[k++] = a; for (double a : xref) v[k++] = a; for (double a : uref) v[k++] = a; }

    template <class T>
    void dynamics(const T* x, const T* u, const T*, const double* d, const T&, T* xdot) const
    {
        // airframe constants
        const double mass = 2.5, Ixx = 0.25, Iyy = 0.12, Izz = 0.32, g = 9.81;
        const double rho_air = 1.2, Sref = 0.45, bref = 2.0, cref = 0.23;
        const double CL0 = 0.3, CLa = 4.5, CLde = 0.4, CD0 = 0.03, Kind = 0.05, CYb = -0.3;
        const double Cm0 = 0.02, Cma = -0.6, Cmq = -8.0, Cmde = -0.9;
        const double Clb = -0.06, Clp = -0.5, Cnb = 0.06, Cnr = -0.1, Cndr = -0.05, Cldr = 0.005;
        const double lam_q = 1.0;

        T vx = x[0], vy = x[1], vz = x[2];
        T wx = x[3], wy = x[4], wz = x[5];
        T q0 = x[9], q1 = x[10], q2 = x[11], q3 = x[12];

        // rotation matrix body -> NED (row-major entries)
        T r00 = 1.0 - 2.0 * (q2 * q2 + q3 * q3);
        T r01 = 2.0 * (q1 * q2 - q0 * q3);
        T r02 = 2.0 * (q1 * q3 + q0 * q2);
        T r10 = 2.0 * (q1 * q2 + q0 * q3);
        T r11 = 1.0 - 2.0 * (q1 * q1 + q3 * q3);
        T r12 = 2.0 * (q2 * q3 - q0 * q1);
        T r20 = 2.0 * (q1 * q3 - q0 * q2);
        T r21 = 2.0 * (q2 * q3 + q0 * q1);
        T r22 = 1.0 - 2.0 * (q1 * q1 + q2 * q2);

        // apparent wind in body frame: va = v - R^T [W,0,0]
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

        // v_dot = F/m + R^T [0,0,g] - w x v
        xdot[0] = Fx / mass + r20 * g - (wy * vz - wz * vy);
        xdot[1] = Fy / mass + r21 * g - (wz * vx - wx * vz);
        xdot[2] = Fz / mass + r22 * g - (wx * vy - wy * vx);

        // moments
        T ph = (0.5 * bref) * wx / Va, qh = (0.5 * cref) * wy / Va, rh = (0.5 * bref) * wz / Va;
        T Lm = (qbarS * bref) * (Clb * beta + Clp * ph + Cldr * u[2]);
        T Mm = (qbarS * cref) * (Cm0 + Cma * alpha + Cmq * qh + Cmde * u[1]);
        T Nm = (qbarS * bref) * (Cnb * beta + Cnr * rh + Cndr * u[2]);
        xdot[3] = (Lm - (Izz - Iyy) * wy * wz) / Ixx;
        xdot[4] = (Mm - (Ixx - Izz) * wz * wx) / Iyy;
        xdot[5] = (Nm - (Iyy - Ixx) * wx * wy) / Izz;

        // r_dot = R v
        xdot[6] = r00 * vx + r01 * vy + r02 * vz;
        xdot[7] = r10 * vx + r11 * vy + r12 * vz;
        xdot[8] = r20 * vx + r21 * vy + r22 * vz;

        // q_dot = 0.5 q (x) [0,w] + lam (1 - |q|^2) q
        T nq = lam_q * (1.0 - (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3));
        xdot[9]  = 0.5 * (-q1 * wx - q2 * wy - q3 * wz) + nq * q0;
        xdot[10] = 0.5 * (q0 * wx + q2 * wz - q3 * wy) + nq * q1;
        xdot[11] = 0.5 * (q0 * wy - q1 * wz + q3 * wx) + nq * q2;
        xdot[12] = 0.5 * (q0 * wz + q1 * wy - q2 * wx) + nq * q3;
    }
    template <class T>
    void lagrange(const T* x, const T* u, const T*, const double*, double, T& L) const
    {
        T acc = Q[0] * ((x[0] - xref[0]) * (x[0] - xref[0]));
        for (int i = 1; i < 13; ++i) acc = acc + Q[i] * ((x[i] - xref[i]) * (x[i] - xref[i]));
        for (int i = 0; i < 3; ++i) acc = acc + R[i] * ((u[i] - uref[i]) * (u[i] - uref[i]));
        L = acc;
    }
    template <class T>
    void mayer(const T* x, const T*, const T*, const double*, double, T& M) const
    {
        T acc = QN[0] * ((x[0] - xref[0]) * (x[0] - xref[0]));
        for (int i = 1; i < 13; ++i) acc = acc + QN[i] * ((x[i] - xref[i]) * (x[i] - xref[i]));
        M = acc;
    }
    template <class T> void ineq(const T*, const T*, const T*, const double*, double, T*) const {}
};

} // namespace orc
