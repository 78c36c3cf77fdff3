// pmb_cta.hpp — block-cooperative vocabulary: one CTA (1..8 warps) works on one OCP/QP instance.
//
// Conventions that keep the arithmetic identical for every block size (and identical to the CPU oracle):
//   * element-wise work is strided over all threads of the block (order free);
//   * order-sensitive reductions keep their canonical shape (oracle/canon.hpp): "tree32" sums are computed by warp 0
//     (lane l accumulates elements l, l+32, ... sequentially, then an xor butterfly) and broadcast through shared memory;
//   * max / inf-norm reductions are exact and therefore free to use the whole block;
//   * scalars returned by a cooperative routine are valid in EVERY thread (uniform control flow follows from that).
#pragma once
#include "pmb_warp.hpp"

namespace pmb {

struct Cta {
    static constexpr int BC_SLOTS = 16;       // broadcast ring: a slot is reused only after BC_SLOTS further block syncs
    static constexpr int RED_DOUBLES = 8 * 12; // per-warp partials of up to 12 simultaneous max-reductions, <= 8 warps
    static constexpr int SCRATCH_DOUBLES = BC_SLOTS + RED_DOUBLES;
    Warp w;
    double* bc;      // [BC_SLOTS] broadcast slots followed by [RED_DOUBLES] reduction scratch (shared memory)
    int ctr;
    int vw;          // virtual warp id = (physical warp id - rot) mod nwarps

    /** `rot` rotates the warp numbering: the serial phases of the algorithms run on *virtual* warp 0, and the hardware
     *  pins physical warp w of every CTA to scheduler w % 4 — without the rotation the serial warps of all CTAs resident on
     *  an SM would share one scheduler while the other three idle.  Callers pass a different `rot` to co-resident CTAs. */
    PMB_DEV Cta(const Warp& w_, double* scratch, int rot = 0) : w(w_), bc(scratch), ctr(0)
    {
        const int nw = (w.nthreads() + 31) >> 5;
        int v = w.warp_id() - (rot % nw);
        vw = v < 0 ? v + nw : v;
    }
    PMB_DEV int tid() const { return (vw << 5) | w.lane(); }
    PMB_DEV int nthreads() const { return w.nthreads(); }
    PMB_DEV int lane() const { return w.lane(); }
    PMB_DEV int warp_id() const { return vw; }
    PMB_DEV int nwarps() const { return (w.nthreads() + 31) >> 5; }
    PMB_DEV void sync() const { w.block_sync(); }

    /** value held by thread `src` -> every thread (one block sync) */
    PMB_DEV double bcast(double v, int src = 0)
    {
        double* slot = bc + (ctr & (BC_SLOTS - 1));
        ++ctr;
        if (tid() == src) *slot = v;
        w.block_sync();
        return *slot;
    }
    PMB_DEV int bcast_int(int v, int src = 0) { return (int)bcast((double)v, src); }

    /** exact maximum over the block of K values per thread (m[k] >= identity for idle threads); result in every thread */
    template <int K>
    PMB_DEV void max_all(double (&m)[K])
    {
        PMB_UNROLL
        for (int k = 0; k < K; ++k)
            for (int off = 16; off >= 1; off >>= 1) { const double o = w.shfl_xor(m[k], off); if (o > m[k]) m[k] = o; }
        const int nw = nwarps();
        if (nw == 1) return;
        double* red = bc + BC_SLOTS;
        if (w.lane() == 0) for (int k = 0; k < K; ++k) red[vw * K + k] = m[k];
        w.block_sync();
        PMB_UNROLL
        for (int k = 0; k < K; ++k) {
            double v = red[k];
            for (int q = 1; q < nw; ++q) { const double o = red[q * K + k]; if (o > v) v = o; }
            m[k] = v;
        }
        w.block_sync();
    }
    PMB_DEV double max_all1(double v) { double m[1] = {v}; max_all<1>(m); return m[0]; }
};

/** canonical "tree32" sum of per-element terms term(i), i < n: computed by warp 0, broadcast to the block */
template <class F>
PMB_DEV double sum_tree32(Cta& c, int n, F term)
{
    double acc = 0.0;
    if (c.warp_id() == 0) {
        for (int i = c.lane(); i < n; i += 32) acc = term(i, acc);
    }
    for (int off = 16; off >= 1; off >>= 1) acc = acc + c.w.shfl_xor(acc, off);
    return c.bcast(acc, 0);
}

} // namespace pmb
