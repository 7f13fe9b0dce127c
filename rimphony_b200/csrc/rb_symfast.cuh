// rb_symfast.cuh -- j_I/Q/V and alpha_I/Q/V of one point on the product
// ("fast") path: the Symphony harmonic sum + (n, gamma) double integral of
//   src/symphony.rs:66-187 (compute), :196-295 (n_integration),
//   :312-389 (gamma_integral), :398-479 (gamma_integrand)
// evaluated with the compact engine of rb_engine.cuh.
//
// Same integrand, same integration domain, same truncation rule as the
// reference; what differs is how the quadrature is organised:
//
//   * all six integrands share every gamma node (one J_n, J_{n+1}, f, df pair
//     per node) and the two Stokes-V lobes fall out of the left/right halves;
//   * the gamma integral is seeded where the integrand lives.  For large n the
//     integrand is a peak of relative half-width ~ n^(-1/3) around gamma_peak:
//     with gamma = gamma_peak + t (gamma_+ - gamma_peak), z/n = 1 - t^2/2 - ... and
//     J_n(z)^2 ~ exp(-(2n/3) |t|^3).  The seeds are [-T, 0], [0, T] with
//     T = kPeakSpan n^(-1/3) plus the two outer remainders, instead of the
//     reference's bisection cascade from the full [gamma_-, gamma_+] (15-45 rule
//     applications per gamma integral at n >= 1e3, measured);
//   * the integral over continuous n marches in u = ln n with panels of fixed
//     width in u (n G(n) is a smooth bump in u: exponential rise, power-law
//     decay), instead of linear chunks grown x10 with a derivative probe each;
//     it stops by the reference's rule |chunk| < |sum| / 1e5 (symphony.rs:225).
#pragma once

#include "rb_bessel.cuh"
#include "rb_dist.cuh"
#include "rb_engine.cuh"

namespace rb {

constexpr double kPeakSpan = 3.2;    // seed half-width in units of n^(-1/3)
constexpr double kInnerFloor = 0.125; // acceptance floor of a gamma panel, fraction of the integral so far
constexpr double kPanelWidth = 2.302585092994046; // outer panel width in u = ln n (one decade)
constexpr int kMaxPanels = 64;

struct SymFastWS {
    EngLevel inner, outer;
    LeungOrder on, on1;
};

// Warp-uniform context of one point.
template <int KIND>
struct SymFastCtx {
    const Dist *d;
    SymFastWS *ws;
    double s, cos_th, sin_th;
    double epsrel_gamma;
};

// J_n(x) for the argument range of the Symphony integrand (0 <= x <= n): the
// Debye and Meissel expansions appear once each (pkgw_bessel_j, bessel.c:318-357).
RB_FN double leung_j_below(const LeungOrder &o, double x)
{
    if (o.kind != kOrderLeung || !(x <= o.n))
        return leung_j(o, x);
    const double eps = (o.n - x) / o.n;
    const bool use_debye = !(eps > o.hi_minus);
    const bool use_meissel = !(eps < o.lo_minus) && x != o.n;
    double dv = 0.0, mv = 0.0;
    if (use_debye)
        dv = leung_debye_eps(o.n, x);
    if (use_meissel)
        mv = leung_meissel_first(o, x);
    if (use_debye && use_meissel) {
        const double eta = log(eps) * kLog10e;
        const double pos = (eta - o.eta_lo_minus) / (kMinusEtaB - kMinusEtaA);
        return dv * (1.0 - pos) + mv * pos;
    }
    return use_debye ? dv : mv;
}

// The six gamma integrands at one node (symphony.rs:398-479).
template <int KIND>
RB_FN void sym_node(const SymFastCtx<KIND> &cx, double n, double gamma, double (&out)[6])
{
    const double s = cx.s, costh = cx.cos_th, sinth = cx.sin_th;
    const double beta = sqrt(1.0 - 1.0 / (gamma * gamma));
    const double cos_xi = (s * gamma - n) / (s * gamma * beta * costh);
    const double sin_xi = sqrt(1.0 - cos_xi * cos_xi);
    const double m = (costh - beta * cos_xi) / sinth;
    const double big_n = beta * sin_xi;

    double gamma_sin_xi;
    if (beta < 0.1) {
        gamma_sin_xi = gamma * sin_xi;
    } else {
        const double bc = beta * costh;
        const double beta2_costh2 = bc * bc;
        const double s_on_r = 2.0 * n / (s * (beta2_costh2 - 1.0));
        const double r = 1.0 - 1.0 / beta2_costh2;
        gamma_sin_xi = sqrt(r * (gamma * (gamma + s_on_r)) - (n * n / (s * s * beta2_costh2)));
    }
    const double z = s * beta * sinth * gamma_sin_xi;

    // J_n(z), J_{n+1}(z): one copy of the evaluator, two trips
    double jv[2];
#pragma unroll 1
    for (int k = 0; k < 2; k++)
        jv[k] = leung_j_below(k ? cx.ws->on1 : cx.ws->on, z);
    const double jn = jv[0];
    double djn;
    if (n >= 1e15)
        djn = NAN; // bessel.c:382-388
    else if (z == 0.0)
        djn = (n >= 2.0) ? 0.0 : ((n == 0.0) ? -jv[1] : n * jn / DBL_MIN - jv[1]);
    else
        djn = n * jn / z - jv[1];

    const double mj = m * jn;
    const double njp = big_n * djn;

    double f, dfdg, dfdcx;
    dist_eval<KIND>(*cx.d, gamma, cos_xi, f, dfdg, dfdcx);
    const double dfdcx_factor = (beta * costh - cos_xi) / (gamma - 1.0 / gamma);
    const double f_abs = dfdg + dfdcx_factor * dfdcx;

    const double g2 = gamma * gamma;
    const double pol_i = g2 * (mj * mj + njp * njp);
    const double pol_q = g2 * (mj * mj - njp * njp);
    const double pol_v = g2 * (2.0 * mj * njp);
    out[0] = pol_i * f;
    out[1] = pol_i * f_abs;
    out[2] = pol_q * f;
    out[3] = pol_q * f_abs;
    out[4] = pol_v * f;
    out[5] = pol_v * f_abs;
}

// G(n): the gamma integral at harmonic number n for all eight accumulators,
//   0 j_I, 1 a_I, 2 j_Q, 3 a_Q, 4 j_V(+), 5 a_V(+), 6 j_V(-), 7 a_V(-)
// ((+) = gamma > gamma_peak, symphony.rs:356-363), multiplied by wa / wb and
// parked in column `col` of the outer tile (A rows / B rows).
template <int KIND>
RB_FN_NOINLINE void sym_gamma_integral(Warp &w, const SymFastCtx<KIND> &cx, double n, int col, double wa, double wb)
{
    SymFastWS &ws = *cx.ws;
    const double s = cx.s, costh = cx.cos_th, sinth = cx.sin_th;
    const double nos = n / s;
    const double root = sqrt(nos * nos - sinth * sinth);
    const double sin2 = sinth * sinth;
    const double gamma_minus = (nos - fabs(costh) * root) / sin2;
    const double gamma_plus = (nos + fabs(costh) * root) / sin2;
    const double gamma_peak = 0.5 * (gamma_plus + gamma_minus);
    const double rel_width = (s < 1e6) ? 1.0 : exp(-0.27 * log(n) - 0.1);
    // gamma = gamma_peak + t half, t in [-1, 1]
    const double half = (gamma_plus - gamma_peak) * rel_width;

    warp_fence();
#ifdef RB_DEVICE_BUILD
    if (w.lane == 0)
#endif
    {
        leung_prepare(n, ws.on);
        leung_prepare(n + 1.0, ws.on1);
    }

    // seeds, pushed so that the two central panels are popped first
    PanelStack stk;
    stk.reset(&ws.inner);
    const double span = kPeakSpan * exp(-(1.0 / 3.0) * log(n)) / rel_width;
    if (span < 0.5) {
        stk.push(w, -1.0, -span, 0);
        stk.push(w, span, 1.0, 1);
        stk.push(w, -span, 0.0, 0);
        stk.push(w, 0.0, span, 1);
    } else {
        stk.push(w, -1.0, 0.0, 0);
        stk.push(w, 0.0, 1.0, 1);
    }
    stk.seal();

    // est: integral of |f| over the accepted panels; big: the largest |panel value| seen so far,
    // accepted or not (the central seeds are evaluated first, so it knows the scale at once)
    PerChan<double> sum_l, sum_r, est, big;
    RB_FOR_CHAN(c, kEngChan)
    {
        sum_l[c] = 0.0;
        sum_r[c] = 0.0;
        est[c] = 0.0;
        big[c] = 0.0;
    }

    while (stk.sp > 0) {
        double ta, tb;
        int side;
        stk.pop(ta, tb, side);
        const double tc = 0.5 * (ta + tb), thl = 0.5 * (tb - ta);
        warp_fence(); // the previous reduce has finished reading the tile

#ifdef RB_DEVICE_BUILD
        {
            double vals[6];
            sym_node<KIND>(cx, n, gamma_peak + half * (tc + thl * w.xk), vals);
            tile_store<6>(ws.inner.tile, w, w.lane, vals);
        }
#else
        for (int l = 0; l < 32; l++) {
            double vals[6];
            sym_node<KIND>(cx, n, gamma_peak + half * (tc + thl * LANE_X[l]), vals);
            tile_store<6>(ws.inner.tile, w, l, vals);
        }
#endif
        w.n_apply_lanes++;
        warp_fence();

        PerChan<double> r, e;
        tile_reduce(ws.inner.tile, 6, thl * half, r, e);

        PerChan<bool> ok;
        RB_FOR_CHAN(c, kEngChan) { ok[c] = true; }
        RB_FOR_CHAN(c, 6)
        {
            big[c] = fmax(big[c], fabs(r[c]));
            ok[c] = panel_ok(r[c], e[c], cx.epsrel_gamma, kInnerFloor * fmax(est[c] + fabs(r[c]), big[c]));
        }
        const bool accept = chan_all(ok, 6);
#ifdef RB_TRACE_INNER
        RB_TRACE_INNER(n, ta, tb, r, e, ok, est);
#endif
        if (accept || !stk.room(2) || panel_too_small(ta, tb)) {
            if (!accept)
                w.status |= kStatusCapHit;
            RB_FOR_CHAN(c, 6)
            {
                est[c] += fabs(r[c]);
                if (side)
                    sum_r[c] += r[c];
                else
                    sum_l[c] += r[c];
            }
        } else {
            // the half nearer the peak (t = 0) is popped first
            if (side) {
                stk.push(w, tc, tb, side);
                stk.push(w, ta, tc, side);
            } else {
                stk.push(w, ta, tc, side);
                stk.push(w, tc, tb, side);
            }
            stk.seal();
        }
    }

    // park the eight accumulators in the outer tile
    double *ot = ws.outer.tile;
#ifdef RB_DEVICE_BUILD
    if ((w.lane & 3) == 0)
#endif
    {
        RB_FOR_CHAN(c, 6)
        {
            if (c < 4) {
                const double v = sum_l[c] + sum_r[c];
                ot[c * kEngRow + col] = wa * v;
                ot[(kEngChan + c) * kEngRow + col] = wb * v;
            } else {
                ot[c * kEngRow + col] = wa * sum_r[c];
                ot[(kEngChan + c) * kEngRow + col] = wb * sum_r[c];
                ot[(c + 2) * kEngRow + col] = wa * sum_l[c];
                ot[(kEngChan + c + 2) * kEngRow + col] = wb * sum_l[c];
            }
        }
    }
}

RB_FN void tile_clear(const Warp &w, double *tile)
{
#ifdef RB_DEVICE_BUILD
    for (int i = w.lane; i < kEngTile; i += 32)
        tile[i] = 0.0;
#else
    (void)w;
    for (int i = 0; i < kEngTile; i++)
        tile[i] = 0.0;
#endif
}

// All six j/alpha coefficients of one point, dimensionless (lib.rs:178-191);
// out6 = j_I, a_I, j_Q, a_Q, j_V, a_V; lobes4 = j_V(+), j_V(-), a_V(+), a_V(-).
template <int KIND>
RB_FN void symphony_point_fast(Warp &w, const Dist &dist, double s, double theta, double epsrel_gamma,
                               double epsrel_n, SymFastWS &ws, double (&out6)[6], double (&lobes4)[4])
{
    constexpr double kNMax = 30.0, kTolerance = 1e5;

    SymFastCtx<KIND> cx;
    cx.d = &dist;
    cx.ws = &ws;
    cx.s = s;
    cx.cos_th = cos(theta);
    cx.sin_th = sin(theta);
    cx.epsrel_gamma = epsrel_gamma;

    const double n_minus = s * fabs(cx.sin_th);
    const long long n_lo = (long long)(n_minus + 1.0);
    const long long n_hi = (long long)(n_minus + 1.0 + kNMax);
    const double n_start = floor(n_minus + 1.0 + kNMax);

    // the first 30 harmonics, discretely (symphony.rs:96-108): a tile with unit weights
    warp_fence();
    tile_clear(w, ws.outer.tile);
    warp_fence();
    {
        int col = 0;
        for (long long n = n_lo; n < n_hi; n++, col++)
            sym_gamma_integral<KIND>(w, cx, (double)n, tile_col(col), 1.0, 0.0);
    }
    warp_fence();
    PerChan<double> ans, unused;
    tile_reduce(ws.outer.tile, kEngChan, 1.0, ans, unused);

    // the rest, treating n as continuous (symphony.rs:124-140), in u = ln n
    PerChan<bool> active;
    RB_FOR_CHAN(c, kEngChan) { active[c] = (ans[c] - ans[c] == 0.0); } // finite so far
    double u = log(n_start);
    PanelStack stk;

    for (int panel = 0; panel < kMaxPanels; panel++) {
        PerChan<double> chunk;
        RB_FOR_CHAN(c, kEngChan) { chunk[c] = 0.0; }
        stk.reset(&ws.outer);
        stk.push(w, u, u + kPanelWidth, 0);
        stk.seal();

        while (stk.sp > 0) {
            double ua, ub;
            int tag;
            stk.pop(ua, ub, tag);
            const double uc = 0.5 * (ua + ub), uhl = 0.5 * (ub - ua);
            warp_fence();
#pragma unroll 1
            for (int j = 0; j < 31; j++) {
                const double n = exp(uc + uhl * LANE_X[j]);
                sym_gamma_integral<KIND>(w, cx, n, tile_col(j), LANE_WK[j] * n, LANE_WD[j] * n);
            }
            warp_fence();
            PerChan<double> r, e;
            tile_reduce(ws.outer.tile, kEngChan, uhl, r, e);

            PerChan<bool> ok;
            RB_FOR_CHAN(c, kEngChan)
            {
                ok[c] = !active[c] || panel_ok(r[c], e[c], epsrel_n, fabs(ans[c] + chunk[c]));
            }
            const bool accept = chan_all(ok, kEngChan);
#ifdef RB_TRACE_FAST
            RB_TRACE_FAST(panel, ua, ub, r, e, ok, ans, chunk, w.n_apply_lanes);
#endif
            if (accept || !stk.room(2) || panel_too_small(ua, ub)) {
                if (!accept)
                    w.status |= kStatusCapHit;
                RB_FOR_CHAN(c, kEngChan) { chunk[c] += r[c]; }
            } else {
                stk.push(w, ua, uc, 0);
                stk.push(w, uc, ub, 0);
                stk.seal();
            }
        }

        PerChan<bool> done;
        RB_FOR_CHAN(c, kEngChan)
        {
            ans[c] += chunk[c];
            // loop condition of the reference; a NaN contribution also ends it
            if (!(fabs(chunk[c]) >= fabs(ans[c] / kTolerance)))
                active[c] = false;
            done[c] = !active[c];
        }
        if (chan_all(done, kEngChan))
            break;
        u += kPanelWidth;
        if (panel == kMaxPanels - 1)
            w.status |= kStatusCapHit;
    }

    // dimensional constants outside the integrals (symphony.rs:173-183)
    const double two_pi_e = kTwoPi * kElectronCharge;
    const double pre_j = two_pi_e * two_pi_e / (kSpeedLight * fabs(cx.cos_th));
    const double pre_a = -1.0 * two_pi_e * two_pi_e / (2.0 * kMassElectron * kSpeedLight * fabs(cx.cos_th));

    double total[kEngChan];
#pragma unroll
    for (int c = 0; c < kEngChan; c++)
        total[c] = chan_get(ans, c);

    out6[0] = total[0] * pre_j;
    out6[1] = total[1] * pre_a;
    out6[2] = total[2] * pre_j;
    out6[3] = total[3] * pre_a;
    lobes4[0] = total[4] * pre_j;
    lobes4[1] = total[6] * pre_j;
    lobes4[2] = total[5] * pre_a;
    lobes4[3] = total[7] * pre_a;
    out6[4] = lobes4[0] + lobes4[1];
    out6[5] = lobes4[2] + lobes4[3];
}

} // namespace rb
