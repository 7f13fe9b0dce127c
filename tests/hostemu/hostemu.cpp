// tests/hostemu/hostemu.cpp -- DEVELOPMENT HARNESS, not part of the product.
//
// Compiles the device headers of rimphony_b200/csrc with g++ and
// -DRB_HOST_EMU, which turns the lane dimension of the warp-cooperative
// quadrature into an explicit loop.  It lets the algorithmic logic of the CUDA
// kernels (joint adaptive quadrature, Leung Bessel restructuring, fused
// integrands) be exercised and compared with the oracle on a machine without a
// GPU.  The package rimphony_b200 never loads this library; GPU parity tests
// (-m gpu) go through the real C ABI.
#include <cstdlib>
#include <cstring>
#include <vector>

// The chunk-growth vote of the Symphony product path (rb_symfast.cuh: delta_n grows when ALL active
// accumulators ask for it, whereas the reference grows it per coefficient, symphony.rs:243-258):
// count the votes and those in which the accumulators disagreed.
static long g_votes = 0, g_split_votes = 0;
#define RB_TRACE_VOTE(chunk_no, n_lo, delta_n, active, grow)                                   \
    do {                                                                                       \
        int na_ = 0, ng_ = 0;                                                                  \
        for (int c_ = 0; c_ < 8; c_++)                                                         \
            if (active.v[c_]) {                                                                \
                na_++;                                                                         \
                if (grow.v[c_])                                                                \
                    ng_++;                                                                     \
            }                                                                                  \
        g_votes++;                                                                             \
        if (ng_ != 0 && ng_ != na_)                                                            \
            g_split_votes++;                                                                   \
    } while (0)

#include "../../rimphony_b200/csrc/rb_symphony.cuh"
#include "../../rimphony_b200/csrc/rb_heyvaerts.cuh"
#include "../../rimphony_b200/csrc/rb_symfast.cuh"
#include "../../rimphony_b200/csrc/rb_heyfast.cuh"

using namespace rb;

extern "C" {

void emu_vote_stats(long *out2)
{
    out2[0] = g_votes;
    out2[1] = g_split_votes;
}

double emu_leung_j(double n, double x)
{
    LeungOrder o;
    leung_prepare(n, o);
    return leung_j(o, x);
}

double emu_leung_dj(double n, double x)
{
    LeungOrder o, o1;
    leung_prepare(n, o);
    leung_prepare(n + 1.0, o1);
    double j, dj;
    leung_j_and_dj(o, o1, x, j, dj);
    return dj;
}

} // extern "C"

template <int KIND>
static int make_dist(const double *params, int n_params, Dist &d, unsigned &status)
{
    if (!dist_from_params<KIND>(params, n_params, d))
        return -1;
    Warp w;
    w.init();
    std::vector<double> store(1024 * IntervalList<1>::doubles_per_interval);
    IntervalList<1> list;
    list.bind(store.data(), 1024);
    const bool ok = dist_normalize<KIND>(w, d, KIND == kDistThermalJuettner ? params[0] : 0.0, list);
    if (!ok)
        status |= kStatusNormFailed;
    status |= w.status;
    return 0;
}

template <int KIND, bool FUSED>
static void run_symphony(const Dist &d, double s, double theta, const double *eps, double *out6, double *lobes4, unsigned *info)
{
    constexpr int GC = 1024, NC = 1024;
    auto *ws = new SymWorkspace<FUSED, GC, NC>();
    Warp w;
    w.init();
    double o6[6], l4[4];
    symphony_point<KIND, FUSED, GC, NC>(w, d, s, theta, eps[0], eps[1], *ws, o6, l4);
    memcpy(out6, o6, sizeof(o6));
    memcpy(lobes4, l4, sizeof(l4));
    info[0] = w.n_apply_lanes;
    info[1] = w.status | ((unsigned)w.max_list << 8);
    delete ws;
}

template <int KIND>
static void run_symphony_fast(const Dist &d, double s, double theta, const double *eps, double *out6, double *lobes4, unsigned *info)
{
    auto *ws = new SymFastWS();
    Warp w;
    w.init();
    double o6[6], l4[4];
    symphony_point_fast<KIND>(w, d, s, theta, eps[0], eps[1], *ws, o6, l4);
    memcpy(out6, o6, sizeof(o6));
    memcpy(lobes4, l4, sizeof(l4));
    info[0] = w.n_apply_lanes;
    info[1] = w.status;
    if (w.status & kStatusRerouted) { // what the launcher does: hand the point to the faithful kernel
        if (ws->snap[kSnapValid] == 1.0) {
            constexpr int GC = 1024, NC = 1024;
            auto *fws = new SymWorkspace<false, GC, NC>();
            Warp w2;
            w2.init();
            symphony_tail_faithful<KIND, GC, NC>(w2, d, s, theta, eps[0], eps[1], *fws, ws->snap, o6, l4);
            memcpy(out6, o6, sizeof(o6));
            memcpy(lobes4, l4, sizeof(l4));
            info[0] += w2.n_apply_lanes;
            info[1] |= w2.status & 0xffu;
            delete fws;
        } else {
            unsigned info2[2];
            run_symphony<KIND, false>(d, s, theta, eps, out6, lobes4, info2);
            info[0] += info2[0];
            info[1] |= info2[1] & 0xffu;
        }
    }
    delete ws;
}

template <int KIND, bool FUSED>
static void run_heyvaerts(const Dist &d, double s, double theta, const double *eps, double *out2, unsigned *info)
{
    constexpr int IC = 1024, OC = 1024;
    auto *ws = new HeyWorkspace<FUSED, IC, OC>();
    Warp w;
    w.init();
    double o2[2];
    heyvaerts_point<KIND, FUSED, IC, OC>(w, d, s, theta, eps[2], eps[3], *ws, o2);
    out2[0] = o2[0];
    out2[1] = o2[1];
    info[0] = w.n_apply_lanes;
    info[1] = w.status | ((unsigned)w.max_list << 8);
    delete ws;
}

template <int KIND>
static void run_heyvaerts_fast(const Dist &d, double s, double theta, const double *eps, double *out2, unsigned *info)
{
    auto *ws = new HeyFastWS();
    Warp w;
    w.init();
    double o2[2];
    heyvaerts_point_fast<KIND>(w, d, s, theta, eps[2], eps[3], *ws, o2);
    out2[0] = o2[0];
    out2[1] = o2[1];
    info[0] = w.n_apply_lanes;
    info[1] = w.status;
    delete ws;
}

template <int KIND>
static int point_kind(const double *params, int n_params, int fused, int which, double s, double theta,
                      const double *eps, double *out8, double *lobes4, unsigned *info)
{
    Dist d;
    unsigned status = 0;
    if (make_dist<KIND>(params, n_params, d, status) != 0)
        return -1;
    info[0] = info[1] = info[2] = info[3] = 0;
    for (int c = 0; c < 8; c++)
        out8[c] = NAN;
    if (which & 1) {
        if (fused == 2)
            run_symphony_fast<KIND>(d, s, theta, eps, out8, lobes4, info);
        else if (fused)
            run_symphony<KIND, true>(d, s, theta, eps, out8, lobes4, info);
        else
            run_symphony<KIND, false>(d, s, theta, eps, out8, lobes4, info);
    }
    if (which & 2) {
        if (fused == 2)
            run_heyvaerts_fast<KIND>(d, s, theta, eps, out8 + 6, info + 2);
        else if (fused)
            run_heyvaerts<KIND, true>(d, s, theta, eps, out8 + 6, info + 2);
        else
            run_heyvaerts<KIND, false>(d, s, theta, eps, out8 + 6, info + 2);
    }
    info[1] |= status;
    info[3] |= 0;
    return 0;
}

extern "C" {

// which: bit0 symphony (out8[0..5]), bit1 heyvaerts (out8[6..7]).
// info: [0] symphony GK applications, [1] symphony status, [2] heyvaerts GK applications, [3] status
// eps: epsrel of {symphony gamma, symphony n, heyvaerts inner, heyvaerts outer}
int emu_point(int kind, const double *params, int n_params, int fused, int which, double s, double theta,
              const double *eps, double *out8, double *lobes4, unsigned *info)
{
    switch (kind) {
    case kDistPowerLaw:
        return point_kind<kDistPowerLaw>(params, n_params, fused, which, s, theta, eps, out8, lobes4, info);
    case kDistThermalJuettner:
        return point_kind<kDistThermalJuettner>(params, n_params, fused, which, s, theta, eps, out8, lobes4, info);
    case kDistPitchyPL:
        return point_kind<kDistPitchyPL>(params, n_params, fused, which, s, theta, eps, out8, lobes4, info);
    case kDistPitchyKappa:
        return point_kind<kDistPitchyKappa>(params, n_params, fused, which, s, theta, eps, out8, lobes4, info);
    }
    return -1;
}

double emu_norm(int kind, const double *params, int n_params)
{
    Dist d;
    unsigned status = 0;
    int rc = -1;
    switch (kind) {
    case kDistPowerLaw: rc = make_dist<kDistPowerLaw>(params, n_params, d, status); break;
    case kDistThermalJuettner: rc = make_dist<kDistThermalJuettner>(params, n_params, d, status); break;
    case kDistPitchyPL: rc = make_dist<kDistPitchyPL>(params, n_params, d, status); break;
    case kDistPitchyKappa: rc = make_dist<kDistPitchyKappa>(params, n_params, d, status); break;
    }
    return rc == 0 ? d.norm : NAN;
}

void emu_dist_eval(int kind, const double *params, int n_params, double gamma, double cos_xi, double *out3)
{
    Dist d;
    double f = NAN, a = NAN, b = NAN;
    switch (kind) {
    case kDistPowerLaw: dist_from_params<kDistPowerLaw>(params, n_params, d); d.norm = 1; dist_eval<kDistPowerLaw>(d, gamma, cos_xi, f, a, b); break;
    case kDistThermalJuettner: dist_from_params<kDistThermalJuettner>(params, n_params, d); d.norm = 1; dist_eval<kDistThermalJuettner>(d, gamma, cos_xi, f, a, b); break;
    case kDistPitchyPL: dist_from_params<kDistPitchyPL>(params, n_params, d); d.norm = 1; dist_eval<kDistPitchyPL>(d, gamma, cos_xi, f, a, b); break;
    case kDistPitchyKappa: dist_from_params<kDistPitchyKappa>(params, n_params, d); d.norm = 1; dist_eval<kDistPitchyKappa>(d, gamma, cos_xi, f, a, b); break;
    }
    out3[0] = f; out3[1] = a; out3[2] = b;
}

} // extern "C"

template <int KIND>
static double diag_kind(const double *params, int n_params, int coeff, int stokes, double s, double theta, int what,
                        double a, double b)
{
    Dist d;
    unsigned status = 0;
    if (make_dist<KIND>(params, n_params, d, status) != 0)
        return NAN;
    constexpr int GC = 1024, NC = 1024;
    auto *ws = new SymWorkspace<false, GC, NC>();
    Warp w;
    w.init();
    const double v = symphony_diagnostic<KIND, GC, NC>(w, d, coeff, stokes, s, theta, what, a, b, 1e-3, 1e-3, *ws);
    delete ws;
    return v;
}

extern "C" double emu_symphony_diag(int kind, const double *params, int n_params, int coeff, int stokes, double s,
                                    double theta, int what, double a, double b)
{
    switch (kind) {
    case kDistPowerLaw: return diag_kind<kDistPowerLaw>(params, n_params, coeff, stokes, s, theta, what, a, b);
    case kDistThermalJuettner: return diag_kind<kDistThermalJuettner>(params, n_params, coeff, stokes, s, theta, what, a, b);
    case kDistPitchyPL: return diag_kind<kDistPitchyPL>(params, n_params, coeff, stokes, s, theta, what, a, b);
    case kDistPitchyKappa: return diag_kind<kDistPitchyKappa>(params, n_params, coeff, stokes, s, theta, what, a, b);
    }
    return NAN;
}

// --- building blocks of the product path, one at a time (tests/test_device_headers_on_host.py) -----------
// what: 0 rb_exp, 1 rb_log, 2 rb_rcp, 3 rb_sqrt (out[0]); 4 rb_div(a, b)
extern "C" double emu_lean_math(int what, double a, double b)
{
    switch (what) {
    case 0: return rb_exp(a);
    case 1: return rb_log(a);
    case 2: return rb_rcp(a);
    case 3: return rb_sqrt(a);
    default: return rb_div(a, b);
    }
}

// J_n(x), J_{n+1}(x) twice: out4 = {single n, single n + 1, pair n, pair n + 1}.  Integer orders below 29 go
// through the Miller recurrences, orders >= 30 through pkgw_bessel_j and the shared-coefficient Debye pair.
extern "C" void emu_bessel_pair(double n, double x, double *out4)
{
    LeungOrder o0, o1;
    leung_prepare(n, o0);
    leung_prepare(n + 1.0, o1);
    out4[0] = leung_j(o0, x);
    out4[1] = leung_j(o1, x);
    if (o0.kind == kOrderInteger && o1.kind == kOrderInteger) {
        bessel_jn_pair_small_int(o0.nint, x, out4[2], out4[3]);
    } else {
        double jv[2];
        leung_j_pair_below(o0, o1, x, jv);
        out4[2] = jv[0];
        out4[3] = jv[1];
    }
}

// J_sigma, J_{sigma-1}, Y_sigma, Y_{sigma-1} at x: out8 = {plain x 4, prepared-order x 4}
extern "C" void emu_jy_pair(double sigma, double x, double *out8)
{
    bessel_jy_pair(sigma, x, out8[0], out8[1], out8[2], out8[3]);
    JYOrder o;
    jy_prepare(sigma, o);
    bessel_jy_pair_prepared(o, x, out8[4], out8[5], out8[6], out8[7]);
}

// I_{1/3}, I_{-1/3}, I_{2/3}, I_{-2/3} at g
extern "C" void emu_i_thirds(double g, double *out4) { bessel_i_thirds(g, out4[0], out4[1], out4[2], out4[3]); }
