// Host-side solver planner of the C ABI: pdm_solver_plan (include/pdm.h).
//
// Restates, in float32 and in the reference's operand order, what DPM_Solver.sample derives per step on the host:
//   NoiseScheduleVP('discrete')   dpm_solver_pp.py:55-169   (log-alpha table, interpolate_fn :9-52, lambda / inverse lambda)
//   get_time_steps / fast orders  dpm_solver_pp.py:330-405
//   singlestep 1S / 2S / 3S coefficients (data prediction, solver_type 'dpm_solver')   :432-457, :511-557, :700-766
//   multistep 1 / 2M / 3M coefficients                                                 :602-677, driver :995-1017
// Every solver scalar is data independent, so ONE flat table (PDM_PLAN_STRIDE floats per network evaluation) drives the
// whole device loop (pdm_sample).  The Python host layer (dpm_solver_pp.build_plan) builds the same table with torch CPU
// float32 ops; the two agree to a few float32 ulp, not bit for bit: torch evaluates exp / log / log1p / expm1 with SLEEF
// (<= 1 ulp) where libm is correctly rounded in ~99 % of the cases, and torch.linspace rounds per SIMD chunk (the result
// depends on the host's vector width).  tests/test_plan_cpu.py pins the difference (<= 2e-6 relative per coefficient).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pdm.h"
#include "common.cuh"

namespace pdm {
namespace {

// torch.linspace(start, end, n) for float32 in its scalar form: ascending from `start` in the first half, descending from
// `end` in the second (aten/src/ATen/native/cpu/RangeFactoriesKernel.cpp)
std::vector<float> linspace(float start, float end, int n) {
    std::vector<float> v(n);
    if (n == 1) {
        v[0] = start;
        return v;
    }
    const float step = (end - start) / (float)(n - 1);
    const int half = n / 2;
    for (int i = 0; i < n; ++i) v[i] = i < half ? start + step * (float)i : end - step * (float)(n - i - 1);
    return v;
}

struct Schedule {
    std::vector<float> t, la, la_rev, t_rev;  // knots t_i = i / N (ascending), log alpha(t_i); reversed copies
    explicit Schedule(const float* betas, int n) {
        // log_alphas = 0.5 * log(1 - betas).cumsum(0)  (dpm_solver_pp.py:101-103); torch's CPU cumsum accumulates in double
        t = linspace(1.f / (float)n, 1.f, n);
        la.resize(n);
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            acc += (double)logf(1.f - betas[i]);
            la[i] = 0.5f * (float)acc;
        }
        la_rev.assign(la.rbegin(), la.rend());
        t_rev.assign(t.rbegin(), t.rend());
    }
    // interpolate_fn (dpm_solver_pp.py:9-52) for one query: the segment whose right knot is the first knot >= x, clamped to
    // the end segments (linear extrapolation)
    static float pwl(float x, const std::vector<float>& xp, const std::vector<float>& yp) {
        const int K = (int)xp.size();
        int idx = 0;  // number of knots strictly below x
        {
            int lo = 0, hi = K;
            while (lo < hi) {
                const int mid = (lo + hi) / 2;
                if (xp[mid] < x) lo = mid + 1; else hi = mid;
            }
            idx = lo;
        }
        int lo = idx - 1;
        if (lo < 0) lo = 0;
        if (lo > K - 2) lo = K - 2;
        const float sx = xp[lo], ex = xp[lo + 1], sy = yp[lo], ey = yp[lo + 1];
        return sy + (x - sx) * (ey - sy) / (ex - sx);
    }
    float log_mean(float x) const { return pwl(x, t, la); }
    float alpha(float x) const { return expf(log_mean(x)); }
    float sigma(float x) const { return sqrtf(1.f - expf(2.f * log_mean(x))); }
    float lam(float x) const {
        const float lm = log_mean(x);
        return lm - 0.5f * logf(1.f - expf(2.f * lm));
    }
    float inv_lam(float l) const {
        // -0.5 * logaddexp(0, -2 l) -> interpolate on the flipped tables (dpm_solver_pp.py:162-164)
        const float a = 0.f, b = -2.f * l;
        const float m = a > b ? a : b;
        const float lae = m + log1pf(expf(-fabsf(a - b)));
        return pwl(-0.5f * lae, la_rev, t_rev);
    }
};

struct Rec {
    float v[PDM_PLAN_STRIDE];
    Rec() { std::memset(v, 0, sizeof(v)); }
};

Rec rec(const Schedule& ns, float t_eval, float n_time, float A, float B_img, bool has_c, float C_img, float B_msk, float C_msk,
        int stage, bool last) {
    Rec r;
    r.v[0] = t_eval * n_time;
    r.v[1] = ns.alpha(t_eval);
    r.v[2] = ns.sigma(t_eval);
    r.v[3] = A;
    r.v[4] = B_img;
    r.v[5] = has_c ? C_img : 0.f;
    r.v[6] = B_msk;
    r.v[7] = has_c ? C_msk : 0.f;
    r.v[8] = (float)stage;
    r.v[9] = has_c ? 1.f : 0.f;
    r.v[10] = last ? 1.f : 0.f;
    return r;
}
// pass-through mask stream (enable_mask_opt=False): m_out = A_msk * m_base + B_msk * P0, no difference term
Rec rec_split(Rec r, float A_msk, float B_msk) {
    r.v[6] = B_msk;
    r.v[7] = 0.f;
    r.v[11] = A_msk;
    r.v[12] = 1.f;
    return r;
}

// one singlestep update s -> t (mirrors dpm_solver_pp._step_records of the Python host layer)
void step_records(const Schedule& ns, float s, float t, int order, bool have_r, float r1, float r2, bool mask_opt,
                  float n_time, std::vector<Rec>& out) {
    const float lam_s = ns.lam(s), lam_t = ns.lam(t);
    const float h = lam_t - lam_s;
    const float sig_s = ns.sigma(s), sig_t = ns.sigma(t);
    const float a_t = expf(ns.log_mean(t));
    if (order == 1) {
        const float phi_1 = (expf(-h) - 1.f) / (-1.f);
        const float B = a_t * phi_1;
        Rec r = rec(ns, s, n_time, sig_t / sig_s, B, false, 0.f, B, 0.f, 0, true);
        out.push_back(mask_opt ? r : rec_split(r, 0.f, 1.f));
        return;
    }
    if (order == 2) {
        // default r1 = 0.5 is a Python double: 0.5 / r1 is evaluated in double (exactly 1) before it meets a float32 tensor
        const float r1f = have_r ? r1 : 0.5f;
        const float half_over_r1 = have_r ? 0.5f / r1 : 1.f;
        const float s1 = ns.inv_lam(lam_s + r1f * h);
        const float sig_s1 = ns.sigma(s1);
        const float a_s1 = expf(ns.log_mean(s1));
        const float phi_11 = expm1f(-r1f * h);
        const float phi_1 = expm1f(-h);
        const float B0 = a_s1 * phi_11;
        const float Bt = a_t * phi_1;
        const float Ct = half_over_r1 * (a_t * phi_1);
        Rec r0 = rec(ns, s, n_time, sig_s1 / sig_s, -B0, false, 0.f, B0, 0.f, 0, false);  // '+' quirk on the mask (:536-539)
        Rec r1r = rec(ns, s1, n_time, sig_t / sig_s, -Bt, true, -Ct, -Bt, -Ct, 1, true);
        out.push_back(mask_opt ? r0 : rec_split(r0, 1.f, 0.f));
        out.push_back(mask_opt ? r1r : rec_split(r1r, 0.f, 1.f));
        return;
    }
    // order 3; defaults r1 = 1/3, r2 = 2/3 are Python doubles: r2 / r1 = 2 and 1 / r2 = 1.5 are formed in double
    const float r1f = have_r ? r1 : (float)(1.0 / 3.0);
    const float r2f = have_r ? r2 : (float)(2.0 / 3.0);
    const float r2_over_r1 = have_r ? r2 / r1 : (float)((2.0 / 3.0) / (1.0 / 3.0));
    const float inv_r2 = have_r ? 1.f / r2 : (float)(1.0 / (2.0 / 3.0));
    const float s1 = ns.inv_lam(lam_s + r1f * h);
    const float s2 = ns.inv_lam(lam_s + r2f * h);
    const float sig_s1 = ns.sigma(s1), sig_s2 = ns.sigma(s2);
    const float a_s1 = expf(ns.log_mean(s1)), a_s2 = expf(ns.log_mean(s2));
    const float phi_11 = expm1f(-r1f * h);
    const float phi_12 = expm1f(-r2f * h);
    const float phi_1 = expm1f(-h);
    const float phi_22 = expm1f(-r2f * h) / (r2f * h) + 1.f;
    const float phi_2 = phi_1 / h + 1.f;
    const float B0 = a_s1 * phi_11;
    const float B1 = a_s2 * phi_12;
    const float C1 = r2_over_r1 * (a_s2 * phi_22);
    const float Bt = a_t * phi_1;
    const float Ct = inv_r2 * (a_t * phi_2);
    (void)phi_12;
    Rec q0 = rec(ns, s, n_time, sig_s1 / sig_s, -B0, false, 0.f, B0, 0.f, 0, false);  // '+' quirk (:730-733)
    Rec q1 = rec(ns, s1, n_time, sig_s2 / sig_s, -B1, true, C1, -B1, C1, 1, false);
    Rec q2 = rec(ns, s2, n_time, sig_t / sig_s, -Bt, true, Ct, -Bt, Ct, 2, true);
    out.push_back(mask_opt ? q0 : rec_split(q0, 1.f, 0.f));
    out.push_back(mask_opt ? q1 : rec_split(q1, 1.f, 0.f));
    out.push_back(mask_opt ? q2 : rec_split(q2, 0.f, 1.f));
}

// one multistep update from t_hist[o-1] to t with o cached predictions (dpm_solver_pp.py:432-446, 606-628, 649-669)
Rec multistep_record(const Schedule& ns, const float* t_hist, int o, float t, float n_time) {
    Rec r;
    const float p0 = t_hist[o - 1];
    const float lam0 = ns.lam(p0), lam_t = ns.lam(t);
    const float sig0 = ns.sigma(p0), sig_t = ns.sigma(t);
    const float a_t = expf(ns.log_mean(t));
    const float h = lam_t - lam0;
    r.v[0] = p0 * n_time;
    r.v[1] = ns.alpha(p0);
    r.v[2] = ns.sigma(p0);
    r.v[3] = sig_t / sig0;
    r.v[11] = (float)o;
    r.v[15] = 1.f;  // record kind: multistep
    if (o == 1) {
        r.v[4] = a_t * ((expf(-h) - 1.f) / (-1.f));
        return r;
    }
    const float lam1 = ns.lam(t_hist[o - 2]);
    const float h_0 = lam0 - lam1;
    const float r0 = h_0 / h;
    const float B = a_t * (expf(-h) - 1.f);
    r.v[4] = B;
    r.v[7] = 1.f / r0;
    r.v[12] = 0.5f * B;
    if (o == 3) {
        const float lam2 = ns.lam(t_hist[o - 3]);
        const float h_1 = lam1 - lam2;
        const float r1 = h_1 / h;
        r.v[5] = a_t * ((expf(-h) - 1.f) / h + 1.f);
        r.v[6] = a_t * ((expf(-h) - 1.f + h) / (h * h) - 0.5f);
        r.v[8] = 1.f / r1;
        r.v[9] = r0 / (r0 + r1);
        r.v[10] = 1.f / (r0 + r1);
    }
    return r;
}

std::vector<float> time_steps(const Schedule& ns, int skip_type, float t_T, float t_0, int N) {
    if (skip_type == PDM_SKIP_TIME_UNIFORM) return linspace(t_T, t_0, N + 1);
    if (skip_type == PDM_SKIP_LOGSNR) {
        std::vector<float> l = linspace(ns.lam(t_T), ns.lam(t_0), N + 1);
        for (auto& v : l) v = ns.inv_lam(v);
        return l;
    }
    // t2: linspace(t_T ** 0.5, t_0 ** 0.5, N + 1) ** 2 (the square roots are Python doubles)
    std::vector<float> q = linspace((float)sqrt((double)t_T), (float)sqrt((double)t_0), N + 1);
    for (auto& v : q) v = v * v;
    return q;
}

std::vector<int> fast_orders(int steps, int order) {
    std::vector<int> o;
    if (order == 3) {
        const int K = steps / 3 + 1, rem = steps % 3;
        const int ntail = rem == 0 ? 2 : 1;
        o.assign(K - ntail, 3);
        if (rem == 0) { o.push_back(2); o.push_back(1); }
        else if (rem == 1) o.push_back(1);
        else o.push_back(2);
    } else {
        o.assign(steps / 2, 2);
        if (steps % 2) o.push_back(1);
    }
    return o;
}

thread_local std::string g_plan_error;

}  // namespace

const char* plan_last_error() { return g_plan_error.c_str(); }

int solver_plan(const float* betas, int n_betas, int steps, int order, int method, int skip_type, float eps, float T,
                int mask_opt, float n_time, float* out, int cap, int* n_evals) {
    try {
        PDM_REQUIRE(betas && n_betas >= 2 && n_evals, "pdm_solver_plan: null / short beta table");
        PDM_REQUIRE(steps >= 1, "pdm_solver_plan: steps must be >= 1");
        PDM_REQUIRE(skip_type >= PDM_SKIP_TIME_UNIFORM && skip_type <= PDM_SKIP_T2, "pdm_solver_plan: bad skip_type");
        const Schedule ns(betas, n_betas);
        std::vector<Rec> recs;
        if (method == PDM_METHOD_FAST) {
            PDM_REQUIRE(order == 2 || order == 3, "order must >= 2");
            const std::vector<int> orders = fast_orders(steps, order);
            const std::vector<float> ts = time_steps(ns, skip_type, T, eps, steps);
            int i = 0;
            for (int o : orders) {
                const float h = ns.lam(ts[i + o]) - ns.lam(ts[i]);
                const float r1 = o <= 1 ? 0.f : (ns.lam(ts[i + 1]) - ns.lam(ts[i])) / h;
                const float r2 = o <= 2 ? 0.f : (ns.lam(ts[i + 2]) - ns.lam(ts[i])) / h;
                step_records(ns, ts[i], ts[i + o], o, true, r1, r2, mask_opt != 0, n_time, recs);
                i += o;
            }
        } else if (method == PDM_METHOD_SINGLESTEP) {
            PDM_REQUIRE(order >= 1 && order <= 3, "Solver order must be 1 or 2 or 3");
            const int n_steps = steps / order;
            PDM_REQUIRE(n_steps >= 1, "pdm_solver_plan: steps < order");
            const std::vector<float> ts = time_steps(ns, skip_type, T, eps, n_steps);
            for (int i = 0; i < n_steps; ++i) step_records(ns, ts[i], ts[i + 1], order, false, 0.f, 0.f, mask_opt != 0, n_time, recs);
        } else if (method == PDM_METHOD_MULTISTEP) {
            PDM_REQUIRE(order >= 1 && order <= 3, "Solver order must be 1 or 2 or 3");
            PDM_REQUIRE(steps >= order, "pdm_solver_plan: multistep needs steps >= order");
            const std::vector<float> ts = time_steps(ns, skip_type, T, eps, steps);
            for (int k = 0; k < steps; ++k) {
                const int o = k + 1 < order ? k + 1 : order;
                recs.push_back(multistep_record(ns, &ts[k - o + 1], o, ts[k + 1], n_time));
            }
        } else {
            PDM_REQUIRE(false, "pdm_solver_plan: unsupported method");
        }
        *n_evals = (int)recs.size();
        if (out) {
            PDM_REQUIRE(cap >= (int)recs.size(), "pdm_solver_plan: output capacity too small");
            for (size_t i = 0; i < recs.size(); ++i) std::memcpy(out + i * PDM_PLAN_STRIDE, recs[i].v, sizeof(recs[i].v));
        }
        return 0;
    } catch (const std::exception& e) {
        g_plan_error = e.what();
        return 1;
    }
}

}  // namespace pdm
