// CPU check of include/deb_ensemble.hpp (the C++ host mirror of the crate's builder API): built against tests/support/abi_on_oracle.cpp,
// i.e. with the CPU oracle behind the C ABI.  Exit code 0 = every check passed.
#include <cmath>
#include <cstdio>

#include "deb_ensemble.hpp"

static int failures = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); failures++; } \
    } while (0)

int main() {
    using deb::CrossingDirection;
    using deb::EnsembleIVP;
    using deb::ExplicitRungeKutta;
    using deb::System;
    const double PI = 3.14159265358979323846;
    {   // t_eval: SURVEY.md appendix A, L100 / L10 t_eval known answers for the (1, 1, 1) Lorenz trajectory
        const int n = 8;
        std::vector<double> y0(3 * n);
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) y0[3 * i + c] = 1.0 + 1e-3 * i * (c + 1);
        auto sol = EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 100.0, y0).t_eval({1.0, 2.5, 100.0})
                       .method(ExplicitRungeKutta::dopri5().rtol(1e-8)).solve();
        const deb::Solution s = sol.at(0);
        CHECK(s.status == deb::Status::Complete && s.steps.accepted == 6008 && s.steps.rejected == 411 && s.evals.function == 44525);
        CHECK(s.t.size() == 3 && s.t[0] == 1.0 && s.t[1] == 2.5 && s.t[2] == 100.0);
        CHECK(s.y[0][0] == -9.378567031548476 && s.y[1][2] == 24.70312694098664);
        CHECK(s.y[2][0] == -0x1.8c65ec78fb24fp+1 && s.y[2][1] == -0x1.4d4949b09d281p+1 && s.y[2][2] == 0x1.5bf4de76ed043p+4);
        CHECK(sol.ok(n - 1) && sol.at(n - 1).t.size() == 3);
    }
    {   // even(dt) + terminal event, per-trajectory parameters (examples/ode/03_logistic_growth as a sweep)
        const int n = 40;
        std::vector<double> y0(n, 1.0), km(2 * n);
        for (int i = 0; i < n; i++) { km[2 * i] = 1.0; km[2 * i + 1] = 10.0 + 0.25 * i; }
        auto sol = EnsembleIVP::ode(System::logistic_equation(1.0, 10.0).sweep(km), 0.0, 10.0, y0).even(2.0)
                       .event(deb::Event::linear(-9.0, 0.0, {1.0}).terminal()).method(ExplicitRungeKutta::dopri5().rtol(1e-10).atol(1e-10)).solve();
        for (int i = 0; i < n; i++) {
            const deb::Solution s = sol.at(i);
            const double m = km[2 * i + 1], t_star = std::log(9.0 * (m - 1.0) / (m - 9.0));
            CHECK(s.status == deb::Status::Interrupted);
            CHECK(s.t.size() >= 2 && s.t[0] == 0.0 && s.t[1] == 2.0);
            CHECK(std::fabs(s.t.back() - t_star) < 1e-2 && std::fabs(s.y.back()[0] - 9.0) < 5e-2);
            for (size_t r = 1; r < s.t.size(); r++) CHECK(s.t[r] > s.t[r - 1]);
        }
        // even(dt) alone: rows at 0, 2, ..., 10 (the last one is the exact final state)
        auto ev = EnsembleIVP::ode(System::logistic_equation(1.0, 10.0), 0.0, 10.0, {1.0}).even(2.0).method(ExplicitRungeKutta::dopri5()).solve();
        const deb::Solution s = ev.at(0);
        CHECK(s.t.size() == 6 && s.t[0] == 0.0 && s.t[5] == 10.0 && s.y[0][0] == 1.0 && std::fabs(s.y[5][0] - 10.0 / (1.0 + 9.0 * std::exp(-10.0))) < 1e-4);
    }
    {   // per-step recorders on the harmonic oscillator x'' = -x, x(0) = 1: x = cos t
        const System osc = System::harmonic_oscillator(1.0);
        auto every = EnsembleIVP::ode(osc, 0.0, 10.0, {1.0, 0.0}).every_step(400).method(ExplicitRungeKutta::dopri5().rtol(1e-8)).solve();
        deb::Solution s = every.at(0);
        CHECK((int)s.t.size() == s.steps.accepted + 1 && s.t.front() == 0.0 && s.t.back() == 10.0);
        auto dense = EnsembleIVP::ode(osc, 0.0, 10.0, {1.0, 0.0}).dense(3, 1200).method(ExplicitRungeKutta::dopri5().rtol(1e-8)).solve();
        s = dense.at(0);
        CHECK((int)s.t.size() == 3 * s.steps.accepted + 1);
        for (size_t r = 0; r < s.t.size(); r++) CHECK(std::fabs(s.y[r][0] - std::cos(s.t[r])) < 1e-4);
        auto cross = EnsembleIVP::ode(osc, 0.0, 10.0, {1.0, 0.0, 2.0, 0.0}).crossing(0, 0.0, CrossingDirection::Both, 8)
                         .method(ExplicitRungeKutta::dop853().rtol(1e-10).atol(1e-10)).solve();
        for (int i = 0; i < 2; i++) {
            s = cross.at(i);
            CHECK(s.t.size() == 3);
            for (size_t k = 0; k < s.t.size(); k++) CHECK(std::fabs(s.t[k] - (PI / 2.0 + (double)k * PI)) < 1e-3);
        }
        auto down = EnsembleIVP::ode(osc, 0.0, 10.0, {1.0, 0.0}).crossing(0, 0.0, CrossingDirection::Negative, 8)
                        .method(ExplicitRungeKutta::dopri5().rtol(1e-9).atol(1e-9)).solve();
        CHECK(down.at(0).t.size() == 2);  // pi/2 and 5 pi/2
        auto plane = EnsembleIVP::ode(osc, 0.0, 10.0, {1.0, 0.0}).hyperplane_crossing({0.0, 0.0}, {1.0, 1.0}, {0, 1}, CrossingDirection::Both, 8)
                         .method(ExplicitRungeKutta::dopri5().rtol(1e-9).atol(1e-9)).solve();
        s = plane.at(0);  // cos t - sin t = 0 at pi/4 + k pi
        CHECK(s.t.size() == 3);
        for (size_t k = 0; k < s.t.size(); k++) CHECK(std::fabs(s.t[k] - (PI / 4.0 + (double)k * PI)) < 1e-4);
        // plain solve() + a non-terminal event: every step plus the event rows, in time order
        auto evs = EnsembleIVP::ode(osc, 0.0, 10.0, {1.0, 0.0}).event(deb::Event::linear(0.0, 0.0, {1.0, 0.0}).direction(CrossingDirection::Positive), 400)
                       .method(ExplicitRungeKutta::dopri5().rtol(1e-8)).solve();
        s = evs.at(0);
        CHECK((int)s.t.size() == s.steps.accepted + 1 + 1 && s.status == deb::Status::Complete);  // one upward zero crossing in [0, 10]: 3 pi/2
        // a row capacity that is too small is an error of the caller, not silently truncated data
        auto small = EnsembleIVP::ode(osc, 0.0, 10.0, {1.0, 0.0}).every_step(5).method(ExplicitRungeKutta::dopri5()).solve();
        bool thrown = false;
        try { small.at(0); } catch (const std::length_error&) { thrown = true; }
        CHECK(thrown);
    }
    {   // errors are the crate's Error variants, per trajectory; option mistakes are caught on the host
        auto sol = EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 100.0, {1.0, 1.0, 1.0, 2.0, 1.0, 1.0})
                       .method(ExplicitRungeKutta::dopri5().rtol(1e-8).max_steps(100)).solve();
        for (int i = 0; i < 2; i++) {
            bool thrown = false;
            try { sol.at(i); } catch (const deb::Error& e) { thrown = e.kind == deb::Error::MaxSteps && e.t > 0.0 && e.t < 100.0 && e.y.size() == 3; }
            CHECK(thrown && !sol.ok(i));
        }
        auto bad = EnsembleIVP::ode(System::exponential_growth(1.0), 0.0, 0.0, {1.0}).method(ExplicitRungeKutta::dopri5()).solve();
        bool thrown = false;
        try { bad.at(0); } catch (const deb::Error& e) { thrown = e.kind == deb::Error::BadInput; }
        CHECK(thrown);
        auto stiff = EnsembleIVP::ode(System::exponential_growth(1.0000001), 0.0, 10.0, {1.0})
                         .method(ExplicitRungeKutta::dopri5().h_max(0.002).max_steps(100000)).solve();
        thrown = false;
        try { stiff.at(0); } catch (const deb::Error& e) { thrown = e.kind == deb::Error::Stiffness; }
        CHECK(thrown && stiff.accepted[0] == 1499);
        thrown = false;
        try { EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, {1.0, 1.0, 1.0}).method(ExplicitRungeKutta::dopri5().rtol({1e-6, 1e-6})).solve(); }
        catch (const std::invalid_argument&) { thrown = true; }
        CHECK(thrown);
        thrown = false;
        try { EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, {1.0, 1.0}); } catch (const std::invalid_argument&) { thrown = true; }
        CHECK(thrown);
        thrown = false;
        try { System::from_source(1, "dydt[0] = y[0];", {}); } catch (const deb::CallError& e) { thrown = e.code == DEB_ERR_UNSUPPORTED; }
        CHECK(thrown);  // (this build has the oracle behind the ABI)
        thrown = false;
        try { System::sensitivity_from_source(1, "dydt[0] = p[0] * y[0];", "J[0] = p[0];", "Jp[0] = y[0];", {1.0}); } catch (const deb::CallError& e) { thrown = e.code == DEB_ERR_UNSUPPORTED; }
        CHECK(thrown);
        // vector tolerances and fixed-step methods
        auto vt = EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, {1.0, 1.0, 1.0}).t_eval({0.5, 1.0})
                      .method(ExplicitRungeKutta::rkf45().rtol({1e-8, 1e-8, 1e-8}).atol({1e-9, 1e-9, 1e-9})).solve();
        auto st = EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, {1.0, 1.0, 1.0}).t_eval({0.5, 1.0})
                      .method(ExplicitRungeKutta::rkf45().rtol(1e-8).atol(1e-9)).solve();
        CHECK(vt.at(0).y[1][0] == st.at(0).y[1][0] && vt.at(0).steps.accepted == st.at(0).steps.accepted);
        auto rk4 = EnsembleIVP::ode(System::exponential_growth(1.0), 0.0, 1.0, {1.0}).t_eval({1.0}).method(ExplicitRungeKutta::rk4(0.01)).solve();
        CHECK(rk4.at(0).steps.accepted == 100 && std::fabs(rk4.at(0).y[0][0] - std::exp(1.0)) < 1e-9);
    }
    {   // IVP::sde: Ornstein-Uhlenbeck paths with Euler-Maruyama and Milstein; the sample moments follow the SDE's
        const int n = 4000;
        auto run = [&](const deb::ExplicitRungeKutta& m, uint64_t seed) {
            return deb::EnsembleSDE::sde(deb::SdeSystem::ornstein_uhlenbeck(0.5, 1.0, 0.3), 0.0, 2.0, std::vector<double>(n, 5.0), seed).t_eval({1.0, 2.0}).method(m).solve();
        };
        auto em = run(ExplicitRungeKutta::euler(0.01), 7), em2 = run(ExplicitRungeKutta::euler(0.01), 7), other = run(ExplicitRungeKutta::euler(0.01), 8);
        auto mil = run(ExplicitRungeKutta::milstein(0.01), 7);
        double mean = 0.0;
        for (int i = 0; i < n; i++) mean += em.at(i).y[1][0];
        mean /= n;
        CHECK(std::fabs(mean - (1.0 + 4.0 * std::exp(-1.0))) < 0.03);  // E[y(2)] = mu + (y0 - mu) e^{-theta t}
        CHECK(em.at(17).t.size() == 2 && em.at(17).t[1] == 2.0 && em.at(17).steps.accepted == 200);
        CHECK(em.at(17).y[1][0] == em2.at(17).y[1][0] && em.at(17).y[1][0] != other.at(17).y[1][0]);  // the stream is a function of (seed, path)
        CHECK(std::fabs(mil.at(17).y[1][0] - em.at(17).y[1][0]) < 1e-12);  // additive noise: the Milstein correction vanishes
    }
    {   // IVP::pde: heat equation by the method of lines, tests/pde/method_of_lines.rs:37-70 (decay of the first sine mode)
        const int n = 41;
        std::vector<double> u0(n);
        for (int i = 0; i < n; i++) u0[i] = std::sin(PI * i / (n - 1.0));
        const double alpha = 0.1, dx = 1.0 / (n - 1.0);
        auto heat = deb::solve_heat_mol(u0, 0.0, 1.0, alpha, ExplicitRungeKutta::rk4(0.2 * dx * dx / alpha), 0.0, 0.1);
        CHECK(heat.status == DEB_STATUS_COMPLETE && std::fabs(heat.t - 0.1) < 1e-12 && heat.steps > 0);
        CHECK(std::fabs(heat.u[n / 2] - std::exp(-alpha * PI * PI * 0.1)) < 5e-4 && heat.u[0] == u0[0] && heat.u[n - 1] == u0[n - 1]);  // Dirichlet nodes do not move
    }
    std::printf(failures ? "%d check(s) FAILED\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
