// The C++ host mirror of the crate's builder API (include/deb_ensemble.hpp) in use: the calls read like the reference's own tests and
// examples, for N problems at a time.
//
//   g++ -std=c++17 -O2 -I../../include builder_api.cpp -L../../differential-equations_b200 -ldeb200 -Wl,-rpath,<dir of libdeb200.so> -o builder_api
//
// Exit code: 0 = every check passed, 3 = no CUDA device (the library has no CPU fallback), 1 = a check failed.
#include <cmath>
#include <cstdio>

#include "deb_ensemble.hpp"

static int failures = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); failures++; } \
    } while (0)

int main() {
    using deb::EnsembleIVP;
    using deb::ExplicitRungeKutta;
    using deb::System;
    try {
        // --- benches/solvers/adaptive_step.rs:191-205 (Lorenz, dopri5().rtol(1e-8)) over an ensemble; trajectory 0 = (1, 1, 1):
        //     the known answers of SURVEY.md appendix A (L100), bit for bit
        {
            const int n = 2048;
            std::vector<double> y0(3 * n);
            for (int i = 0; i < n; i++)
                for (int c = 0; c < 3; c++) y0[3 * i + c] = 1.0 + 1e-3 * i * (c + 1);
            auto sol = EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 100.0, y0)
                           .t_eval({1.0, 2.5, 100.0})
                           .method(ExplicitRungeKutta::dopri5().rtol(1e-8))
                           .with_stats()
                           .solve();
            const deb::Solution s = sol.at(0);
            CHECK(s.status == deb::Status::Complete && s.steps.accepted == 6008 && s.steps.rejected == 411 && s.evals.function == 44525);
            CHECK(s.t.size() == 3 && s.t[0] == 1.0 && s.t[1] == 2.5 && s.t[2] == 100.0);
            CHECK(s.y[0][0] == -9.378567031548476 && s.y[0][1] == -8.357039113339846 && s.y[0][2] == 29.36231537793399);
            CHECK(s.y[2][0] == -0x1.8c65ec78fb24fp+1 && s.y[2][1] == -0x1.4d4949b09d281p+1 && s.y[2][2] == 0x1.5bf4de76ed043p+4);
            CHECK(sol.stats_counts[0] == n && sol.stats_counts[2] == n);
            double sum_x = 0.0;
            for (int i = 0; i < n; i++) sum_x += sol.at(i).y[0][0];
            CHECK(std::fabs(sum_x - sol.stats_sums[0]) <= 1e-9 * std::fabs(sum_x) + 1e-9);
            std::printf("Lorenz x %d: kernel %.2f ms, %d launches, trajectory 0 reproduces the known answers\n", n, sol.kernel_ms, sol.gpu_launches);
        }
        // --- examples/ode/03_logistic_growth: even(2.0) + a terminal event at 90 % of the carrying capacity, swept over the capacity
        {
            const int n = 300;
            std::vector<double> y0(n, 1.0), km(2 * n);
            for (int i = 0; i < n; i++) { km[2 * i] = 1.0; km[2 * i + 1] = 10.0 + 0.05 * i; }
            // g = y - 9 (a linear event): every trajectory reaches 9 before t = 10
            auto sol = EnsembleIVP::ode(System::logistic_equation(1.0, 10.0).sweep(km), 0.0, 10.0, y0)
                           .even(2.0)
                           .event(deb::Event::linear(-9.0, 0.0, {1.0}).terminal())
                           .method(ExplicitRungeKutta::dop853().rtol(1e-12).atol(1e-12))
                           .solve();
            int interrupted = 0;
            double worst_y = 0.0, worst_t = 0.0;
            for (int i = 0; i < n; i++) {
                const deb::Solution s = sol.at(i);
                interrupted += s.status == deb::Status::Interrupted;
                worst_y = std::fmax(worst_y, std::fabs(s.y.back()[0] - 9.0));
                // closed form: y(t) = m / (1 + (m - 1) e^{-t})  =>  t* = ln(9 (m - 1) / (m - 9))
                const double m = km[2 * i + 1];
                worst_t = std::fmax(worst_t, std::fabs(s.t.back() - std::log(9.0 * (m - 1.0) / (m - 9.0))));
            }
            // (the event is located on the reference's dense output as written -- DOP853's Horner factors, SURVEY.md 8a row 7 -- hence 1e-2)
            CHECK(interrupted == n && worst_y < 0.05 && worst_t < 0.01);
            std::printf("logistic sweep: %d of %d stopped by the event; event state within %.1e of 9, event time within %.1e of the closed form\n",
                        interrupted, n, worst_y, worst_t);
        }
        // --- examples/ode/08_damped_oscillator: a user-defined right-hand side, zero crossings of x located on the dense output
        {
            System osc = System::from_source(2, "dydt[0] = y[1]; dydt[1] = -p[0] * y[1] - p[1] * y[0];", {0.5, 1.0});
            auto sol = EnsembleIVP::ode(osc, 0.0, 20.0, {1.0, 0.0, 2.0, 0.0})
                           .crossing(0, 0.0, deb::CrossingDirection::Both, 32)
                           .method(ExplicitRungeKutta::dopri5().rtol(1e-8).atol(1e-8))
                           .solve();
            const double w = std::sqrt(15.0) / 4.0;  // x(t) = e^{-t/4} (cos wt + sin wt / (4w)): zeros at wt = atan(-4w) + k pi
            for (int i = 0; i < 2; i++) {
                const deb::Solution s = sol.at(i);
                CHECK(s.t.size() == 6);
                for (size_t k = 0; k < s.t.size(); k++) {
                    CHECK(std::fabs(s.t[k] - (std::atan(-4.0 * w) + (double)(k + 1) * M_PI) / w) < 1e-5);
                    CHECK(std::fabs(s.y[k][0]) < 1e-9);
                }
            }
            std::printf("damped oscillator (user-defined right-hand side): zero crossings match the closed form\n");
        }
        // --- errors are the crate's Error variants, per trajectory (tests/ode/errors.rs)
        {
            auto sol = EnsembleIVP::ode(System::lorenz(10.0, 28.0, 8.0 / 3.0), 0.0, 100.0, {1.0, 1.0, 1.0})
                           .method(ExplicitRungeKutta::dopri5().rtol(1e-8).max_steps(100))
                           .solve();
            bool thrown = false;
            try { sol.at(0); } catch (const deb::Error& e) { thrown = e.kind == deb::Error::MaxSteps && e.t > 0.0 && e.t < 100.0 && e.y.size() == 3; }
            CHECK(thrown);
            auto bad = EnsembleIVP::ode(System::exponential_growth(1.0), 0.0, 0.0, {1.0}).method(ExplicitRungeKutta::dopri5()).solve();
            thrown = false;
            try { bad.at(0); } catch (const deb::Error& e) { thrown = e.kind == deb::Error::BadInput; }
            CHECK(thrown);
            std::printf("MaxSteps and BadInput come back as the reference's Error variants\n");
        }
        // --- examples/ode/16_sensitivity pattern: forward sensitivities generated from diff / jacobian / jacobian_p (logistic growth, p = {k, m})
        {
            System sens = System::sensitivity_from_source(1, "dydt[0] = p[0] * y[0] * (1.0 - y[0] / p[1]);", "J[0] = p[0] * (1.0 - 2.0 * y[0] / p[1]);",
                                                          "Jp[0] = y[0] * (1.0 - y[0] / p[1]); Jp[1] = p[0] * y[0] * y[0] / (p[1] * p[1]);", {1.0, 10.0});
            auto sol = EnsembleIVP::ode(sens, 0.0, 2.0, {1.0, 0.0, 0.0}).t_eval({2.0}).method(ExplicitRungeKutta::dop853().rtol(1e-11).atol(1e-11)).solve();
            const std::vector<double> z = sol.at(0).y[0];  // [y, dy/dk, dy/dm]
            const double e = std::exp(-2.0), d = 1.0 + 9.0 * e;
            CHECK(z.size() == 3 && std::fabs(z[0] - 10.0 / d) < 1e-8 && std::fabs(z[1] - 10.0 * 9.0 * 2.0 * e / (d * d)) < 1e-7 && std::fabs(z[2] - (1.0 - e) / (d * d)) < 1e-8);
            std::printf("logistic growth with generated forward sensitivities: dy/dk = %.8f, dy/dm = %.8f at t = 2\n", z[1], z[2]);
        }
        // --- examples/sde/03_ornstein_uhlenbeck as an ensemble, and a user-defined SDE that must reproduce it bit for bit
        {
            const int n = 100000;
            auto run = [&](const deb::SdeSystem& sys) {
                return deb::EnsembleSDE::sde(sys, 0.0, 10.0, std::vector<double>(n, 5.0), 42).t_eval({10.0}).method(ExplicitRungeKutta::euler(0.01)).solve();
            };
            auto ou = run(deb::SdeSystem::ornstein_uhlenbeck(0.5, 1.0, 0.3));
            auto user = run(deb::SdeSystem::from_source(1, "dydt[0] = p[0] * (p[1] - y[0]);", "g[0] = p[2];", {0.5, 1.0, 0.3}));
            double mean = 0.0, var = 0.0;
            bool same = true;
            for (int i = 0; i < n; i++) { mean += ou.y_final[i]; same = same && ou.y_final[i] == user.y_final[i]; }
            mean /= n;
            for (int i = 0; i < n; i++) var += (ou.y_final[i] - mean) * (ou.y_final[i] - mean);
            var /= n;
            CHECK(same);
            CHECK(std::fabs(mean - (1.0 + 4.0 * std::exp(-5.0))) < 5e-3 && std::fabs(var - 0.09 * (1.0 - std::exp(-10.0))) < 3e-3);  // sigma^2 / (2 theta) = 0.09
            std::printf("Ornstein-Uhlenbeck x %d: mean %.4f, variance %.4f at t = 10; the user-defined SDE gives the same bits\n", n, mean, var);
        }
        // --- tests/pde/method_of_lines.rs:37-70: heat equation, decay of the first sine mode
        {
            const int n = 4097;
            const double PI = 3.14159265358979323846, alpha = 0.1, dx = 1.0 / (n - 1.0);
            std::vector<double> u0(n);
            for (int i = 0; i < n; i++) u0[i] = std::sin(PI * i / (n - 1.0));
            auto heat = deb::solve_heat_mol(u0, 0.0, 1.0, alpha, ExplicitRungeKutta::rk4(0.2 * dx * dx / alpha).max_steps(1000000), 0.0, 0.01);
            CHECK(heat.status == DEB_STATUS_COMPLETE && std::fabs(heat.u[n / 2] - std::exp(-alpha * PI * PI * 0.01)) < 1e-6);
            std::printf("heat equation, %d nodes, %lld RK4 steps: u(1/2, 0.01) = %.9f (exact %.9f)\n", n, heat.steps, heat.u[n / 2], std::exp(-alpha * PI * PI * 0.01));
        }
    } catch (const deb::CallError& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return e.code == DEB_ERR_NO_DEVICE ? 3 : 1;
    }
    std::printf(failures ? "%d check(s) FAILED\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
