// problems.hpp -- the LCQPs of the reference's examples as plain data (used by the examples and host_tests).
//   warm_up:          /root/reference/examples/warm_up.cpp:32-42
//   OptimizeOnCircle: /root/reference/examples/OptimizeOnCircle.cpp:31-99 (N facets, nV = 2 + 2N, nC = N + 1, nComp = N)
#ifndef LCQPOW_B200_EXAMPLE_PROBLEMS_HPP
#define LCQPOW_B200_EXAMPLE_PROBLEMS_HPP

#include <cmath>
#include <vector>

namespace examples {

struct Problem {
    int nV = 0, nC = 0, nComp = 0;
    std::vector<double> Q, g, L, R, A, lbA, ubA, x0, y0;
};

// min (x-1)^2 + (y-1)^2  s.t. 0 <= x  _|_  y >= 0, started from (1,1)
inline Problem warmUp()
{
    Problem p;
    p.nV = 2; p.nC = 0; p.nComp = 1;
    p.Q = {2.0, 0.0, 0.0, 2.0};
    p.g = {-2.0, -2.0};
    p.L = {1.0, 0.0};
    p.R = {0.0, 1.0};
    p.x0 = {1.0, 1.0};
    p.y0 = {0.0, 0.0, 0.0, 0.0};
    return p;
}

// Closest point to (xr, yr) in the metric [[17,-15],[-15,17]] on the piecewise-linear unit circle with N
// facets: facet i contributes the slack lambda_i of its tangent and the weight theta_i, lambda_i theta_i = 0.
inline Problem circle(int N, double xr, double yr)
{
    const double pi = 3.14159265358979323846;
    Problem p;
    p.nV = 2 + 2 * N; p.nC = N + 1; p.nComp = N;
    const int n = p.nV;
    p.Q.assign((size_t)n * n, 0.0);
    p.g.assign(n, 0.0);
    p.L.assign((size_t)N * n, 0.0);
    p.R.assign((size_t)N * n, 0.0);
    p.A.assign((size_t)(N + 1) * n, 0.0);
    p.lbA.assign(N + 1, 1.0);
    p.ubA.assign(N + 1, 1.0);
    p.x0.assign(n, 1.0);
    p.Q[0] = 17.0; p.Q[1] = -15.0; p.Q[n] = -15.0; p.Q[n + 1] = 17.0;
    for (int i = 2; i < n; ++i) p.Q[(size_t)i * n + i] = 5e-12;
    p.g[0] = -(17.0 * xr - 15.0 * yr);
    p.g[1] = -(-15.0 * xr + 17.0 * yr);
    p.x0[0] = xr; p.x0[1] = yr;
    for (int i = 0; i < N; ++i) {
        const double ang = (2.0 * pi * i) / N;
        p.A[(size_t)i * n + 0] = std::cos(ang);
        p.A[(size_t)i * n + 1] = std::sin(ang);
        p.A[(size_t)i * n + 2 + 2 * i] = 1.0;     // + lambda_i = 1
        p.A[(size_t)N * n + 3 + 2 * i] = 1.0;     // sum theta = 1
        p.L[(size_t)i * n + 2 + 2 * i] = 1.0;
        p.R[(size_t)i * n + 3 + 2 * i] = 1.0;
    }
    return p;
}

}  // namespace examples

#endif
