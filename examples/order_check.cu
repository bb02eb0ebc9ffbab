// examples/order_check.cu — test driver (tests/test_shim_gpu.py) for two properties of the header shim that the reference has by
// construction and a fused engine has to work for:
//   1. d_rho / d_u are "always current" in the reference (src/core/lbm.cuh:148-154): save_checkpoint() followed by save_vtk(), and
//      update_macroscopics() after the next step's increase_ts(), are legal call orders and must not abort.
//   2. reset_forces<Scenario>() re-evaluates Init::apply_forces every step (src/core/macroscopics/macroscopics.cuh:13-48): a scenario
//      whose force depends on time (opt-in: `static constexpr bool time_dependent_forces = true;`) must see it change step by step.
// Build: -DNX=64 -DNY=64 -DSCALE=1 (examples/Makefile, _bin/t_order_check).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "core/lbm.cuh"
#include "functors/includes.cuh"
#include "scenarios/b200_taylor_green.cuh"

// fluid at rest pushed by a uniform force that changes every step; the value is exact in fp32 on host and device alike
__host__ __device__ inline void pulsed_force(int t, float& fx, float& fy) {
    fx = 1.0e-5f * (float)((t % 4) - 1);
    fy = -2.0e-5f * (float)((t % 3) - 1);
}
struct PulsedInit {
    int t;
    explicit PulsedInit(int t) : t(t) {}
    __host__ __device__ void apply_forces(float* rho, float* u, float* force, int node) {
        float fx, fy;
        pulsed_force(t, fx, fy);
        force[get_vec_index(node, 0)] = fx;
        force[get_vec_index(node, 1)] = fy;
    }
    __host__ __device__ void operator()(float* rho, float* u, float* force, int node) {
        rho[node] = 1.0f;
        u[get_vec_index(node, 0)] = 0.0f;
        u[get_vec_index(node, 1)] = 0.0f;
        apply_forces(rho, u, force, node);
    }
};
struct PulsedScenario : public ScenarioTrait<PulsedInit, B200AllFluid, void, BGK<2>> {
    static constexpr bool periodic_x = true, periodic_y = true;
    static constexpr bool time_dependent_forces = true;
    static const char* name() { return "Pulsed"; }
    static InitType init() { return InitType((int)t); }
    static BoundaryType boundary() { return BoundaryType(); }
};

template <typename S>
static void protocol_step(LBM<2>& lbm, bool with_increase = true) {
    if (with_increase) lbm.increase_ts<S>();
    lbm.stream();
    lbm.swap_buffers();
    lbm.apply_boundaries<S>();
    lbm.uncorrected_macroscopics();
    lbm.reset_forces<S>();
    lbm.ibm_step();
    lbm.correct_macroscopics();
    lbm.compute_equilibrium();
    lbm.collide<typename S::CollisionOp>();
}

static float max_abs_diff(const std::vector<float>& a, const std::vector<float>& b) {
    float m = 0.0f;
    for (size_t i = 0; i < a.size(); i++) m = std::fmax(m, std::fabs(a[i] - b[i]));
    return m;
}

int main(int argc, char** argv) {
    const char* ckpt = argc > 1 ? argv[1] : "order_check.ckpt";
    checkCudaErrors(cudaSetDevice(0));
    int failures = 0;
    {   // ---- 1. call orders
        using S = B200TaylorGreenScenario;
        LBM<2> a, b;
        a.allocate<S>(); a.init<S>();
        for (int i = 0; i < 5; i++) protocol_step<S>(a);
        a.save_checkpoint(ckpt);            // closes step 5 ...
        a.save_vtk(5);                      // ... whose rho / u must still be there
        std::vector<float> rho5 = a.h_rho, u5 = a.h_u;
        protocol_step<S>(a);                // step 6 described
        a.increase_ts<S>();                 // step 6 enqueued WITHOUT macroscopics, step 7 begun
        a.update_macroscopics();            // the fields of step 6, rebuilt from its populations
        b.allocate<S>(); b.init<S>();
        for (int i = 0; i < 5; i++) protocol_step<S>(b);
        b.update_macroscopics();
        const float d5r = max_abs_diff(rho5, b.h_rho), d5u = max_abs_diff(u5, b.h_u);
        protocol_step<S>(b);
        b.update_macroscopics();
        const float d6r = max_abs_diff(a.h_rho, b.h_rho), d6u = max_abs_diff(a.h_u, b.h_u);
        const bool ok = d5r == 0.0f && d5u == 0.0f && d6r <= 1e-6f && d6u <= 1e-7f;
        printf("ORDER_CHECK checkpoint-then-vtk max|drho|=%.3e max|du|=%.3e ; recovered step-6 fields max|drho|=%.3e max|du|=%.3e %s\n", d5r, d5u, d6r, d6u, ok ? "OK" : "FAIL");
        failures += !ok;
    }
    {   // ---- 2. forces re-evaluated every step
        using S = PulsedScenario;
        const int steps = 24;
        LBM<2> a;
        a.allocate<S>(); a.init<S>();
        for (int i = 0; i < steps; i++) protocol_step<S>(a);
        a.update_macroscopics();
        LBM<2> c;                           // the same through LBM::run
        S::t = 0.0f;
        c.allocate<S>(); c.init<S>();
        c.run<S>(steps);
        c.update_macroscopics();
        // the same physics straight through the C ABI: a uniform force set anew before every step
        lbm_config cfg;
        lbm_default_config(&cfg);
        cfg.nx = NX; cfg.ny = NY; cfg.periodic_x = 1; cfg.periodic_y = 1; cfg.collision = LBM_BGK; cfg.viscosity = S::viscosity;
        for (int i = 0; i < 9; i++) cfg.S[i] = S::S[i];
        cfg.u_max = S::u_max;
        lbm_handle* h = nullptr;
        LBM_B200_CALL(lbm_create(&cfg, &h));
        std::vector<float> rho((size_t)NX * NY, 1.0f), u((size_t)NX * NY * 2, 0.0f);
        LBM_B200_CALL(lbm_init_fields(h, rho.data(), u.data()));
        for (int t = 1; t <= steps; t++) {
            float fx, fy;
            pulsed_force(t, fx, fy);
            LBM_B200_CALL(lbm_sync(h));
            LBM_B200_CALL(lbm_set_body_force(h, fx, fy));
            LBM_B200_CALL(t == steps ? lbm_step_with_macroscopics(h, 1) : lbm_step(h, 1));
        }
        LBM_B200_CALL(lbm_get_macroscopics(h, rho.data(), u.data()));
        LBM_B200_CALL(lbm_destroy(h));
        double mom = 0.0;
        for (size_t i = 0; i < u.size(); i++) mom += std::fabs(u[i]);
        const float dr = max_abs_diff(a.h_rho, rho), du = max_abs_diff(a.h_u, u), dc = max_abs_diff(c.h_u, u);
        const bool ok = dr == 0.0f && du == 0.0f && dc == 0.0f && mom > 0.0;
        printf("ORDER_CHECK time-dependent force: protocol vs C ABI max|drho|=%.3e max|du|=%.3e, LBM::run vs C ABI max|du|=%.3e, sum|u|=%.3e %s\n", dr, du, dc, mom, ok ? "OK" : "FAIL");
        failures += !ok;
    }
    printf("ORDER_CHECK %s\n", failures ? "FAILED" : "PASSED");
    return failures ? 1 : 0;
}
