// examples/main.cu — main.cu-style driver for the header shim (include/cuda-lbm/): the time loop of the reference's
// src/main.cu:72-153 written against the same LBM<2> / ScenarioTrait interface, with run-time step counts and a
// machine-readable result line.  Build one binary per (scenario, NX, NY) — scenario functors bake the grid in as macros:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -Iinclude/cuda-lbm -I<dir with scenarios/...> \
//        -DUSE_TAYLOR_GREEN -DPERIODIC_X -DPERIODIC_Y examples/main.cu -Lcuda_lbm_b200 -llbm_b200 -o tg
//
// <dir with scenarios/...> is the reference's src/ (its scenario files compile unchanged) or examples/ (this repo's own).
// Any other scenario:  -DSCENARIO_HEADER='"my/scenario.cuh"' -DSCENARIO_TYPE=MyScenario -DNX=.. -DNY=..
//
//   ./tg [--steps N] [--save-int K] [--warmup W] [--vtk] [--dump] [--fast] [--load-ckpt FILE] [--save-ckpt FILE]
//     --warmup W    W untimed steps first (module load, CUDA-graph capture); --steps counts the timed steps that follow
//     --repeat R    the whole segment (init, warm-up, timed steps) R times; times add up (for scenarios that stay physical for a
//                   limited number of steps only, e.g. CM<2,OptimalAdapter> on the impulsively started cavity)
//     --save-int K  every K steps: update_macroscopics + compute_error (and --vtk: save_vtk, --dump: save_macroscopics)
//     --fast        no per-step host calls between save points (LBM::run), the throughput mode
//     --load-ckpt F continue from a checkpoint written by --save-ckpt (same binary); --steps counts the steps still to run
//     --save-ckpt F write the population state after the last step
#include <stdio.h>
#include <chrono>
#include <cstring>
#include <iostream>
#include "core/lbm.cuh"
#include "functors/includes.cuh"
#include "util/timer.cuh"
#include "IBM/IBMBody.cuh"

#if defined(SCENARIO_HEADER)
#include SCENARIO_HEADER
using Scenario = SCENARIO_TYPE;
#elif defined(USE_TAYLOR_GREEN)
#include "scenarios/taylorGreen/taylorGreenScenario.cuh"
using Scenario = TaylorGreenScenario;
#elif defined(USE_POISEUILLE)
#include "scenarios/poiseuille/poiseuilleScenario.cuh"
using Scenario = PoiseuilleScenario;
#elif defined(USE_LID_DRIVEN)
#include "scenarios/lidDrivenCavity/lidDrivenCavityScenario.cuh"
using Scenario = LidDrivenScenario;
#elif defined(USE_FLOW_PAST_CYLINDER)
#include "scenarios/flowPastCylinder/flowPastCylinderScenario.cuh"
using Scenario = FlowPastCylinderScenario;
#else
#error "select a scenario: -DUSE_TAYLOR_GREEN / -DUSE_POISEUILLE / -DUSE_LID_DRIVEN / -DUSE_FLOW_PAST_CYLINDER or -DSCENARIO_HEADER/-DSCENARIO_TYPE"
#endif

int main(int argc, char** argv) {
    int total_timesteps = 1000, save_int = 100, warmup = 0, repeat = 1;
    bool vtk = false, dump = false, fast = false;
    const char *load_ckpt = nullptr, *save_ckpt = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--steps") && i + 1 < argc) total_timesteps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--save-int") && i + 1 < argc) save_int = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = std::max(1, atoi(argv[++i]));
        else if (!strcmp(argv[i], "--vtk")) vtk = true;
        else if (!strcmp(argv[i], "--dump")) dump = true;
        else if (!strcmp(argv[i], "--fast")) fast = true;
        else if (!strcmp(argv[i], "--load-ckpt") && i + 1 < argc) load_ckpt = argv[++i];
        else if (!strcmp(argv[i], "--save-ckpt") && i + 1 < argc) save_ckpt = argv[++i];
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (save_int <= 0) save_int = total_timesteps;
    checkCudaErrors(cudaSetDevice(0));
    cudaDeviceProp prop;
    checkCudaErrors(cudaGetDeviceProperties(&prop, 0));
    constexpr float Re = compute_reynolds(Scenario::u_max, NY, Scenario::viscosity);
    std::cout << "Running " << Scenario::name() << " scenario on " << prop.name << ", " << NX << " x " << NY << std::endl;
    std::cout << "Viscosity: " << Scenario::viscosity << ", Tau: " << Scenario::tau << std::endl;
    std::cout << "Reynolds number: " << Re << std::endl;

    LBM<dimensions> lbm;
    lbm.allocate<Scenario>();
    lbm.init<Scenario>();
    if (load_ckpt) {
        lbm.load_checkpoint<Scenario>(load_ckpt);
        printf("restarted from %s at step %d\n", load_ckpt, lbm.timestep);
    }
    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    float last_error = -1.0f, gpu_ms = 0.0f;
    for (int rep = 0; rep < repeat; rep++) {
    if (rep > 0) lbm.init<Scenario>();
    if (warmup > 0) {
        lbm.run<Scenario>(warmup);
        lbm.synchronize();
    }
    const int first_step = lbm.timestep;
    int t = 0;
    while (t < total_timesteps) {
        const int chunk = std::min(save_int, total_timesteps - t);
        const auto wall0 = std::chrono::steady_clock::now();
        cudaEventRecord(start);
        if (fast) {
            lbm.run<Scenario>(chunk);
        } else {
            for (int k = 0; k < chunk; k++) {
                lbm.increase_ts<Scenario>();
                lbm.stream();
                lbm.swap_buffers();
                lbm.apply_boundaries<Scenario>();
                lbm.uncorrected_macroscopics();
                lbm.reset_forces<Scenario>();
                lbm.ibm_step();
                lbm.correct_macroscopics();
                lbm.compute_equilibrium();
                lbm.collide<Scenario::CollisionOp>();
            }
        }
        t += chunk;
        lbm.finish_step();                  // the last step of the chunk also stores rho and u (+12 B/node on that step only)
        if (lbm.num_slabs() > 1) lbm.synchronize();        // several GPUs (LBM_B200_GPUS): events on this device's stream do not see the others
        cudaEventRecord(stop);
        cudaEventSynchronize(stop);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, start, stop);
        if (lbm.num_slabs() > 1) ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        gpu_ms += ms;
        lbm.update_macroscopics();          // device -> host copy of rho, u: outside the timed region, as in the reference's save path
        if (vtk) lbm.save_vtk(t);
        if (dump) lbm.save_macroscopics(t);
        if constexpr (Scenario::has_analytical_solution) {
            last_error = lbm.compute_error<Scenario>();
            printf("%s[%d]: error, %.4f%%\n", Scenario::name(), first_step + t, last_error);
            if constexpr (lbm_b200_shim::is_centerline_validation<typename Scenario::ValidationType>::value)
                printf("%s[%d]: centre-line NRMSE against Ghia et al. with the samples gathered on the device, %.4f%%\n", Scenario::name(), first_step + t,
                       lbm.centerline_error_device<Scenario>());
#ifndef LBM_B200_NO_DEVICE_ERROR      // needs a __host__ __device__ Validation::operator()(x, y, ux&, uy&), as the reference's are
            if constexpr (lbm_b200_shim::is_field_validation<typename Scenario::ValidationType>::value)
                printf("%s[%d]: L2 error taken on the device, %.4f%%\n", Scenario::name(), first_step + t, lbm.l2_error_device<Scenario>());
#endif
        }
    }
    }
    if (save_ckpt) {
        lbm.save_checkpoint(save_ckpt);
        printf("checkpoint of step %d written to %s\n", lbm.timestep, save_ckpt);
    }
    const double mass = lbm.total_mass();
    double sum_rho = 0.0, sum_u2 = 0.0;
    for (size_t i = 0; i < lbm.h_rho.size(); i++) sum_rho += lbm.h_rho[i];
    for (size_t i = 0; i < lbm.h_u.size(); i++) sum_u2 += (double)lbm.h_u[i] * lbm.h_u[i];
    printf("SHIM_RESULT scenario=%s gpus=%d nx=%d ny=%d steps=%d ms_per_step=%.6f mlups=%.3f error_pct=%.6f mass_per_node=%.9f mean_rho=%.9f sum_u2=%.9e\n",
           Scenario::name(), lbm.num_slabs(), NX, NY, total_timesteps * repeat, gpu_ms / (total_timesteps * repeat),
           lbm_b200_mlups((long long)NX * NY, (long long)total_timesteps * repeat, gpu_ms * 1e-3), last_error, mass / ((double)NX * NY),
           sum_rho / ((double)NX * NY), sum_u2);
    return 0;
}
