// Development probe: are the one-cell (V1) and two-cell packed (V2) instantiations of the collision templates bit-identical ON THE DEVICE?
// (tests/host_math_check.cu checks it on the host.)  Counts mismatching populations per operator over random inputs.
#include <cstdio>
#include <cuda_runtime.h>
#include "../cuda_lbm_b200/csrc/collide.cuh"
using namespace lbm;
__device__ float rnd(unsigned& s) { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); }
__global__ void probe(int* bad, int forced) {
    unsigned s = 12345u + 977u * (blockIdx.x * blockDim.x + threadIdx.x);
    Relax r; r.omega = 1.0f / (3 * 0.0064f + 0.5f); r.quirks = 127;
    const float Sm[9] = {0, r.omega, r.omega, 0, r.omega, 0, r.omega, r.omega, r.omega}, Sc[9] = {0, 0, 0, 1, r.omega, r.omega, 1, 1, 1};
    for (int op = 0; op < 4; op++) {
        for (int i = 0; i < 9; i++) r.S[i] = op >= 2 ? Sc[i] : Sm[i];
        float g[2][9];
        for (int l = 0; l < 2; l++) {
            const float rho = 0.9f + 0.2f * rnd(s), ux = 0.1f * (rnd(s) - 0.5f), uy = 0.1f * (rnd(s) - 0.5f);
            for (int q = 0; q < 9; q++) { const float cu = cx(q) * ux + cy(q) * uy; g[l][q] = wq(q) * rho * (1 + 3 * cu + 4.5f * cu * cu - 1.5f * (ux * ux + uy * uy)) * (1 + 0.05f * (rnd(s) - 0.5f)); }
        }
        const float F[4] = {forced ? 1e-3f * (rnd(s) - 0.5f) : 0.f, forced ? 1e-3f * (rnd(s) - 0.5f) : 0.f, forced ? 1e-3f * (rnd(s) - 0.5f) : 0.f, forced ? 1e-3f * (rnd(s) - 0.5f) : 0.f};
        V1 a[2][9]; V2 b[9];
        for (int q = 0; q < 9; q++) { a[0][q].a = g[0][q]; a[1][q].a = g[1][q]; b[q].a = make_float2(g[0][q], g[1][q]); }
        const Mom<V1> m0 = moments_v(a[0]), m1 = moments_v(a[1]); const Mom<V2> m2 = moments_v(b);
        if (m2.rho.a.x != m0.rho.a || m2.ux.a.y != m1.ux.a || m2.uy.a.x != m0.uy.a) atomicAdd(&bad[4], 1);
        V1 u0x = m0.ux, u0y = m0.uy, u1x = m1.ux, u1y = m1.uy; V2 ux2 = m2.ux, uy2 = m2.uy;
        if (forced) {
            u0x = fma(V1{F[0]}, m0.inv_rho * 0.5f, u0x); u0y = fma(V1{F[1]}, m0.inv_rho * 0.5f, u0y);
            u1x = fma(V1{F[2]}, m1.inv_rho * 0.5f, u1x); u1y = fma(V1{F[3]}, m1.inv_rho * 0.5f, u1y);
            ux2 = fma(V2{make_float2(F[0], F[2])}, m2.inv_rho * 0.5f, ux2); uy2 = fma(V2{make_float2(F[1], F[3])}, m2.inv_rho * 0.5f, uy2);
        }
        const V2 Fx2{make_float2(F[0], F[2])}, Fy2{make_float2(F[1], F[3])};
        const float hi0 = 1.8f + 0.19f * rnd(s), hi1 = 1.8f + 0.19f * rnd(s);
        if (op == 0) { collide_bgk_v(r, a[0], m0.rho, u0x, u0y, forced != 0, V1{F[0]}, V1{F[1]}); collide_bgk_v(r, a[1], m1.rho, u1x, u1y, forced != 0, V1{F[2]}, V1{F[3]}); collide_bgk_v(r, b, m2.rho, ux2, uy2, forced != 0, Fx2, Fy2); }
        else if (op == 1) { collide_mrt_v(r, a[0], m0.rho, u0x, u0y, forced != 0, V1{F[0]}, V1{F[1]}); collide_mrt_v(r, a[1], m1.rho, u1x, u1y, forced != 0, V1{F[2]}, V1{F[3]}); collide_mrt_v(r, b, m2.rho, ux2, uy2, forced != 0, Fx2, Fy2); }
        else if (op == 2) { collide_cm_v<false>(r, a[0], u0x, u0y, forced != 0, V1{F[0]}, V1{F[1]}, V1{1.f}); collide_cm_v<false>(r, a[1], u1x, u1y, forced != 0, V1{F[2]}, V1{F[3]}, V1{1.f}); collide_cm_v<false>(r, b, ux2, uy2, forced != 0, Fx2, Fy2, splat<V2>(1.f)); }
        else { collide_cm_v<true>(r, a[0], u0x, u0y, forced != 0, V1{F[0]}, V1{F[1]}, V1{hi0}); collide_cm_v<true>(r, a[1], u1x, u1y, forced != 0, V1{F[2]}, V1{F[3]}, V1{hi1}); collide_cm_v<true>(r, b, ux2, uy2, forced != 0, Fx2, Fy2, V2{make_float2(hi0, hi1)}); }
        for (int q = 0; q < 9; q++) if (b[q].a.x != a[0][q].a || b[q].a.y != a[1][q].a) atomicAdd(&bad[op], 1);
    }
}
int main() {
    int* d; cudaMalloc(&d, 5 * sizeof(int));
    for (int forced = 0; forced < 2; forced++) {
        cudaMemset(d, 0, 5 * sizeof(int));
        probe<<<64, 128>>>(d, forced);
        int h[5]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("V1V2_PROBE forced=%d mismatching populations of %d: BGK %d MRT %d CM %d CM_OPT %d, moments %d  (%s)\n", forced, 64 * 128 * 9, h[0], h[1], h[2], h[3], h[4], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
