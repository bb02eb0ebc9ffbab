// Development probe: one warp, one 128 x 1 x 1 TMA box load + shifted store through a 3-D tensor map, descriptor in param space
// (mode 0) or in global memory (mode 1).  Isolates tensor-map / PTX problems from the solver.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tmg, int shift) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* buf = (float*)sm;
    unsigned long long* bar = (unsigned long long*)(sm + 512);
    const void* d = MODE == 0 ? (const void*)&tm : (const void*)tmg;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(s32(bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(s32(buf)), "l"((unsigned long long)d), "r"(s32(bar)), "r"(0 - shift), "r"(1), "r"(0) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(bar)) : "memory");
    float4 v = *(float4*)(buf + 4 * lane);
    v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
    *(float4*)(buf + 4 * lane) = v;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((unsigned long long)d), "r"(s32(buf)), "r"(shift), "r"(2), "r"(1) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main() {
    const int nx = 256, ny = 4, np = 2;
    float* d; CK(cudaMalloc(&d, nx * ny * np * 4));
    std::vector<float> h(nx * ny * np);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[3] = {nx, ny, np}, str[2] = {nx * 4, nx * ny * 4};
    cuuint32_t box[3] = {128, 1, 1}, es[3] = {1, 1, 1};
    CUresult r = ((enc_t)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d (query %d)\n", (int)r, (int)qr);
    CUtensorMap* dtm; CK(cudaMalloc(&dtm, sizeof(tm))); CK(cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    for (int mode = 0; mode < 2; mode++)
        for (int shift = 0; shift < 2; shift++) {
            if (mode == 0) probe<0><<<1, 32, 1024>>>(tm, dtm, shift); else probe<1><<<1, 32, 1024>>>(tm, dtm, shift);
            cudaError_t e = cudaDeviceSynchronize();
            printf("mode %d (descriptor in %s) shift %d -> %s\n", mode, mode ? "global" : "param", shift, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
            // plane 1, row 2, x = shift .. : source plane 0 row 1 x = -shift .. + 1
            printf("   out[p1][r2][0..3] = %g %g %g %g  (expect src+1: row1 = %d..)\n", h[nx * ny + 2 * nx + 0], h[nx * ny + 2 * nx + 1], h[nx * ny + 2 * nx + 2], h[nx * ny + 2 * nx + 3], nx);
        }
    printf("TMA_PROBE OK\n");
    return 0;
}
