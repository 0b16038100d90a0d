// tmem_probe.cu -- microbenchmark behind the round-2 redesign of k_beam_encode:
// can tensor memory (TMEM, 256 KB/SM) serve as a lane-private operand store for a NON-tensor-core kernel, read
// with tcgen05.ld while the shared-memory pipe is saturated by the quantile gathers?
//
//   1. semantics: tcgen05.st / tcgen05.ld .32x32b -- thread i of warp w addresses TMEM lane 32*(w%4)+i; warps with the
//      same w%4 see the same data; columns are addressed warp-uniformly.
//   2. throughput of tcgen05.ld.32x32b.x4 alone, of a bank-conflicted LDS gather loop alone, and of both interleaved
//      (do the two pipes overlap?), with 12 warps per SM and one CTA per SM like the real kernel.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu ; run: ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// ties the loaded registers to the wait so that the compiler cannot use them before it
__device__ __forceinline__ void tmem_wait_ld4(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) :: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- semantics
__global__ void k_semantics(uint32_t* out, int* errors)
{
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&s_base, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = s_base;
    const uint32_t q = warp & 3;
    const uint32_t my = base + ((q * 32u) << 16);
    // warps 0..3 write columns 0..511 of their quarter: value = (quarter, lane, column)
    if (warp < 4) {
        for (int c = 0; c < 512; c += 4) {
            const uint32_t v = (q << 24) | ((uint32_t)lane << 16);
            tmem_st4(my + c, v | (c + 0), v | (c + 1), v | (c + 2), v | (c + 3));
        }
        tmem_wait_st();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    // every warp (also warps 4..11: same quarter as warp%4) reads back
    int bad = 0;
    for (int c = 0; c < 512; c += 4) {
        uint32_t a, b, cc, d;
        tmem_ld4(my + c, a, b, cc, d);
        tmem_wait_ld4(a, b, cc, d);
        const uint32_t v = (q << 24) | ((uint32_t)lane << 16);
        bad += (a != (v | (c + 0))) + (b != (v | (c + 1))) + (cc != (v | (c + 2))) + (d != (v | (c + 3)));
    }
    if (bad) atomicAdd(errors, bad);
    if (tid == 0) out[0] = base;
    // overwrite one column from a warp of the second layer (warp 4 + q), read from warp q
    __syncthreads();
    if (warp >= 4 && warp < 8) { tmem_st4(my + 100, 7u, 8u, 9u, 10u + lane); tmem_wait_st(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp < 4) {
        uint32_t a, b, cc, d;
        tmem_ld4(my + 100, a, b, cc, d);
        tmem_wait_ld4(a, b, cc, d);
        if (a != 7u || b != 8u || cc != 9u || d != 10u + lane) atomicAdd(errors, 1000);
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(base, 512);
}

// ---------------------------------------------------------------- throughput
// mode bit 0: LDS gathers (G per iteration, random word offsets -> ~3.5-way conflicts; or `spread` = conflict-free)
// mode bit 1: tcgen05.ld.32x32b.x4 (L per iteration)
// mode bit 2: FP work (F FFMA per gathered value)
template <int MODE, int G, int L>
__global__ void __launch_bounds__(384, 1) k_tp(const uint32_t* __restrict__ offs, int iters, int conflict_free, float* out, long long* cycles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_base;
    float* tab = reinterpret_cast<float*>(smem);                 // 30018 floats = 120 KB like T2
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 30018; i += blockDim.x) tab[i] = (float)(i & 1023) * 1e-3f;
    if (warp == 0) tmem_alloc(&s_base, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t my = s_base + (((warp & 3) * 32u) << 16);
    if (warp < 4) {
        for (int c = 0; c < 512; c += 4) tmem_st4(my + c, c, c + 1, c + 2, c + 3);
        tmem_wait_st();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t o[G > 0 ? G : 1];
#pragma unroll
    for (int g = 0; g < G; ++g) o[g] = conflict_free ? (uint32_t)(4 * (lane + 32 * g)) : 4u * (offs[(tid * G + g) & 65535] % 20000u);
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
    uint32_t tsum = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        uint32_t r[L > 0 ? L : 1][4];
        if (MODE & 2) {
#pragma unroll
            for (int l = 0; l < L; ++l) tmem_ld4(my + ((it * 4 + l * 16) & 508), r[l][0], r[l][1], r[l][2], r[l][3]);
        }
        if (MODE & 1) {
            const uint32_t cb = (uint32_t)(it & 255) * 4u;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float v = *reinterpret_cast<const float*>(smem + o[g] + cb);
                if (MODE & 4) {
                    float d = v + acc[g & 3];
                    acc[g & 3] = fmaf(fmaf(0.5f, d, 0.25f), d, acc[g & 3]);
                } else {
                    acc[g & 3] += v;
                }
            }
        }
        if (MODE & 2) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int l = 0; l < L; ++l) tsum += r[l][0] ^ r[l][1] ^ r[l][2] ^ r[l][3];
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc[0] + acc[1] + acc[2] + acc[3] + (float)tsum;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_base, 512);
}

template <int MODE, int G, int L>
static void run_tp(const char* name, const uint32_t* d_offs, int conflict_free, float* d_out, long long* d_cyc)
{
    const int iters = 20000, grid = 148;
    const size_t smem = 30018 * 4;
    CK(cudaFuncSetAttribute(k_tp<MODE, G, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tp<MODE, G, L><<<grid, 384, smem>>>(d_offs, 1000, conflict_free, d_out, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_tp<MODE, G, L><<<grid, 384, smem>>>(d_offs, iters, conflict_free, d_out, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc[148];
    CK(cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    double c = 0;
    for (int i = 0; i < grid; ++i) c += (double)cyc[i];
    c /= grid;
    const double per_iter = c / iters;                     // cycles per iteration of the whole CTA (12 warps in parallel)
    printf("%-44s %8.3f ms  %9.1f cyc/iter", name, ms, per_iter);
    if (MODE & 1) printf("  | %6.3f cyc per warp-gather (12 warps x %d)", per_iter / (12.0 * (G > 0 ? G : 1)), G);
    if (MODE & 2) printf("  | %6.3f cyc per warp tcgen05.ld.x4 (12 x %d) = %6.1f B/cyc/SM", per_iter / (12.0 * (L > 0 ? L : 1)), L, 12.0 * L * 512.0 / per_iter);
    printf("\n");
}

int main()
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    printf("device: %s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    uint32_t* d_out; int* d_err;
    CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_err, 0, 4));
    k_semantics<<<1, 384>>>(d_out, d_err);
    CK(cudaDeviceSynchronize());
    int err = -1; uint32_t base = 0;
    CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&base, d_out, 4, cudaMemcpyDeviceToHost));
    printf("semantics: tmem base 0x%08x, mismatches %d  (%s)\n", base, err, err == 0 ? "lane-private, shared by warps with equal warp%4: OK" : "FAILED");

    uint32_t* h = (uint32_t*)malloc(65536 * 4);
    uint32_t x = 12345u;
    for (int i = 0; i < 65536; ++i) { x = x * 1664525u + 1013904223u; h[i] = x >> 8; }
    uint32_t* d_offs; float* d_o; long long* d_c;
    CK(cudaMalloc(&d_offs, 65536 * 4)); CK(cudaMalloc(&d_o, 148 * 384 * 4)); CK(cudaMalloc(&d_c, 148 * 8));
    CK(cudaMemcpy(d_offs, h, 65536 * 4, cudaMemcpyHostToDevice));

    run_tp<1, 40, 0>("LDS gather, random banks", d_offs, 0, d_o, d_c);
    run_tp<1, 40, 0>("LDS gather, conflict-free", d_offs, 1, d_o, d_c);
    run_tp<2, 0, 4>("tcgen05.ld.x4 alone (4/iter)", d_offs, 0, d_o, d_c);
    run_tp<2, 0, 14>("tcgen05.ld.x4 alone (14/iter)", d_offs, 0, d_o, d_c);
    run_tp<3, 40, 4>("gather random + 4 tcgen05.ld.x4", d_offs, 0, d_o, d_c);
    run_tp<3, 40, 14>("gather random + 14 tcgen05.ld.x4", d_offs, 0, d_o, d_c);
    run_tp<3, 40, 14>("gather conflict-free + 14 tcgen05.ld.x4", d_offs, 1, d_o, d_c);
    run_tp<5, 40, 0>("gather random + fp", d_offs, 0, d_o, d_c);
    run_tp<7, 40, 5>("gather random + fp + 5 tcgen05.ld.x4", d_offs, 0, d_o, d_c);
    run_tp<7, 120, 14>("gather random + fp + 14 ld.x4 per 120", d_offs, 0, d_o, d_c);
    run_tp<5, 120, 0>("gather random + fp (120)", d_offs, 0, d_o, d_c);
    return err == 0 ? 0 : 1;
}
