// Micro-benchmarks behind the decoder's design decisions: one warp, dependent chains, clock64.
// Each kernel runs ITER iterations of an unrolled body and reports cycles per iteration.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096

// ---- 1. predicate as a SEL operand vs predicate as an instruction guard, on a dependent chain
__global__ void k_sel_chain(uint32_t *out, uint32_t x, uint32_t lim) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 a, b;\n\t"
                         "setp.lt.u32 p, %0, %1;\n\t"
                         "add.u32 a, %0, 3;\n\t"
                         "xor.b32 b, %0, 5;\n\t"
                         "selp.b32 %0, a, b, p;\n\t}" : "+r"(x) : "r"(lim));
    }
    long long t1 = clock64();
    out[0] = x;
    out[1] = (uint32_t)(t1 - t0);
}
__global__ void k_guard_chain(uint32_t *out, uint32_t x, uint32_t lim) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "setp.lt.u32 p, %0, %1;\n\t"
                         "@p add.u32 %0, %0, 3;\n\t"
                         "@!p xor.b32 %0, %0, 5;\n\t}" : "+r"(x) : "r"(lim));
    }
    long long t1 = clock64();
    out[0] = x;
    out[1] = (uint32_t)(t1 - t0);
}
// ---- 2. plain ALU / IMAD dependent latencies
__global__ void k_add_chain(uint32_t *out, uint32_t x, uint32_t y) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %1;" : "+r"(x) : "r"(y));
    }
    long long t1 = clock64();
    out[0] = x; out[1] = (uint32_t)(t1 - t0);
}
__global__ void k_mul_chain(uint32_t *out, uint32_t x, uint32_t y) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("mul.lo.u32 %0, %0, %1;\n\tshr.u32 %0, %0, 1;" : "+r"(x) : "r"(y));
    }
    long long t1 = clock64();
    out[0] = x; out[1] = (uint32_t)(t1 - t0);
}
// ---- 3. shared-memory pointer chase
__global__ void k_lds_chain(uint32_t *out, uint32_t start) {
    __shared__ uint32_t tab[256];
    for (int i = threadIdx.x; i < 256; i += 32) tab[i] = (uint32_t)((i * 4 + 68) & 1023);
    __syncwarp();
    uint32_t a = (uint32_t)__cvta_generic_to_shared(tab), x = start & 1020;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("{\n\t.reg .b32 t;\n\tadd.u32 t, %1, %0;\n\tld.shared.u32 %0, [t];\n\t}" : "+r"(x) : "r"(a) : "memory");
    }
    long long t1 = clock64();
    out[0] = x; out[1] = (uint32_t)(t1 - t0);
}
// ---- 4. branches on a dependent chain: never taken / always taken (skipping 6 instructions) / taken backward loop
__global__ void k_branch(uint32_t *out, uint32_t x, uint32_t lim, uint32_t y) {
    // lim = 0xFFFFFFFF: p always true (branch always taken); lim = 0: never taken
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "setp.lt.u32 p, %0, %1;\n\t"
                         "add.u32 %0, %0, 3;\n\t"
                         "@p bra.uni SKIP;\n\t"
                         "xor.b32 %0, %0, %2;\n\t"
                         "add.u32 %0, %0, %2;\n\t"
                         "xor.b32 %0, %0, 7;\n\t"
                         "add.u32 %0, %0, %2;\n\t"
                         "xor.b32 %0, %0, 9;\n\t"
                         "add.u32 %0, %0, %2;\n\t"
                         "SKIP:\n\t"
                         "and.b32 %0, %0, 0xffffff;\n\t}" : "+r"(x) : "r"(lim), "r"(y));
    }
    long long t1 = clock64();
    out[0] = x; out[1] = (uint32_t)(t1 - t0);
}
// ---- 5. ALU throughput of two warps on one sub-partition: all 32 lanes vs 16 lanes active
__global__ void k_alu_tput(uint32_t *out, uint32_t x, uint32_t y, int half) {
    if (half && (threadIdx.x & 31) >= 16) return;
    uint32_t a = x, b = x + 1, c = x + 2, d = x + 3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("add.u32 %0, %0, %4;\n\tadd.u32 %1, %1, %4;\n\tadd.u32 %2, %2, %4;\n\tadd.u32 %3, %3, %4;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(y));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = a + b + c + d; out[1] = (uint32_t)(t1 - t0); }
}

static void report(const char *name, uint32_t *d, int per_iter_ops) {
    uint32_t h[2];
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-46s %8.2f cycles / step (%d steps per iteration)\n", name, (double)h[1] / ITER / per_iter_ops, per_iter_ops);
}

int main() {
    uint32_t *d;
    cudaMalloc(&d, 64);
    for (int rep = 0; rep < 2; rep++) {
        k_sel_chain<<<1, 32>>>(d, 1, 0x7fffffff);   if (rep) report("setp -> selp (predicate as data) chain", d, 8);
        k_guard_chain<<<1, 32>>>(d, 1, 0x7fffffff); if (rep) report("setp -> @p op (predicate as guard) chain", d, 8);
        k_add_chain<<<1, 32>>>(d, 1, 3);            if (rep) report("add -> xor (2 dependent ALU ops)", d, 8);
        k_mul_chain<<<1, 32>>>(d, 3, 5);            if (rep) report("mul -> shr (IMAD + ALU)", d, 8);
        k_lds_chain<<<1, 32>>>(d, 0);               if (rep) report("add -> ld.shared pointer chase", d, 8);
        k_branch<<<1, 32>>>(d, 1, 0, 3);            if (rep) report("chain with a branch never taken", d, 8);
        k_branch<<<1, 32>>>(d, 1, 0xffffffffu, 3);  if (rep) report("chain with a forward branch always taken", d, 8);
        k_alu_tput<<<1, 32>>>(d, 1, 3, 0);          if (rep) report("4 independent adds x8, 1 warp, 32 lanes", d, 8);
        k_alu_tput<<<1, 256>>>(d, 1, 3, 0);         if (rep) report("4 independent adds x8, 8 warps (2/SMSP), 32 lanes", d, 8);
        k_alu_tput<<<1, 256>>>(d, 1, 3, 1);         if (rep) report("4 independent adds x8, 8 warps (2/SMSP), 16 lanes", d, 8);
        k_alu_tput<<<1, 512>>>(d, 1, 3, 0);         if (rep) report("4 independent adds x8, 16 warps (4/SMSP), 32 lanes", d, 8);
        k_alu_tput<<<1, 512>>>(d, 1, 3, 1);         if (rep) report("4 independent adds x8, 16 warps (4/SMSP), 16 lanes", d, 8);
    }
    return 0;
}
