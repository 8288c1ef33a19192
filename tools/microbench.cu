// microbench.cu -- per-SM throughput probes that drive the blind-rotation design (run on the B200 via gpurun).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 4096
__constant__ double cW[16];

template <int MODE> __global__ void __launch_bounds__(256) k(double* out, const int* in, int iters) {
    __shared__ double2 sm[2048];
    const int t = threadIdx.x;
    for (int i = t; i < 2048; i += 256) sm[i] = make_double2(i, -i);
    __syncthreads();
    double a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5, a6 = t + 6, a7 = t + 7;
    int x0 = in[t], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double2 s = make_double2(0, 0);
    double c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, c4 = t + 4, c5 = t + 5, c6 = t + 6, c7 = t + 7;      // MODE 13
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // DFMA only
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
            a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
        } else if (MODE == 1) {     // DFMA with constant-bank operand
            a0 = fma(a0, cW[0], cW[1]); a1 = fma(a1, cW[2], cW[3]); a2 = fma(a2, cW[4], cW[5]); a3 = fma(a3, cW[6], cW[7]);
            a4 = fma(a4, cW[8], cW[9]); a5 = fma(a5, cW[10], cW[11]); a6 = fma(a6, cW[12], cW[13]); a7 = fma(a7, cW[14], cW[15]);
        } else if (MODE == 2) {     // LDS.128 only (conflict free), 8 per iter
#pragma unroll
            for (int u = 0; u < 8; u++) { double2 v = sm[(t + 256 * u + it) & 2047]; s.x += v.x; s.y += v.y; }
        } else if (MODE == 3) {     // 8 LDS.128 + 32 DFMA per iter (ratio of the 16-point design)
#pragma unroll
            for (int u = 0; u < 8; u++) { double2 v = sm[(t + 256 * u + it) & 2047]; a0 = fma(a0, v.x, v.y); }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9); a4 = fma(a4, 1.0000001, 1e-9);
                a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9); a1 = fma(a1, 1.0000002, 1e-9);
            }
        } else if (MODE == 4) {     // I2F.F64.S32 (8 per iter)
            a0 += (double)x0; a1 += (double)x1; a2 += (double)x2; a3 += (double)x3;
            a4 += (double)(x0 ^ it); a5 += (double)(x1 ^ it); a6 += (double)(x2 ^ it); a7 += (double)(x3 ^ it);
        } else if (MODE == 5) {     // magic-number int->double (8 per iter): LOP + DADD
            a0 += __hiloint2double(0x43300000, x0 ^ 0x80000000) - 4503601774854144.0;
            a1 += __hiloint2double(0x43300000, x1 ^ 0x80000000) - 4503601774854144.0;
            a2 += __hiloint2double(0x43300000, x2 ^ 0x80000000) - 4503601774854144.0;
            a3 += __hiloint2double(0x43300000, x3 ^ 0x80000000) - 4503601774854144.0;
            a4 += __hiloint2double(0x43300000, (x0 ^ it) ^ 0x80000000) - 4503601774854144.0;
            a5 += __hiloint2double(0x43300000, (x1 ^ it) ^ 0x80000000) - 4503601774854144.0;
            a6 += __hiloint2double(0x43300000, (x2 ^ it) ^ 0x80000000) - 4503601774854144.0;
            a7 += __hiloint2double(0x43300000, (x3 ^ it) ^ 0x80000000) - 4503601774854144.0;
        } else if (MODE == 6) {     // F2I.S64.F64.TRUNC (8 per iter)
            a0 = __longlong_as_double(__double_as_longlong(a0) ^ (long long)__double2ll_rz(a0)); a1 = __longlong_as_double(__double_as_longlong(a1) ^ (long long)__double2ll_rz(a1));
            a2 = __longlong_as_double(__double_as_longlong(a2) ^ (long long)__double2ll_rz(a2)); a3 = __longlong_as_double(__double_as_longlong(a3) ^ (long long)__double2ll_rz(a3));
            a4 = __longlong_as_double(__double_as_longlong(a4) ^ (long long)__double2ll_rz(a4)); a5 = __longlong_as_double(__double_as_longlong(a5) ^ (long long)__double2ll_rz(a5));
            a6 = __longlong_as_double(__double_as_longlong(a6) ^ (long long)__double2ll_rz(a6)); a7 = __longlong_as_double(__double_as_longlong(a7) ^ (long long)__double2ll_rz(a7));
        } else if (MODE == 7) {     // SHFL.XOR 32-bit (8 per iter)
            x0 = __shfl_xor_sync(0xffffffff, x0, 1) + 1; x1 = __shfl_xor_sync(0xffffffff, x1, 2) + 1; x2 = __shfl_xor_sync(0xffffffff, x2, 4) + 1; x3 = __shfl_xor_sync(0xffffffff, x3, 8) + 1;
            x0 = __shfl_xor_sync(0xffffffff, x0, 16) + 1; x1 = __shfl_xor_sync(0xffffffff, x1, 1) + 1; x2 = __shfl_xor_sync(0xffffffff, x2, 2) + 1; x3 = __shfl_xor_sync(0xffffffff, x3, 4) + 1;
        } else if (MODE == 8) {     // 8 SHFL + 8 LDS.128 per iter: do they share a pipe?
            x0 = __shfl_xor_sync(0xffffffff, x0, 1) + 1; x1 = __shfl_xor_sync(0xffffffff, x1, 2) + 1; x2 = __shfl_xor_sync(0xffffffff, x2, 4) + 1; x3 = __shfl_xor_sync(0xffffffff, x3, 8) + 1;
            x0 = __shfl_xor_sync(0xffffffff, x0, 16) + 1; x1 = __shfl_xor_sync(0xffffffff, x1, 1) + 1; x2 = __shfl_xor_sync(0xffffffff, x2, 2) + 1; x3 = __shfl_xor_sync(0xffffffff, x3, 4) + 1;
#pragma unroll
            for (int u = 0; u < 8; u++) { double2 v = sm[(t + 256 * u + it) & 2047]; s.x += v.x; s.y += v.y; }
        } else if (MODE == 9) {     // STS.128 + LDS.128 (8 each) per iter
#pragma unroll
            for (int u = 0; u < 8; u++) sm[(t + 256 * u) & 2047] = make_double2(a0 + u, a1);
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; u++) { double2 v = sm[(t * 8 + u + (t >> 2)) & 2047]; s.x += v.x; s.y += v.y; }
            a0 += s.x;
        } else if (MODE == 11) {    // DMMA m8n8k4 (256 FMA per warp instruction), 4 independent accumulators
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%4}, {%5}, {%0,%1};\n\t"
                         "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%4}, {%5}, {%2,%3};"
                         : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3) : "d"(a6), "d"(a7));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%4}, {%5}, {%0,%1};\n\t"
                         "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%4}, {%5}, {%2,%3};"
                         : "+d"(a4), "+d"(a5), "+d"(s.x), "+d"(s.y) : "d"(a6), "d"(a7));
        } else if (MODE == 12) {    // DMMA m16n8k16 (2048 FMA per warp instruction), 2 independent accumulators
            double b0 = a6, b1 = a7, b2 = a6 + 1, b3 = a7 + 1;
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3)
                         : "d"(b0), "d"(b1), "d"(b2), "d"(b3), "d"(b0), "d"(b1), "d"(b2), "d"(b3), "d"(b0), "d"(b1), "d"(b2), "d"(b3));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(a4), "+d"(a5), "+d"(s.x), "+d"(s.y)
                         : "d"(b0), "d"(b1), "d"(b2), "d"(b3), "d"(b0), "d"(b1), "d"(b2), "d"(b3), "d"(b0), "d"(b1), "d"(b2), "d"(b3));
        } else if (MODE == 13) {    // 4 DMMA m8n8k4 + 32 DFMA per iteration: do the tensor and the vector FP64 paths add up?
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%4}, {%5}, {%0,%1};\n\t"
                         "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%4}, {%5}, {%2,%3};"
                         : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3) : "d"(a6), "d"(a7));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%4}, {%5}, {%0,%1};\n\t"
                         "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%4}, {%5}, {%2,%3};"
                         : "+d"(a4), "+d"(a5), "+d"(s.x), "+d"(s.y) : "d"(a6), "d"(a7));
#pragma unroll
            for (int u = 0; u < 4; u++) {      // loop-carried (c0..c7 live across iterations), so the 32 DFMA stay in the loop
                c0 = fma(c0, 1.0000001, 1e-9); c1 = fma(c1, 1.0000001, 1e-9); c2 = fma(c2, 1.0000001, 1e-9); c3 = fma(c3, 1.0000001, 1e-9);
                c4 = fma(c4, 1.0000001, 1e-9); c5 = fma(c5, 1.0000001, 1e-9); c6 = fma(c6, 1.0000001, 1e-9); c7 = fma(c7, 1.0000001, 1e-9);
            }
        } else if (MODE == 10) {    // integer bit-trick double->int64 trunc (8 per iter), integer pipe only
#pragma unroll
            for (int u = 0; u < 8; u++) {
                uint64_t bits = (uint64_t)__double_as_longlong(a0) + (uint64_t)(x0 + u + it) * 0x10000000001ull;
                uint64_t val = (bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
                int tr = (int)((bits >> 52) & 0x7FF) - 1075;
                uint64_t v2 = tr > 0 ? (tr >= 64 ? 0 : val << tr) : (-tr >= 64 ? 0 : val >> -tr);
                x1 += (int)((bits >> 63) ? 0 - v2 : v2);
            }
        }
    }
    out[blockIdx.x * 256 + t] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + s.x + s.y + x0 + x1 + x2 + x3 + (MODE == 13 ? c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7 : 0.0);
}

template <int MODE> void run(const char* name, double ops_per_iter_per_thread, const char* unit) {
    double* out; int* in; cudaMalloc(&out, 148 * 8 * 256 * 8); cudaMalloc(&in, 1024); cudaMemset(in, 0, 1024);
    for (int occ : {1, 2, 4}) {     // CTAs of 256 threads per SM
        int grid = 148 * occ;
        k<MODE><<<grid, 256>>>(out, in, 64);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, in, ITERS); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double per_sm_per_clk = ops_per_iter_per_thread * ITERS * 256.0 * occ / (ms * 1e-3 * 1.965e9);
        printf("%-44s occ=%d warps/SM=%2d  %8.3f ms  %7.2f %s/clk/SM (at 1965 MHz)\n", name, occ, occ * 8, ms, per_sm_per_clk, unit);
    }
    cudaFree(out); cudaFree(in);
}
int main() {
    double h[16]; for (int i = 0; i < 16; i++) h[i] = (i & 1) ? 1e-9 : 1.0000001;
    cudaMemcpyToSymbol(cW, h, sizeof(h));
    run<0>("DFMA reg", 8, "fma");
    run<1>("DFMA const-bank operand", 8, "fma");
    run<2>("LDS.128", 8 * 16, "B");
    run<3>("8 LDS.128 + 40 DFMA", 40, "fma");
    run<4>("I2F.F64.S32 (+DADD)", 8, "cvt");
    run<5>("magic int->double (LOP+DADD+DADD)", 8, "cvt");
    run<6>("F2I.S64.F64 trunc", 8, "cvt");
    run<7>("SHFL.BFLY b32", 8 * 4, "B");
    run<8>("8 SHFL + 8 LDS.128", 8 * 4 + 8 * 16, "B");
    run<9>("8 STS.128 + 8 LDS.128", 16 * 16, "B");
    run<10>("int bit-trick f64->i64 trunc", 8, "cvt");
    // FP64 tensor path (north star: tensor cores only if a DFT-as-GEMM beats the butterflies).  fma counts per THREAD per iteration:
    // one m8n8k4 = 256 FMA per warp = 8 per thread; one m16n8k16 = 2048 per warp = 64 per thread.
    run<11>("DMMA m8n8k4 x4", 4 * 8, "fma");
    run<12>("DMMA m16n8k16 x2", 2 * 64, "fma");
    run<13>("4 DMMA m8n8k4 + 32 DFMA (both counted)", 4 * 8 + 32, "fma");
    return 0;
}
