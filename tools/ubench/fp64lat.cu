// microbenchmarks: DFMA dependent latency, throughput vs warps/ILP, co-issue with integer work
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int MIXI>
__global__ void k(double* out, long long* cyc, int iters, double a, double b, int ia)
{
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = a + threadIdx.x + j;
    int acc[4] = {ia, ia + 1, ia + 2, ia + 3};
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int j = 0; j < ILP; j++) x[j] = fma(x[j], b, a);
#pragma unroll
            for (int m = 0; m < MIXI; m++) acc[m & 3] = acc[m & 3] * 3 + ia;   // IMAD chain(s)
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += x[j];
    s += acc[0] + acc[1] + acc[2] + acc[3];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP, int MIXI>
void run(const char* name, int warps_per_sm)
{
    double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int iters = 4000;
    int threads = 32 * warps_per_sm;  // one block per SM
    int bs = threads > 1024 ? 1024 : threads; int nb = 148 * ((threads + bs - 1) / bs);
    k<ILP, MIXI><<<nb, bs>>>(out, cyc, 10, 1.0000001, 0.999999, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP, MIXI><<<nb, bs>>>(out, cyc, iters, 1.0000001, 0.999999, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double dfma_per_warp = (double)iters * 8 * ILP;
    double tf = (double)nb * bs * dfma_per_warp * 2 / (ms * 1e-3) * 1e-12;
    printf("%-28s warps/SM %3d ILP %d mixI %d : %7.2f cyc per DFMA-group(per warp), %6.2f TF, %.3f ms\n", name, warps_per_sm, ILP, MIXI,
           (double)c / (iters * 8), tf, ms);
}
int main()
{
    run<1, 0>("latency", 4);      // 1 warp per SMSP, dependent chain
    run<1, 0>("", 8); run<1, 0>("", 16); run<1, 0>("", 32); run<1,0>("",64);
    run<2, 0>("", 16); run<2, 0>("", 32);
    run<4, 0>("", 16); run<4, 0>("", 32);
    run<8, 0>("", 4); run<8, 0>("", 8); run<8, 0>("", 32);
    run<1, 1>("mix 1 IMAD per DFMA", 32); run<1, 2>("mix 2 IMAD per DFMA", 32); run<1, 1>("mix1", 64); run<1, 2>("mix2", 64);
    run<2, 2>("mix ilp2 1:1", 32); run<2, 4>("mix ilp2 2:1", 32);
    return 0;
}
