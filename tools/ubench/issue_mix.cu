// microbenchmark: what does an instruction of each kind cost next to FP64 work on one SM sub-partition?
// Per loop iteration every warp issues NF independent DFMAs (4 chains) interleaved with NX instructions of one
// other kind (4 independent chains, inline PTX so that the kind is what SASS shows); 8 warps per scheduler
// (4 CTAs x 256 threads per SM, the shape of k_chain_eval), so dependent latency is hidden.
// Printed: cycles per warp-iteration and scheduler.  An FP64 warp instruction holds the 16-lane pipe 2 cycles.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>
enum { IMAD, LOP3, ISETP_SEL, LDS64, LDS128, MUFU, F2I, SHFL, VOTE, IADD, FFMA, PRMT, BRA, LDC };
static const char* kNames[] = {"IMAD", "LOP3", "ISETP+SEL (2)", "LDS.64", "LDS.128", "MUFU.RCP64H", "F2I.F64", "SHFL", "VOTE", "IADD3", "FFMA", "PRMT", "BRA (taken, uniform) + ISETP", "LDC"};
__constant__ unsigned kc[64];
template <int OP, int NF, int NX>
__global__ void __launch_bounds__(256, 4) k(double* out, int iters, double a, double b, unsigned ia)
{
    // pointer-chase table: entry (row k, lane l) holds the shared-memory address of entry (row k+1 mod 8, lane l);
    // rows are 32 x 16 bytes (LDS.128, conflict-free) -- the 64-bit variant uses the same entries (2 wavefronts)
    __shared__ uint4 sh[8 * 32];
    const unsigned shbase = (unsigned)__cvta_generic_to_shared(sh);
    {
        const int k = threadIdx.x >> 5, l = threadIdx.x & 31;
        sh[threadIdx.x] = make_uint4(shbase + ((((k + 1) & 7) * 32 + l) << 4), 0, 0, 0);
    }
    __syncthreads();
    double x[4];
    unsigned u[4];
    float f[4];
    double d[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { x[j] = a + threadIdx.x + j; u[j] = ia + j + threadIdx.x; f[j] = (float)a + j; d[j] = a + j; }
    unsigned p[4];
#pragma unroll
    for (int j = 0; j < 4; j++) p[j] = shbase + (((2 * j) * 32 + (threadIdx.x & 31)) << 4);
    for (int i = 0; i < iters; i++) {
        constexpr int R = NF > NX ? NF : NX;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int c = r & 3;
            if (r < NF) x[c] = fma(x[c], b, a);
            if (r < NX) {
                if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(ia), "r"(u[(c + 1) & 3]));
                if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(ia), "r"(u[(c + 1) & 3]));
                if (OP == ISETP_SEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(u[c]) : "r"(ia), "r"(u[(c + 1) & 3]));
                if (OP == LDS64) asm volatile("{.reg .u32 q; ld.shared.v2.u32 {%0, q}, [%0];}" : "+r"(p[c]));
                if (OP == LDS128) asm volatile("{.reg .u32 q, r, s; ld.shared.v4.u32 {%0, q, r, s}, [%0];}" : "+r"(p[c]));
                if (OP == MUFU) { int hi = __double2hiint(d[c]); asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d[c]) : "d"(__hiloint2double(hi, 0))); }
                if (OP == F2I) { int q; asm volatile("cvt.rmi.s32.f64 %0, %1;" : "=r"(q) : "d"(d[c])); d[c] = __hiloint2double(__double2hiint(d[c]), q); }
                if (OP == SHFL) u[c] = __shfl_xor_sync(0xffffffffu, u[c], 1);
                if (OP == VOTE) u[c] = __ballot_sync(0xffffffffu, u[c] & 1);
                if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[c]) : "r"(u[(c + 1) & 3]));
                if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(f[(c + 1) & 3]), "f"((float)b));
                if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u[c]) : "r"(u[(c + 1) & 3]));
                if (OP == BRA) asm volatile("{.reg .pred p; setp.eq.u32 p, %0, 0x12345; @p bra.uni L%=; bra.uni M%=; L%=: add.u32 %0, %0, 1; M%=: }" : "+r"(u[c]));
                if (OP == LDC) u[c] ^= kc[(u[(c + 1) & 3] + r) & 63];
            }
        }
    }
    double s = x[0] + x[1] + x[2] + x[3] + u[0] + u[1] + u[2] + u[3] + f[0] + f[1] + f[2] + f[3] + d[0] + d[1] + d[2] + d[3] + p[0] + p[1] + p[2] + p[3];
    if (s == 123.456) out[0] = s;
}
template <int OP, int NF, int NX>
void run()
{
    static double* out = nullptr;
    if (!out) cudaMalloc(&out, 8);
    const int iters = 2000, nb = 148 * 4;
    k<OP, NF, NX><<<nb, 256>>>(out, 10, 1.0000001, 0.999999, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP, NF, NX><<<nb, 256>>>(out, iters, 1.0000001, 0.999999, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 8.0;  // 8 warps per scheduler
    printf("DFMA %3d + %-30s %3d : %7.1f cycles per warp-iteration and scheduler\n", NF, kNames[OP], NX, cyc);
}
template <int OP>
void sweep()
{
    run<OP, 0, 32>(); run<OP, 32, 8>(); run<OP, 32, 16>(); run<OP, 32, 32>(); run<OP, 32, 64>();
}
int main()
{
    run<IMAD, 32, 0>();
    sweep<IMAD>(); sweep<LOP3>(); sweep<IADD>(); sweep<ISETP_SEL>(); sweep<PRMT>(); sweep<FFMA>(); sweep<LDS64>(); sweep<LDS128>();
    sweep<MUFU>(); sweep<F2I>(); sweep<SHFL>(); sweep<VOTE>(); sweep<BRA>(); sweep<LDC>();
    return 0;
}
