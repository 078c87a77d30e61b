// Issue-rate microbenchmark of the integer / packed-16-bit instructions the mode-2 kernels are made of (sm_100a).
// Every kind runs 8 independent dependency chains per thread, 1024 threads per SM x 2 CTAs, and reports
// warp-instructions per clock per SM (4 = one instruction per scheduler per clock) and Tinstr/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_int ubench_int.cu && ./ubench_int
#include <cuda_runtime.h>

#include <cstdio>

enum Kind { K_IADD3, K_LOP3, K_VIMNMX, K_VIADDMNMX, K_VADD2, K_VMAXS2, K_VIBMAX_SEL, K_VIADDMAX2, K_SEL, K_IMAD, K_PRMT, K_SHF,
            K_MIX_ROW, K_MIX_NEW, K_IMADHI, K_COUNT };
static const char* names[] = {"IADD3", "LOP3", "VIMNMX", "VIADDMNMX", "VIADD.16x2", "VIMNMX.S16x2", "VIMNMX.S16x2+pred+2SEL(3 instr)",
                              "VIADDMNMX.S16x2", "SEL(+ISETP)", "IMAD", "PRMT", "SHF", "row mix (packed DP cell pair)", "new mix (biased fields, arithmetic flags)", "IMAD.HI"};

template <int KIND>
__global__ void __launch_bounds__(512) k(unsigned* out, int iters, unsigned seed) {
    unsigned a[8], b0 = seed | 1u, b1 = seed * 7u + 3u;
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x * (2 * j + 1) + seed + j;
    unsigned acc = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned o = a[(j + 1) & 7];
                if (KIND == K_IADD3) a[j] = a[j] + b0 + o;
                if (KIND == K_LOP3) a[j] = a[j] ^ (o & b0);
                if (KIND == K_VIMNMX) a[j] = (unsigned)max((int)a[j], (int)o);
                if (KIND == K_VIADDMNMX) a[j] = (unsigned)__viaddmax_s32((int)a[j], (int)b0, (int)o);
                if (KIND == K_VADD2) a[j] = __vadd2(a[j], o);
                if (KIND == K_VMAXS2) a[j] = __vmaxs2(a[j], o);
                if (KIND == K_VIBMAX_SEL) {
                    bool ph, pl;
                    a[j] = __vibmax_s16x2(a[j], o, &ph, &pl);
                    acc += (pl ? b0 : 0u) + (ph ? b1 : 0u);  // 2 SEL + 1 IADD3
                }
                if (KIND == K_VIADDMAX2) a[j] = __viaddmax_s16x2(a[j], b0, o);
                if (KIND == K_SEL) a[j] = ((int)a[j] > (int)o) ? b0 : b1 + a[j];
                if (KIND == K_IMAD) a[j] = a[j] * b0 + o;
                if (KIND == K_PRMT) a[j] = __byte_perm(a[j], o, 0x5432);
                if (KIND == K_SHF) a[j] = __funnelshift_r(a[j], o, 7);
                if (KIND == K_IMADHI) a[j] = __umulhi(a[j], b0) + o;
                if (KIND == K_MIX_NEW) {
                    // biased unsigned 16-bit fields: plain 32-bit adds, U16x2 maxima, flags from bit 15 / 31 of a + K - b
                    const unsigned Kc = 0x80008000u, Mk = 0x80008000u;
                    const unsigned sh = 1u << (17 + (j & 7));
                    const unsigned um2 = a[j] + b0;
                    const unsigned yv = __vmaxu2(um2, o);
                    const unsigned fy = o + (Kc - 0x00010001u) - um2;
                    acc = __umulhi(fy & Mk, sh) + acc;
                    const unsigned dd = o + b1;
                    const unsigned h = __vmaxu2(dd, yv);
                    const unsigned xl = __viaddmax_u16x2(h, b0, h + b1);
                    const unsigned x = __vmaxu2(xl, b1);
                    const unsigned t = __vmaxu2(dd, x);
                    const unsigned fd = dd + Kc - x;
                    const unsigned m = __vmaxu2(t, yv);
                    const unsigned ft = t + Kc - yv;
                    const unsigned fx = x + (Kc - 0x00050005u) - m;
                    const unsigned fb = m + Kc - a[(j + 2) & 7];
                    const unsigned bs = __vmaxu2(m, a[(j + 2) & 7]);
                    acc2 = __umulhi(fd & Mk, sh) + acc2;
                    acc3 = __umulhi(ft & Mk, sh) + acc3;
                    acc4 = __umulhi(fx & Mk, sh) + acc4;
                    acc5 = __umulhi(fb & Mk, sh) + acc5;
                    a[j] = bs + yv;
                }
                if (KIND == K_MIX_ROW) {
                    // one packed pair of the DP row: 6 VIADD.16x2, 7 VIMNMX.S16x2 (5 with flags), 1 VIADDMNMX.S16x2
                    const unsigned um2 = __vadd2(a[j], b0);
                    bool p0, p1, p2, p3, p4, p5, p6, p7, p8, p9;
                    const unsigned yv = __vibmax_s16x2(um2, o, &p0, &p1);
                    const unsigned dd = __vadd2(o, b1);
                    const unsigned h = __vmaxs2(dd, yv);
                    const unsigned xl = __viaddmax_s16x2(h, b0, __vadd2(h, b1));
                    const unsigned x = __vmaxs2(xl, b1);
                    const unsigned t = __vibmax_s16x2(dd, x, &p2, &p3);
                    const unsigned m = __vibmax_s16x2(t, yv, &p4, &p5);
                    (void)__vibmax_s16x2(x, __vadd2(m, b0), &p6, &p7);
                    const unsigned bs = __vibmax_s16x2(m, a[j], &p8, &p9);
                    acc |= (p0 ? 1u : 0u) | (p1 ? 2u : 0u) | (p2 ? 4u : 0u) | (p3 ? 8u : 0u) | (p4 ? 16u : 0u) | (p5 ? 32u : 0u) |
                           (p6 ? 64u : 0u) | (p7 ? 128u : 0u) | (p8 ? 256u : 0u) | (p9 ? 512u : 0u);
                    a[j] = __vadd2(bs, yv);
                }
            }
        }
    }
    unsigned r = acc ^ acc2 ^ acc3 ^ acc4 ^ acc5;
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= a[j];
    if (r == 0x12345678u) out[0] = r;
}

template <int KIND>
static void run(unsigned* d_out, int sms, double clk_ghz, double instr_per_step) {
    const int blocks = sms * 2, threads = 512, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<KIND><<<blocks, threads>>>(d_out, iters, (unsigned)rep + 1u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double warp_instr = (double)blocks * threads / 32.0 * iters * 64.0 * instr_per_step;
    const double per_s = warp_instr * 32.0 / (best * 1e-3);
    printf("%-36s %8.3f ms  %7.2f Tthread-instr/s  %5.2f warp-instr/clk/SM (at %.3f GHz)\n", names[KIND], best, per_s / 1e12,
           warp_instr / (best * 1e-3) / (clk_ghz * 1e9) / sms, clk_ghz);
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    unsigned* d_out;
    cudaMalloc(&d_out, 4);
    printf("SMs %d, clock %.3f GHz (nominal max)\n", sms, ghz);
    run<K_IADD3>(d_out, sms, ghz, 1);
    run<K_LOP3>(d_out, sms, ghz, 1);
    run<K_VIMNMX>(d_out, sms, ghz, 1);
    run<K_VIADDMNMX>(d_out, sms, ghz, 1);
    run<K_VADD2>(d_out, sms, ghz, 1);
    run<K_VMAXS2>(d_out, sms, ghz, 1);
    run<K_VIBMAX_SEL>(d_out, sms, ghz, 3.125);
    run<K_VIADDMAX2>(d_out, sms, ghz, 1);
    run<K_SEL>(d_out, sms, ghz, 3);
    run<K_IMAD>(d_out, sms, ghz, 1);
    run<K_PRMT>(d_out, sms, ghz, 1);
    run<K_SHF>(d_out, sms, ghz, 1);
    run<K_MIX_ROW>(d_out, sms, ghz, 29);
    run<K_MIX_NEW>(d_out, sms, ghz, 27);
    run<K_IMADHI>(d_out, sms, ghz, 1);
    cudaFree(d_out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
