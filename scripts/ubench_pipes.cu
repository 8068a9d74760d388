// ubench_pipes.cu — issue-rate microbenchmarks that decide how the decode / pack inner loops are
// written (which pipe an instruction class runs on and at what rate on sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/ubench_pipes scripts/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 256
#define ITER 4096

template <int K>
__global__ void __launch_bounds__(128) kern(uint32_t *out, uint32_t seed, unsigned long long *cycles)
{
    __shared__ uint16_t lut[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) lut[i] = (uint16_t)((i * 2654435761u) >> 20);
    __syncthreads();
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1, c = seed ^ 0x55aa, d = threadIdx.x * 7 + 1;
    uint32_t e = a ^ b, f = c + d, g = a + 11, h = b + 13;
    unsigned long long p64[4] = {a * 77ull + 1, c * 91ull + 3, e * 13ull + 5, f * 17ull + 7};
    uint32_t base;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(base) : "l"(lut));
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int r = 0; r < REP / 8; r++) {
            if (K == 0) {  // SHF funnel, variable shift, 8 independent chains
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(e) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(f) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(g) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(h) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
            } else if (K == 1) {  // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c) : "r"(b), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e) : "r"(b), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(f) : "r"(b), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(b), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h) : "r"(b), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c) : "r"(b), "r"(d));
            } else if (K == 2) {  // IMAD (mad.lo)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(e) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(f) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(g) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(h) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
            } else if (K == 3) {  // IMAD.HI (mul.hi)
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(c) : "r"(b));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(e) : "r"(b));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(f) : "r"(b));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(g) : "r"(b));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(h) : "r"(b));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(c) : "r"(b));
            } else if (K == 4) {  // IMAD.WIDE (mad.wide.u32), 4 chains of 64-bit, data-dependent multiplicand
#define MW(v) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(v) : "r"(d))
                MW(p64[0]); MW(p64[1]); MW(p64[2]); MW(p64[3]); MW(p64[0]); MW(p64[1]); MW(p64[2]); MW(p64[3]);
            } else if (K == 5) {  // PRMT
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(e) : "r"(b), "r"(d));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(f) : "r"(b), "r"(d));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(g) : "r"(b), "r"(d));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(h) : "r"(b), "r"(d));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
            } else if (K == 6) {  // 1:1 mix SHF + IMAD
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(e) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(f) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(g) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(h) : "r"(b), "r"(d));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c) : "r"(b), "r"(d));
            } else if (K == 7) {  // 1:1 mix LOP3 + IMAD.WIDE
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e) : "r"(b), "r"(d));
                MW(p64[0]);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(f) : "r"(b), "r"(d));
                MW(p64[1]);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(b), "r"(d));
                MW(p64[2]);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h) : "r"(b), "r"(d));
                MW(p64[3]);
            } else if (K == 8) {  // LDS.U16, random banks, 8 independent
                uint32_t x0, x1, x2, x3;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x0) : "r"(base + ((a & 0xfff) << 1)));
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x1) : "r"(base + ((c & 0xfff) << 1)));
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x2) : "r"(base + ((e & 0xfff) << 1)));
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x3) : "r"(base + ((f & 0xfff) << 1)));
                a = a * 5 + x0; c = c * 5 + x1; e = e * 5 + x2; f = f * 5 + x3;
            } else if (K == 9) {  // dependent chain: shf -> lop3 -> lds.u16 -> shf (decode step latency)
                uint32_t x;
                asm volatile("shf.r.clamp.b32 %0, %1, 0, 18;" : "=r"(x) : "r"(a));
                asm volatile("lop3.b32 %0, %1, 0x1ffe, %2, 0xf8;" : "=r"(x) : "r"(x), "r"(base));
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(x));
                asm volatile("shf.l.wrap.b32 %0, %1, %0, %2;" : "+r"(a) : "r"(b), "r"(x));
            } else if (K == 10) {  // ISETP + SEL pairs
                asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(a) : "r"(b), "r"(d));
                asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(c) : "r"(b), "r"(d));
                asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(e) : "r"(b), "r"(d));
                asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(f) : "r"(b), "r"(d));
            } else if (K == 11) {  // ATOMS.OR spread addresses
                asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(base + ((a & 0x7ff) << 2)), "r"(b) : "memory");
                asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(base + ((c & 0x7ff) << 2)), "r"(b) : "memory");
                a = a * 5 + 1; c = c * 5 + 3;
            } else if (K == 12) {  // ATOMS.ADD spread addresses (histogram style)
                asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(base + ((a & 0xff) << 2)), "r"(1) : "memory");
                asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(base + ((c & 0xff) << 2)), "r"(1) : "memory");
                a = a * 5 + 1; c = c * 5 + 3;
            } else if (K == 13) {  // STS.32 conflict free + LDS.32 conflict free
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(base + (threadIdx.x & 31) * 4 + ((a & 15) << 7)), "r"(b) : "memory");
                uint32_t x;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(base + (threadIdx.x & 31) * 4 + ((c & 15) << 7)));
                a += x; c = c * 5 + 3;
            }
        }
    }
    const unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h ^ (uint32_t)(p64[0] ^ p64[1] ^ p64[2] ^ p64[3]) ^ (uint32_t)((p64[0] ^ p64[1] ^ p64[2] ^ p64[3]) >> 32);
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int K>
void run(const char *name, int per_iter, int ctas_per_sm, uint32_t *out, unsigned long long *cyc)
{
    const int grid = 148 * ctas_per_sm;
    kern<K><<<grid, 128>>>(out, 12345, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<K><<<grid, 128>>>(out, 12345, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // every CTA runs the same instruction count, all CTAs resident at once: kernel time = CTA time
    const double n_warp_inst_per_sm = (double)ITER * (REP / 8) * per_iter * ctas_per_sm * 4;
    const double mhz = c / (ms * 1e3);
    printf("%-34s warps/SMSP %d: %8.3f ms  %10llu cyc (%.0f MHz)  warp-inst/clk/SMSP %.3f\n", name, ctas_per_sm, ms, c,
           mhz, n_warp_inst_per_sm / c / 4);
}

int main()
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 16 * 128 * 4); cudaMalloc(&cyc, 8);
    for (int occ : {1, 4, 8}) {
        run<0>("SHF.L.W variable", 8, occ, out, cyc);
        run<1>("LOP3", 8, occ, out, cyc);
        run<2>("IMAD", 8, occ, out, cyc);
        run<3>("IMAD.HI", 8, occ, out, cyc);
        run<4>("IMAD.WIDE (+6 mov)", 8, occ, out, cyc);
        run<5>("PRMT", 8, occ, out, cyc);
        run<6>("SHF+IMAD 1:1", 8, occ, out, cyc);
        run<7>("LOP3+IMAD.WIDE 1:1", 8, occ, out, cyc);
        run<8>("LDS.U16 random (+3 ALU each)", 4, occ, out, cyc);
        run<9>("chain shf-lop3-lds-shf (4 inst)", 4, occ, out, cyc);
        run<10>("ISETP+SEL", 8, occ, out, cyc);
        run<11>("ATOMS.OR spread (+imad)", 2, occ, out, cyc);
        run<12>("ATOMS.ADD 256 bins (+imad)", 2, occ, out, cyc);
        run<13>("STS+LDS conflict-free", 2, occ, out, cyc);
    }
    return 0;
}
