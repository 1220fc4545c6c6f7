// Integer-pipe micro-benchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the alignment kernels are bound by
// the SM integer pipes (DPX s16x2 min/max/add, LOP3, IADD3, SHF, PRMT, IMAD).  This program
// measures warp-instruction throughput per SM per clock for those ops so that bench.py can
// report the roofline against a measured integer peak.  Output: one JSON object on stdout.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo int_peak.cu -o int_peak
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHAINS 8
#define ITERS 4096

enum Op { OP_VIADDMNMX_RELU = 0, OP_VIMNMX3_RELU, OP_VIMNMX, OP_VIADD16X2, OP_PRMT, OP_LOP3, OP_IADD3, OP_IMAD,
          OP_SHF, OP_ADD64, OP_MIX_DPX_IMAD, OP_MIX_SW, OP_MYERS, OP_MYERS_ILP2, OP_COUNT };

static const char *op_names[OP_COUNT] = {"viaddmnmx_s16x2_relu", "vimnmx3_s16x2_relu", "vimnmx_s16x2", "viadd_16x2", "prmt",
                                         "lop3", "iadd3", "imad", "shf_funnel", "add_u64", "mix_dpx_imad",
                                         "mix_sw_cell", "myers_word_step", "myers_word_step_ilp2"};
// lane-ops issued per loop iteration and chain (for the throughput figure)
// (myers_word_step*: ONE unit per iteration and chain = one 64-row Myers/Hyyro word step, edlib.cpp:411-443, with all
//  operands in registers - the figure is word steps per second, the roof of the edit-distance kernels)
static const int op_count_per_iter[OP_COUNT] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 7, 1, 1};

template <int OP>
__global__ void __launch_bounds__(256) peak_kernel(uint32_t *out, uint32_t seed)
{
    uint32_t a[CHAINS], b[CHAINS];
    unsigned long long w[CHAINS];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        a[c] = seed + tid * 31u + c;
        b[c] = seed * 7u + c * 13u + tid;
        w[c] = ((unsigned long long)a[c] << 32) | b[c];
    }
    const uint32_t k1 = seed | 0x00010001u, k2 = (seed >> 3) | 0x00030002u;
    uint32_t carry_p[CHAINS], carry_m[CHAINS];
    int score = 0;
#pragma unroll
    uint64_t eqv[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { carry_p[c] = 0; carry_m[c] = 0; eqv[c] = (uint64_t)(tid * 2654435761u + c * 40503u) * 0x9E3779B97F4A7C15ull; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == OP_VIADDMNMX_RELU) a[c] = __viaddmax_s16x2_relu(a[c], k1, b[c]);
            else if (OP == OP_VIMNMX3_RELU) a[c] = __vimax3_s16x2_relu(a[c], k1, b[c]);
            else if (OP == OP_VIMNMX) a[c] = __vmaxs2(a[c], b[c] ^ k1);
            else if (OP == OP_VIADD16X2) a[c] = __vadd2(a[c], k1);
            else if (OP == OP_PRMT) a[c] = __byte_perm(a[c], b[c], 0x5410 + (k1 & 0x2222));
            else if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(k1));   // one 3-input LOP3
            else if (OP == OP_IADD3) a[c] = a[c] + b[c] + k1;
            else if (OP == OP_IMAD) a[c] = a[c] * k1 + b[c];
            else if (OP == OP_SHF) a[c] = __funnelshift_l(a[c], b[c], k1 & 31);
            else if (OP == OP_ADD64) w[c] = w[c] + (unsigned long long)k2 + ((unsigned long long)k1 << 32);
            else if (OP == OP_MIX_DPX_IMAD) {
                a[c] = __viaddmax_s16x2_relu(a[c], k1, b[c]);
                b[c] = b[c] * k2 + k1;
            } else if (OP == OP_MYERS || OP == OP_MYERS_ILP2) {
                // one Myers/Hyyro word step on a 64-bit word (the kernels' myers_word, ed_kernels.cuh): Pv = w[c], Mv = (a, b),
                // the match vector is a rotating register value, the horizontal carries travel from chain to chain like they
                // travel from word to word of a column.  CHAINS independent words per thread (8: what a thread of the
                // multi-word kernels holds); the ILP2 variant runs two independent columns of four words.
                uint64_t Pv = w[c], Mv = ((uint64_t)a[c] << 32) | b[c];
                uint64_t Eq = eqv[c] ^ (uint64_t)(uint32_t)it;   // a match vector that changes per column at the cost of one LOP3
                const int src = (OP == OP_MYERS_ILP2) ? ((c & 3) == 0 ? -1 : c - 1) : (c == 0 ? -1 : c - 1);
                const uint32_t hp = src < 0 ? 1u : carry_p[src], hm = src < 0 ? 0u : carry_m[src];
                const uint64_t Xv = Eq | Mv;
                Eq |= (uint64_t)hm;
                const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                uint64_t Ph = Mv | ~(Xh | Pv);
                uint64_t Mh = Pv & Xh;
                carry_p[c] = (uint32_t)(Ph >> 63);
                carry_m[c] = (uint32_t)(Mh >> 63);
                Ph = (Ph << 1) | (uint64_t)hp;
                Mh = (Mh << 1) | (uint64_t)hm;
                Pv = Mh | ~(Xv | Ph);
                Mv = Ph & Xv;
                w[c] = Pv; a[c] = (uint32_t)(Mv >> 32); b[c] = (uint32_t)Mv;
                score += (int)carry_p[c] - (int)carry_m[c];
            } else if (OP == OP_MIX_SW) {
                // the op mix of one packed Smith-Waterman cell pair (see sw_kernels.cuh)
                uint32_t u = __viaddmax_s16x2_relu(b[c], k1, a[c]);
                uint32_t s = __byte_perm(a[c], b[c], 0x5410);
                uint32_t hp = __viaddmax_s16x2_relu(a[c], s, u);
                uint32_t hm = __vadd2(hp, k2);
                uint32_t h = __vmaxs2(hp, b[c]);
                b[c] = __viaddmax_s16x2(b[c], k1, hm);
                a[c] = __vmaxs2(a[c], h);
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= a[c] ^ b[c] ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32) ^ carry_p[c] ^ carry_m[c];
    acc ^= (uint32_t)score;
    if (acc == 0x12345678u) out[tid] = acc; // keep the chains alive
}

template <int OP>
static double run_one(uint32_t *d_out, int sms, int clock_khz)
{
    const int threads = 256, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) peak_kernel<OP><<<blocks, threads>>>(d_out, 12345u + i);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        peak_kernel<OP><<<blocks, threads>>>(d_out, 777u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double lane_ops = (double)blocks * threads * (double)ITERS * CHAINS * op_count_per_iter[OP];
    return lane_ops / (best * 1e-3); // lane-ops per second
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no device\n"); return 1; }
    uint32_t *d_out;
    cudaMalloc(&d_out, sizeof(uint32_t) * 256 * prop.multiProcessorCount * 8);
    double r[OP_COUNT];
    r[0] = run_one<0>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[1] = run_one<1>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[2] = run_one<2>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[3] = run_one<3>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[4] = run_one<4>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[5] = run_one<5>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[6] = run_one<6>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[7] = run_one<7>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[8] = run_one<8>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[9] = run_one<9>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[10] = run_one<10>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[11] = run_one<11>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[12] = run_one<12>(d_out, prop.multiProcessorCount, prop.clockRate);
    r[13] = run_one<13>(d_out, prop.multiProcessorCount, prop.clockRate);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"lane_ops_per_s\": {", prop.name, prop.multiProcessorCount,
           prop.clockRate);
    for (int i = 0; i < OP_COUNT; ++i) printf("%s\"%s\": %.4e", i ? ", " : "", op_names[i], r[i]);
    printf("}, \"lanes_per_clk_per_sm_at_max_clock\": {");
    for (int i = 0; i < OP_COUNT; ++i)
        printf("%s\"%s\": %.2f", i ? ", " : "", op_names[i], r[i] / ((double)prop.multiProcessorCount * prop.clockRate * 1e3));
    printf("}}\n");
    cudaFree(d_out);
    return 0;
}
