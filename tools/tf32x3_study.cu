// tf32x3_study.cu -- parity study for DESIGN.md 4.3: could the TargetNetwork contractions (K = 32..128 forward / dgrad,
// K = 2048 points for wgrad) run on the tensor cores as error-compensated 3xTF32 inside the 1e-5 bar against fp32 torch.mm?
//
// Measurement tool, not product code.  Builds with
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tf32x3_study tools/tf32x3_study.cu
// and runs on the GPU box.  For the reference network 3 -> 32 -> 64 -> 128 -> 64 -> 3 (model/target_network.py:31-38) with
// per-sample weights ~ N(0, 0.15^2) and input points ~ N(0, 0.6^2) (the test distribution of tests/test_target_network_gpu.py)
// it evaluates every layer three ways from the SAME fp32 inputs,
//   (1) fp64 accumulation                                   -- ground truth,
//   (2) fp32 FFMA chain in ascending k                      -- what the product kernel and torch.mm's fp32 path do,
//   (3) 3xTF32 on mma.sync.m16n8k8 (a = hi + lo, both tf32; lo*hi + hi*lo + hi*hi into ONE fp32 accumulator tile),
// and the wgrad contraction dW = Z^T A over the 2048 points of a sample the same three ways, and prints the worst error of
// (2) and (3) against (1) and of (3) against (2), normalised by the largest magnitude of the tensor (the tests' metric).
// mma.sync is the legacy tensor-core path of sm_100a; tcgen05 shares the fp32 accumulation data path in question (does it
// round or truncate, and how does that bias grow with K).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));               \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

// C[m][n] = act(bias[n] + sum_k A[m][k] * W[n][k]);  A [M][K], W [N][K] row-major, C [M][N]
__global__ void gemm_f64(int M, int N, int K, const float *A, const float *W, const float *bias, int relu, double *C64, float *C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int m = i / N, n = i % N;
    double s = bias ? (double)bias[n] : 0.0;
    for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)W[n * K + k];
    if (relu) s = s > 0 ? s : 0;
    C64[i] = s;
    if (C) C[i] = (float)s;
}
__global__ void gemm_f32(int M, int N, int K, const float *A, const float *W, const float *bias, int relu, float *C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int m = i / N, n = i % N;
    float s = bias ? bias[n] : 0.f;
    for (int k = 0; k < K; ++k) s = __fmaf_rn(A[m * K + k], W[n * K + k], s);
    if (relu) s = fmaxf(s, 0.f);
    C[i] = s;
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split(float x, uint32_t &hi, uint32_t &lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// one warp per 16 x 8 output tile; K padded with zeros; PASSES = 1: plain TF32, 3: error-compensated 3xTF32
template <int PASSES>
__global__ void gemm_tf32(int M, int N, int K, const float *A, const float *W, const float *bias, int relu, float *C) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int tn = (N + 7) / 8, tm = (M + 15) / 16;
    if (tile >= tm * tn) return;
    const int m0 = (tile / tn) * 16, n0 = (tile % tn) * 8;
    const int g = lane >> 2, t = lane & 3;
    float c[4];
    {
        const int nA = n0 + 2 * t, nB = nA + 1;
        const float bA = (bias && nA < N) ? bias[nA] : 0.f, bB = (bias && nB < N) ? bias[nB] : 0.f;
        c[0] = bA, c[1] = bB, c[2] = bA, c[3] = bB;
    }
    auto ld = [&](const float *P, int r, int R, int k) { return (r < R && k < K) ? P[r * K + k] : 0.f; };
    for (int k0 = 0; k0 < K; k0 += 8) {
        const float af[4] = {ld(A, m0 + g, M, k0 + t), ld(A, m0 + g + 8, M, k0 + t), ld(A, m0 + g, M, k0 + t + 4), ld(A, m0 + g + 8, M, k0 + t + 4)};
        const float bf[2] = {ld(W, n0 + g, N, k0 + t), ld(W, n0 + g, N, k0 + t + 4)};
        uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) split(af[i], ah[i], al[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) split(bf[i], bh[i], bl[i]);
        if (PASSES == 3) {  // small terms first
            mma_tf32(c, al, bh);
            mma_tf32(c, ah, bl);
        }
        mma_tf32(c, ah, bh);
    }
    const int r0 = m0 + g, r1 = r0 + 8, cA = n0 + 2 * t, cB = cA + 1;
    float v[4] = {c[0], c[1], c[2], c[3]};
    if (relu)
        for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
    if (r0 < M && cA < N) C[r0 * N + cA] = v[0];
    if (r0 < M && cB < N) C[r0 * N + cB] = v[1];
    if (r1 < M && cA < N) C[r1 * N + cA] = v[2];
    if (r1 < M && cB < N) C[r1 * N + cB] = v[3];
}

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static float randn() {  // Box-Muller on a 64-bit LCG
    auto u = []() {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        return ((rng >> 11) + 1) * (1.0 / 9007199254740993.0);
    };
    return (float)(sqrt(-2.0 * log(u())) * cos(6.283185307179586 * u()));
}

struct Err {
    double f32 = 0, t3 = 0, t1 = 0, t3_vs_f32 = 0;
};
static Err compare(int n, const double *ref, const float *f32, const float *t3, const float *t1) {
    double scale = 0;
    for (int i = 0; i < n; ++i) scale = fmax(scale, fabs(ref[i]));
    Err e;
    for (int i = 0; i < n; ++i) {
        e.f32 = fmax(e.f32, fabs(f32[i] - ref[i]) / scale);
        e.t3 = fmax(e.t3, fabs(t3[i] - ref[i]) / scale);
        e.t1 = fmax(e.t1, fabs(t1[i] - ref[i]) / scale);
        e.t3_vs_f32 = fmax(e.t3_vs_f32, fabs((double)t3[i] - (double)f32[i]) / scale);
    }
    return e;
}

int main() {
    const int P = 2048, dims[6] = {3, 32, 64, 128, 64, 3};
    std::vector<float> x(P * 3);
    for (auto &v : x) v = randn() * 0.6f;
    float *dA32, *dA3, *dA1, *dW, *dB, *dC32, *dC3, *dC1, *dRef32;
    double *dC64;
    const int MAXA = P * 128;
    CK(cudaMalloc(&dA32, MAXA * 4)); CK(cudaMalloc(&dA3, MAXA * 4)); CK(cudaMalloc(&dA1, MAXA * 4));
    CK(cudaMalloc(&dC32, MAXA * 4)); CK(cudaMalloc(&dC3, MAXA * 4)); CK(cudaMalloc(&dC1, MAXA * 4)); CK(cudaMalloc(&dRef32, MAXA * 4));
    CK(cudaMalloc(&dC64, MAXA * 8)); CK(cudaMalloc(&dW, 128 * 128 * 4)); CK(cudaMalloc(&dB, 128 * 4));
    CK(cudaMemcpy(dA32, x.data(), P * 3 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA3, x.data(), P * 3 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA1, x.data(), P * 3 * 4, cudaMemcpyHostToDevice));
    printf("TargetNetwork forward, every layer from the SAME fp32 inputs of its own chain (errors relative to the largest |value| of the layer output)\n");
    printf("%-22s %12s %12s %12s %14s\n", "layer", "fp32 vs f64", "3xTF32 vs f64", "TF32 vs f64", "3xTF32 vs fp32");
    std::vector<double> r64(MAXA);
    std::vector<float> c32(MAXA), c3(MAXA), c1(MAXA);
    for (int l = 0; l < 5; ++l) {
        const int K = dims[l], N = dims[l + 1], relu = l < 4;
        std::vector<float> w(N * K), b(N);
        for (auto &v : w) v = randn() * 0.15f;
        for (auto &v : b) v = randn() * 0.15f;
        CK(cudaMemcpy(dW, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
        const int total = P * N, tiles = ((P + 15) / 16) * ((N + 7) / 8);
        // per-layer errors: all three from the fp32 chain's input (isolates the layer); the chains continue on their own outputs
        gemm_f64<<<(total + 255) / 256, 256>>>(P, N, K, dA32, dW, dB, relu, dC64, nullptr);
        gemm_f32<<<(total + 255) / 256, 256>>>(P, N, K, dA32, dW, dB, relu, dC32);
        gemm_tf32<3><<<(tiles + 3) / 4, 128>>>(P, N, K, dA32, dW, dB, relu, dC3);
        gemm_tf32<1><<<(tiles + 3) / 4, 128>>>(P, N, K, dA32, dW, dB, relu, dC1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r64.data(), dC64, total * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c32.data(), dC32, total * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c3.data(), dC3, total * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c1.data(), dC1, total * 4, cudaMemcpyDeviceToHost));
        Err e = compare(total, r64.data(), c32.data(), c3.data(), c1.data());
        char name[64];
        snprintf(name, sizeof name, "L%d  K=%-3d -> %-3d", l + 1, K, N);
        printf("%-22s %12.2e %12.2e %12.2e %14.2e\n", name, e.f32, e.t3, e.t1, e.t3_vs_f32);
        // end-to-end chains: each variant feeds on its own previous output
        gemm_tf32<3><<<(tiles + 3) / 4, 128>>>(P, N, K, dA3, dW, dB, relu, dC3);
        gemm_tf32<1><<<(tiles + 3) / 4, 128>>>(P, N, K, dA1, dW, dB, relu, dC1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(dA32, dC32, total * 4, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(dA3, dC3, total * 4, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(dA1, dC1, total * 4, cudaMemcpyDeviceToDevice));
        if (l == 4) {
            CK(cudaMemcpy(c32.data(), dA32, total * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(c3.data(), dA3, total * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(c1.data(), dA1, total * 4, cudaMemcpyDeviceToHost));
            double scale = 0, d3 = 0, d1 = 0;
            for (int i = 0; i < total; ++i) scale = fmax(scale, fabs(c32[i]));
            for (int i = 0; i < total; ++i) d3 = fmax(d3, fabs(c3[i] - c32[i]) / scale), d1 = fmax(d1, fabs(c1[i] - c32[i]) / scale);
            printf("END TO END (5 layers, each chain on its own activations): 3xTF32 vs fp32 chain %.2e   single-pass TF32 vs fp32 chain %.2e   (bar: 1e-5)\n", d3, d1);
        }
    }
    // wgrad: dW[o][k] = sum_p Z[p][o] * A[p][k] over the 2048 points of a sample: a K = 2048 contraction (M = 64, N = 128)
    {
        const int O = 64, Kc = 128;
        std::vector<float> zt(O * P), at(Kc * P);  // stored transposed: rows = output channel / input channel, columns = points
        for (auto &v : zt) v = randn() * 0.3f;
        for (auto &v : at) v = fmaxf(randn() * 0.5f, 0.f);  // post-ReLU activations
        float *dZ, *dAt;
        CK(cudaMalloc(&dZ, zt.size() * 4)); CK(cudaMalloc(&dAt, at.size() * 4));
        CK(cudaMemcpy(dZ, zt.data(), zt.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAt, at.data(), at.size() * 4, cudaMemcpyHostToDevice));
        const int total = O * Kc, tiles = ((O + 15) / 16) * ((Kc + 7) / 8);
        gemm_f64<<<(total + 255) / 256, 256>>>(O, Kc, P, dZ, dAt, nullptr, 0, dC64, nullptr);
        gemm_f32<<<(total + 255) / 256, 256>>>(O, Kc, P, dZ, dAt, nullptr, 0, dC32);
        gemm_tf32<3><<<(tiles + 3) / 4, 128>>>(O, Kc, P, dZ, dAt, nullptr, 0, dC3);
        gemm_tf32<1><<<(tiles + 3) / 4, 128>>>(O, Kc, P, dZ, dAt, nullptr, 0, dC1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r64.data(), dC64, total * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c32.data(), dC32, total * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c3.data(), dC3, total * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c1.data(), dC1, total * 4, cudaMemcpyDeviceToHost));
        Err e = compare(total, r64.data(), c32.data(), c3.data(), c1.data());
        printf("%-22s %12.2e %12.2e %12.2e %14.2e\n", "wgrad K=2048 (64x128)", e.f32, e.t3, e.t1, e.t3_vs_f32);
        // the accumulator's rounding mode: add 2048 equal terms that are not representable sums
        double bias3 = 0, bias32 = 0;
        for (int i = 0; i < total; ++i) bias3 += (c3[i] - r64[i]), bias32 += (c32[i] - r64[i]);
        printf("mean signed error of the K=2048 sums (a truncating accumulator shows as a NEGATIVE-towards-zero bias on positive sums): 3xTF32 %.3e, fp32 chain %.3e, mean |value| %.3e\n",
               bias3 / total, bias32 / total, [&] { double s = 0; for (int i = 0; i < total; ++i) s += fabs(r64[i]); return s / total; }());
    }
    return 0;
}
