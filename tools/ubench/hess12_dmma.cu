// A/B microbenchmark for SURVEY row (g) / north_star "tensor cores only for the batched 12 x 12 dense Hessian blocks":
// the per-tet congruence product H12 = P^T H9 P (P = dF/dx, 9 x 12; ref: libuipc finite_element/constitutions/stable_neo_hookean_3d.cu:111-165,
// fem_utils.cu dFdx) for a batch of tetrahedra, float64,
//   A  one tet per THREAD on the FP64 pipe (DFMA), using P's Kronecker structure P[(3b+a),(3v+c)] = W[v][b] delta_ac
//      (what a register kernel does: 594 DFMA per tet instead of the dense 2268),
//   B  one tet per WARP on the tensor cores: mma.sync.aligned.m8n8k4.f64 (DMMA), T = H9 P (12 DMMA on zero-padded 16 x 16 x 12 tiles),
//      T through shared memory into B-fragments, H12 = P^T T (12 DMMA).
// Both read the same L2-resident inputs ([tet][81] Hessians, [tet][12] shape-gradient rows) and write the full 12 x 12 result; B is
// checked against A on the host. Prints ns per tet and tets per second for both with one CTA of 256 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hess12_dmma hess12_dmma.cu && ./hess12_dmma
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int THREADS = 256; // 8 warps per SM, up to 255 registers per thread: the DFMA arm keeps a whole 9 x 9 Hessian in registers

// reps == 1: full 12 x 12 results to `out` (correctness). reps > 1: the product is repeated on rescaled shape gradients with the inputs
// held on chip and only a checksum leaves the SM, so that the arithmetic pipes -- not HBM -- are what is timed.
__global__ void __launch_bounds__(THREADS, 1) k_dfma(const double* __restrict__ H9, const double* __restrict__ Wg, double* __restrict__ out, int nt, int reps)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
        const double* H = H9 + (size_t)81 * t;
        double W[4][3];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) W[v][b] = Wg[(size_t)12 * t + 3 * v + b];
        double* o = out + (size_t)144 * t;
        double chk = 0.0;
#pragma unroll 1
        for (int rep = 0; rep < reps; ++rep) {
        if (rep) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int b = 0; b < 3; ++b) W[v][b] *= 1.0009765625;
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { // rows 3 b + a of T = H9 P(:, w): the only ones the (., a) entries of the blocks need
                double T[3][3];       // T[b][c] = sum_b' H[(3b+a),(3b'+c)] W[w][b']
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int r = 3 * b + a;
                        T[b][c] = H[r * 9 + c] * W[w][0] + H[r * 9 + 3 + c] * W[w][1] + H[r * 9 + 6 + c] * W[w][2];
                    }
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (v > w) continue;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const double s = W[v][0] * T[0][c] + W[v][1] * T[1][c] + W[v][2] * T[2][c];
                        if (reps == 1) {
                            o[(3 * v + a) * 12 + 3 * w + c] = s;
                            o[(3 * w + c) * 12 + 3 * v + a] = s;
                        } else
                            chk += s;
                    }
                }
            }
        }
        }
        if (reps > 1) o[0] = chk;
    }
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(THREADS, 1) k_dmma(const double* __restrict__ H9, const double* __restrict__ Wg, double* __restrict__ out, int nt, int reps)
{
    __shared__ double Ts[THREADS / 32][16][17];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int r8 = lane >> 2, k4 = lane & 3;
    for (int t = blockIdx.x * wpb + warp; t < nt; t += gridDim.x * wpb) {
        const double* H = H9 + (size_t)81 * t;
        const double* W = Wg + (size_t)12 * t;
        // P fragments: pf[ki][ni] = P[4 ki + k4][8 ni + r8] (B operand of step 1 == A operand, transposed, of step 2)
        double pf[3][2];
#pragma unroll
        for (int ki = 0; ki < 3; ++ki)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int r = 4 * ki + k4, q = 8 * ni + r8;
                pf[ki][ni] = (r < 9 && q < 12 && (r % 3) == (q % 3)) ? W[3 * (q / 3) + r / 3] : 0.0;
            }
        double hf[2][3]; // H9 fragments (A operand of step 1)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) {
                const int r = 8 * mi + r8, c = 4 * ki + k4;
                hf[mi][ki] = (r < 9 && c < 9) ? H[r * 9 + c] : 0.0;
            }
        double d[2][2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) d[mi][ni][0] = d[mi][ni][1] = 0.0;
#pragma unroll 1
        for (int rep = 0; rep < reps; ++rep) {
        if (rep) {
#pragma unroll
            for (int ki = 0; ki < 3; ++ki)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) pf[ki][ni] *= 1.0009765625;
        }
        double acc[2][2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) {
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], hf[mi][ki], pf[ki][ni]);
            }
        // T -> shared memory -> B fragments of step 2
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                Ts[warp][8 * mi + r8][8 * ni + 2 * k4] = acc[mi][ni][0];
                Ts[warp][8 * mi + r8][8 * ni + 2 * k4 + 1] = acc[mi][ni][1];
            }
        __syncwarp();
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
            double tf[2];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) tf[ni] = Ts[warp][4 * ki + k4][8 * ni + r8];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma(d[mi][ni][0], d[mi][ni][1], pf[ki][mi], tf[ni]);
        }
        __syncwarp();
        }
        double* o = out + (size_t)144 * t;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int r = 8 * mi + r8, c = 8 * ni + 2 * k4;
                if (r < 12 && c < 12) { o[r * 12 + c] = d[mi][ni][0]; o[r * 12 + c + 1] = d[mi][ni][1]; }
            }
    }
}

int main()
{
    const int nt = 148 * THREADS * 8; // 682 k tets: 315 tet-chunks of the 2160-tet gel
    std::vector<double> H((size_t)81 * nt), W((size_t)12 * nt);
    srand(1);
    for (int t = 0; t < nt; ++t) {
        double A[81];
        for (int i = 0; i < 81; ++i) A[i] = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) H[(size_t)81 * t + 9 * i + j] = A[9 * i + j] + A[9 * j + i];
        for (int i = 0; i < 12; ++i) W[(size_t)12 * t + i] = 400.0 * (rand() / (double)RAND_MAX - 0.5);
    }
    double *dH, *dW, *oA, *oB;
    cudaMalloc(&dH, H.size() * 8); cudaMalloc(&dW, W.size() * 8); cudaMalloc(&oA, (size_t)144 * nt * 8); cudaMalloc(&oB, (size_t)144 * nt * 8);
    cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(oA, 0, (size_t)144 * nt * 8); cudaMemset(oB, 0, (size_t)144 * nt * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float msA = 0, msB = 0, msA1 = 0, msB1 = 0;
    const int REPS = 64;
    for (int rep = 0; rep < 3; ++rep) { // on-chip repetitions: the arithmetic pipes are timed
        cudaEventRecord(e0); k_dfma<<<148, THREADS>>>(dH, dW, oA, nt, REPS); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&msA, e0, e1);
        cudaEventRecord(e0); k_dmma<<<148, THREADS>>>(dH, dW, oB, nt, REPS); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&msB, e0, e1);
    }
    for (int rep = 0; rep < 3; ++rep) { // one product per tet, full results written: HBM-bound for both
        cudaEventRecord(e0); k_dfma<<<148, THREADS>>>(dH, dW, oA, nt, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&msA1, e0, e1);
        cudaEventRecord(e0); k_dmma<<<148, THREADS>>>(dH, dW, oB, nt, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&msB1, e0, e1);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    std::vector<double> a((size_t)144 * 4096), b((size_t)144 * 4096);
    cudaMemcpy(a.data(), oA, a.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), oB, b.size() * 8, cudaMemcpyDeviceToHost);
    double worst = 0, mag = 0;
    for (size_t i = 0; i < a.size(); ++i) { worst = fmax(worst, fabs(a[i] - b[i])); mag = fmax(mag, fabs(a[i])); }
    const double bytes = (81.0 + 12.0 + 144.0) * 8.0 * nt;
    const double prodA = (double)nt * REPS / (msA * 1e-3), prodB = (double)nt * REPS / (msB * 1e-3);
    printf("{\"tets\": %d, \"on_chip_reps\": %d, \"dfma_products_per_s\": %.4e, \"dmma_products_per_s\": %.4e, \"dmma_over_dfma_time\": %.3f, "
           "\"dfma_ns_per_product_per_sm\": %.3f, \"dmma_ns_per_product_per_sm\": %.3f, "
           "\"hbm_bound_single_product\": {\"dfma_ms\": %.4f, \"dmma_ms\": %.4f, \"dfma_gbs\": %.1f, \"dmma_gbs\": %.1f}, "
           "\"max_abs_diff\": %.3e, \"max_abs\": %.3e}\n",
           nt, REPS, prodA, prodB, msB / msA, 148e9 / prodA, 148e9 / prodB, msA1, msB1, bytes / msA1 * 1e-6, bytes / msB1 * 1e-6, worst, mag);
    return 0;
}
