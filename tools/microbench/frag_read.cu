// Micro-benchmark: how fast can a B200 stream the encoder's feature tensor x[B * nodes, K] (fp32, row-major) when every CTA task
// reads a block of ROWS graph rows of ONE node slot (row stride nodes * K * 4 = 14.4 KB) restricted to a fragment of F columns?
//   F = 64 is the access pattern of the encoder kernels (one K block of 256 bytes per row at a time), larger F = longer contiguous runs.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o frag_read frag_read.cu ; run: ./frag_read
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 2)
k_read(const float4* __restrict__ x, float* __restrict__ out, int B, int nodes, int K4 /*K / 4*/, int F4 /*fragment, float4s*/, int rows,
       int n_frag, int n_rb, int order) {
    // task -> (node, fragment, row block); order 0: row blocks fastest (slot-major), 1: node fastest, 2: fragment fastest
    float acc = 0.f;
    const int n_tasks = nodes * n_frag * n_rb;
    for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        int node, frag, rb;
        if (order == 0) { rb = task % n_rb; frag = (task / n_rb) % n_frag; node = task / (n_rb * n_frag); }
        else if (order == 1) { node = task % nodes; frag = (task / nodes) % n_frag; rb = task / (nodes * n_frag); }
        else { frag = task % n_frag; node = (task / n_frag) % nodes; rb = task / (n_frag * nodes); }
        const int c0 = frag * F4, c1 = min(c0 + F4, K4), w = c1 - c0;
        const int r0 = rb * rows, r1 = min(r0 + rows, B);
        const int total = (r1 - r0) * w;
        for (int e0 = 0; e0 < total; e0 += 4 * 512) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 512 + threadIdx.x;
                const int r = e / w, c = e - r * w;
                v[u] = e < total ? __ldg(x + ((size_t)(r0 + r) * nodes + node) * K4 + c0 + c) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 1.2345f) out[0] = acc;
}

int main() {
    const int B = 16384, nodes = 4, K = 900, K4 = K / 4;
    const size_t n = (size_t)B * nodes * K;
    float* x; float* out;
    cudaMalloc(&x, n * 4); cudaMalloc(&out, 4);
    cudaMemset(x, 0, n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int F, rows; } cfgs[] = {{64, 256}, {128, 128}, {192, 96}, {256, 64}, {448, 36}, {900, 16}};
    for (int order = 0; order < 3; ++order)
    for (auto c : cfgs) {
        const int F4 = c.F / 4 > K4 ? K4 : c.F / 4;
        const int n_frag = (K4 + F4 - 1) / F4, n_rb = (B + c.rows - 1) / c.rows;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_read<<<148 * 2, 512>>>((const float4*)x, out, B, nodes, K4, F4, c.rows, n_frag, n_rb, order);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("order %d fragment %4d cols (%5d B) x %3d rows: %.3f ms  %.0f GB/s\n", order, c.F, F4 * 16, c.rows, ms, n * 4 / ms / 1e6);
    }
    // plain contiguous read of the same bytes
    {
        const int n_rb = (int)(n / 4 / (4 * 512 * 8));
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_read<<<148 * 2, 512>>>((const float4*)x, out, (int)(n / 4 / 16384), 1, 16384, 16384, 1, 1, (int)(n / 4 / 16384), 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("contiguous: %.3f ms  %.0f GB/s\n", ms, n * 4 / ms / 1e6);
        (void)n_rb;
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
