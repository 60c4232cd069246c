// Fused step metrics: everything Base_Lightning.calculate_losses_step derives from (y_pred, y) besides the
// gradient-carrying loss, in two launches and without a host sync.
//
//  Replaces, for the classification head (gnnLightning.py:L131-151, L285-348; customMetrics.py:L6-54):
//    softmax per foot -> 16-class conversion (product of per-foot probabilities) -> argmax -> MulticlassAccuracy;
//    per-foot argmax -> four BinaryF1Score (the reference: 4 sklearn confusion matrices on the host + a python loop over B);
//    CrossEntropyLossMetric (sum-reduced CE / rows).
//  The 16-class argmax equals "every foot's 2-way argmax" (the joint probability factorises; torch.argmax takes the first
//  maximum, i.e. class 0 on a per-foot tie, which is what the per-foot rule below does too).
//  For the regression heads (L124-130): sum of squared / absolute errors -> MSE, RMSE, L1.
//  Deterministic: per-block partials in fixed order, summed by one final block; no atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mshgnn {

constexpr int MET_SLOTS = 32;
constexpr int MET_BLOCKS = 296;
// slots: classification  0 ce_sum  1 rows  2 graphs_all_feet_right  3 graphs  4+4*leg+{0,1,2,3} tp fp fn tn
//                        20 ce_mean (fp32-rounded like the reference)  21 accuracy  22..25 F1 of leg 0..3
//        regression      0 sse  1 sae  2 n   20 mse  21 rmse  22 l1

__device__ __forceinline__ double met_label(const void* p, int dtype, int64_t i) {
    if (dtype == 1) return ((const double*)p)[i];
    if (dtype == 2) return (double)((const long long*)p)[i];
    return (double)((const float*)p)[i];
}

__global__ void __launch_bounds__(256)
k_metrics_partial(const int kind, const int feet, const float* __restrict__ out, const void* __restrict__ labels, const int label_dtype,
                  const int64_t n, double* __restrict__ partial) {
    __shared__ double sh[8][MET_SLOTS];
    double a[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) a[i] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    if (kind == 1) {
        // n = graphs; out [n*feet, 2] logits, labels [n, feet] in {0,1}
        for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < n; g += stride) {
            bool all_ok = true;
            for (int f = 0; f < feet; ++f) {
                const int64_t r = g * feet + f;
                const float2 l = *reinterpret_cast<const float2*>(out + 2 * r);
                const int y = met_label(labels, label_dtype, r) != 0.0;
                const double l0 = (double)l.x, l1 = (double)l.y;
                const double m = l0 > l1 ? l0 : l1;
                a[0] += m + log(exp(l0 - m) + exp(l1 - m)) - (y ? l1 : l0);
                const int p = l.y > l.x;           // argmax of the 2-way softmax, first maximum on ties
                all_ok = all_ok && (p == y);
                if (f < 4) a[4 + 4 * f + (p ? (y ? 0 : 1) : (y ? 2 : 3))] += 1.0;
            }
            a[1] += (double)feet; a[2] += all_ok ? 1.0 : 0.0; a[3] += 1.0;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
            const double d = (double)out[i] - met_label(labels, label_dtype, i);
            a[0] += d * d; a[1] += fabs(d); a[2] += 1.0;
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 20; ++i) {
        double v = a[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[wid][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 20) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * MET_SLOTS + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(32)
k_metrics_final(const int kind, const double* __restrict__ partial, const int n_blocks, double* __restrict__ batch, double* __restrict__ epoch) {
    __shared__ double s[MET_SLOTS];
    const int t = threadIdx.x;
    double v = 0.0;
    if (t < 20)
        for (int b = 0; b < n_blocks; ++b) v += partial[(int64_t)b * MET_SLOTS + t];
    s[t] = v;
    __syncwarp();
    double d = 0.0;
    auto f1 = [&](int leg) {
        const double tp = s[4 + 4 * leg], fp = s[5 + 4 * leg], fn = s[6 + 4 * leg];
        const double pr = tp / (tp + fp), rc = tp / (tp + fn);
        const double f = 2.0 * (pr * rc) / (pr + rc);
        return f != f ? 0.0 : f;                   // torch.nan_to_num (customMetrics.py:L54)
    };
    if (kind == 1) {
        if (t == 20) d = (double)((float)s[0] / (float)s[1]);          // summed_loss.float() / total_num (customMetrics.py:L25)
        else if (t == 21) d = s[2] / s[3];
        else if (t >= 22 && t < 26) d = f1(t - 22);
    } else {
        if (t == 20) d = s[0] / s[2];
        else if (t == 21) d = sqrt(s[0] / s[2]);
        else if (t == 22) d = s[1] / s[2];
    }
    batch[t] = t < 20 ? v : d;
    if (epoch != nullptr && t < 20) epoch[t] += v;
}

}  // namespace mshgnn
