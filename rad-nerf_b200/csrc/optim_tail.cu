// optim_tail.cu -- the tail of a training step as ONE sweep over every trainable tensor.
//
// The reference ends each step with  scaler.step(optimizer); scaler.update(); [ema.update()]  (nerf/utils.py:1171-1182)
// on torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15) over ~60 tensors in 7-9 parameter groups (main.py:204,
// nerf/network.py:329-361): GradScaler's unscale pass (read + write every gradient), a host sync on found_inf, Adam's
// foreach chain (lerp, mul, addcmul, sqrt, div, add, addcdiv: seven sweeps with temporaries) and, before the next
// backward, zero_grad (one more write of every gradient).  Here: one kernel reads p, g, m, v once and writes p, m, v
// (and g = 0) once -- 32 B per element, HBM-bound; the inf flag is read on the device (no host sync), the step counters
// live on the device, bias corrections are evaluated in double per CTA.
//
// Work decomposition: a tensor is cut into chunks of RN_ADAM_CHUNK elements, one CTA per chunk; the descriptor of tensor t
// carries the index of its first chunk, a CTA finds its tensor by binary search (<= 6 probes for 60 tensors).
#include "common.cuh"

namespace rn {
namespace {

constexpr uint32_t kChunk = RN_ADAM_CHUNK;   // elements per CTA
constexpr uint32_t kThreads = 256;
constexpr uint32_t kVecPerThread = kChunk / (4 * kThreads);   // float4 per thread per array

struct GroupTable { rn_adam_group g[RN_ADAM_MAX_GROUPS]; };

struct Hyper {               // per-CTA constants, fp32 exactly as torch rounds its Python doubles
    float one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, neg_step_size, bc2_sqrt, inv_scale;
    int skip;
};

__device__ __forceinline__ uint32_t find_tensor(const rn_adam_tensor* __restrict__ t, uint32_t n, uint32_t chunk) {
    uint32_t lo = 0, hi = n;   // last tensor with first_chunk <= chunk
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (t[mid].first_chunk <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

// torch/optim/adam.py _single_tensor_adam, in its operation order and with the roundings of torch's own kernels:
//   grad = grad.add(param, alpha=weight_decay)                       g + wd * p       (one fma)
//   exp_avg.lerp_(grad, 1 - beta1)                                   fma(g - m, w, m) (Lerp.h, |w| < 0.5)
//   exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)     fma((1-b2) * g, g, v * b2)
//   denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)    IEEE sqrt, division, add
//   param.addcdiv_(exp_avg, denom, value=-step_size)                 p + (-step_size * m) / denom
// every operation is spelled with an explicit-rounding intrinsic so the CPU checker (oracle.c o_adam_step) is bit-identical
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const Hyper& h) {
    g = __fmul_rn(g, h.inv_scale);
    if (h.weight_decay != 0.0f) g = __fmaf_rn(h.weight_decay, p, g);
    m = __fmaf_rn(__fsub_rn(g, m), h.one_minus_beta1, m);
    v = __fmaf_rn(__fmul_rn(h.one_minus_beta2, g), g, __fmul_rn(v, h.beta2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), h.bc2_sqrt), h.eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(h.neg_step_size, m), denom));
}

__global__ void __launch_bounds__(kThreads)
adam_tail_kernel(const rn_adam_tensor* __restrict__ tensors, uint32_t n_tensors, GroupTable groups,
                 const rn_adam_group* __restrict__ groups_dev, const float* __restrict__ grad_scale,
                 const float* __restrict__ found_inf, uint32_t zero_grads) {
    __shared__ rn_adam_tensor T;
    __shared__ Hyper H;
    if (threadIdx.x == 0) {
        const uint32_t t = find_tensor(tensors, n_tensors, blockIdx.x);
        T = tensors[t];
        const rn_adam_group G = groups_dev ? groups_dev[T.group] : groups.g[T.group];
        const double step = (double)(*T.step) + 1.0;
        const double bc1 = 1.0 - pow(G.beta1, step);
        const double bc2 = 1.0 - pow(G.beta2, step);
        H.one_minus_beta1 = (float)(1.0 - G.beta1);
        H.beta2 = (float)G.beta2;
        H.one_minus_beta2 = (float)(1.0 - G.beta2);
        H.eps = (float)G.eps;
        H.weight_decay = (float)G.weight_decay;
        H.neg_step_size = (float)(-(G.lr / bc1));
        H.bc2_sqrt = (float)sqrt(bc2);
        H.inv_scale = grad_scale ? 1.0f / *grad_scale : 1.0f;     // GradScaler scales are powers of two: exact either way
        H.skip = found_inf ? (*found_inf != 0.0f) : 0;
    }
    __syncthreads();
    const Hyper h = H;
    const uint64_t base = (uint64_t)(blockIdx.x - T.first_chunk) * kChunk;
    const uint32_t count = (uint32_t)min((uint64_t)kChunk, T.n - base);
    float* __restrict__ p = T.param + base;
    float* __restrict__ g = T.grad + base;
    float* __restrict__ m = T.exp_avg + base;
    float* __restrict__ v = T.exp_avg_sq + base;
    const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
    uint32_t done = 0;
    if (aligned) {
        const uint32_t n4 = count >> 2;
        float4 P[kVecPerThread], Gr[kVecPerThread], M[kVecPerThread], V[kVecPerThread];
        if (!h.skip) {
            // all loads of the chunk in flight before the first use (16 x 16 B per thread)
#pragma unroll
            for (uint32_t k = 0; k < kVecPerThread; ++k) {
                const uint32_t i = k * kThreads + threadIdx.x;
                if (i < n4) {
                    P[k] = reinterpret_cast<const float4*>(p)[i];
                    Gr[k] = reinterpret_cast<const float4*>(g)[i];
                    M[k] = reinterpret_cast<const float4*>(m)[i];
                    V[k] = reinterpret_cast<const float4*>(v)[i];
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < kVecPerThread; ++k) {
                const uint32_t i = k * kThreads + threadIdx.x;
                if (i < n4) {
                    adam_one(P[k].x, Gr[k].x, M[k].x, V[k].x, h);
                    adam_one(P[k].y, Gr[k].y, M[k].y, V[k].y, h);
                    adam_one(P[k].z, Gr[k].z, M[k].z, V[k].z, h);
                    adam_one(P[k].w, Gr[k].w, M[k].w, V[k].w, h);
                    reinterpret_cast<float4*>(p)[i] = P[k];
                    reinterpret_cast<float4*>(m)[i] = M[k];
                    reinterpret_cast<float4*>(v)[i] = V[k];
                }
            }
        }
        if (zero_grads) {
#pragma unroll
            for (uint32_t k = 0; k < kVecPerThread; ++k) {
                const uint32_t i = k * kThreads + threadIdx.x;
                if (i < n4) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        done = n4 << 2;
    }
    for (uint32_t i = done + threadIdx.x; i < count; i += kThreads) {
        if (!h.skip) {
            float pp = p[i], mm = m[i], vv = v[i];
            adam_one(pp, g[i], mm, vv, h);
            p[i] = pp; m[i] = mm; v[i] = vv;
        }
        if (zero_grads) g[i] = 0.0f;
    }
}

// the group table of the next launches, written into device memory in stream order (see rn_adam_groups_store)
__global__ void adam_store_groups_kernel(GroupTable groups, uint32_t n, rn_adam_group* __restrict__ dst) {
    if (threadIdx.x < n) dst[threadIdx.x] = groups.g[threadIdx.x];
}

// step += 1 for every tensor of the launch unless the step was skipped (torch's fused Adam rolls its device-side step
// back on found_inf the same way)
__global__ void adam_advance_kernel(const rn_adam_tensor* __restrict__ tensors, uint32_t n_tensors,
                                    const float* __restrict__ found_inf) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tensors) return;
    if (found_inf && *found_inf != 0.0f) return;
    *tensors[t].step += 1.0f;
}

// torch_ema.ExponentialMovingAverage.update (requirements.txt: torch-ema; ema.py `update`):
//   tmp = s_param - param; tmp.mul_(one_minus_decay); s_param.sub_(tmp)      -- one launch for all tensors instead of 3 each
__global__ void __launch_bounds__(kThreads)
ema_update_kernel(const rn_adam_tensor* __restrict__ tensors, uint32_t n_tensors, float one_minus_decay) {
    __shared__ rn_adam_tensor T;
    if (threadIdx.x == 0) T = tensors[find_tensor(tensors, n_tensors, blockIdx.x)];
    __syncthreads();
    const uint64_t base = (uint64_t)(blockIdx.x - T.first_chunk) * kChunk;
    const uint32_t count = (uint32_t)min((uint64_t)kChunk, T.n - base);
    const float* __restrict__ p = T.param + base;
    float* __restrict__ s = T.ema + base;
    uint32_t done = 0;
    if (((((uintptr_t)p) | ((uintptr_t)s)) & 15) == 0) {
        const uint32_t n4 = count >> 2;
        for (uint32_t i = threadIdx.x; i < n4; i += kThreads) {
            const float4 a = reinterpret_cast<const float4*>(p)[i];
            float4 b = reinterpret_cast<float4*>(s)[i];
            b.x = __fsub_rn(b.x, __fmul_rn(__fsub_rn(b.x, a.x), one_minus_decay));
            b.y = __fsub_rn(b.y, __fmul_rn(__fsub_rn(b.y, a.y), one_minus_decay));
            b.z = __fsub_rn(b.z, __fmul_rn(__fsub_rn(b.z, a.z), one_minus_decay));
            b.w = __fsub_rn(b.w, __fmul_rn(__fsub_rn(b.w, a.w), one_minus_decay));
            reinterpret_cast<float4*>(s)[i] = b;
        }
        done = n4 << 2;
    }
    for (uint32_t i = done + threadIdx.x; i < count; i += kThreads)
        s[i] = __fsub_rn(s[i], __fmul_rn(__fsub_rn(s[i], p[i]), one_minus_decay));
}

}  // namespace
}  // namespace rn

using namespace rn;

static int fill_group_table(GroupTable& table, const rn_adam_group* groups, uint32_t n_groups) {
    RN_REQUIRE(groups, "null pointer");
    RN_REQUIRE(n_groups >= 1 && n_groups <= RN_ADAM_MAX_GROUPS, "between 1 and RN_ADAM_MAX_GROUPS parameter groups");
    for (uint32_t i = 0; i < n_groups; ++i) {
        RN_REQUIRE(groups[i].beta1 >= 0.0 && groups[i].beta1 < 1.0 && groups[i].beta2 >= 0.0 && groups[i].beta2 < 1.0,
                   "betas must lie in [0, 1)");
        RN_REQUIRE(groups[i].eps >= 0.0 && groups[i].weight_decay >= 0.0, "eps and weight_decay must be >= 0");
    }
    for (uint32_t i = 0; i < RN_ADAM_MAX_GROUPS; ++i) table.g[i] = groups[i < n_groups ? i : 0];
    return RN_OK;
}

extern "C" int rn_adam_step(const rn_adam_tensor* tensors, uint32_t n_tensors, uint32_t n_chunks, const rn_adam_group* groups,
                            uint32_t n_groups, const float* grad_scale, const float* found_inf, uint32_t flags, void* stream) {
    if (n_tensors == 0 || n_chunks == 0) return RN_OK;
    RN_REQUIRE(tensors && groups, "null pointer");
    GroupTable table = {};
    const rn_adam_group* on_device = nullptr;
    if (flags & RN_ADAM_GROUPS_ON_DEVICE) {
        RN_REQUIRE(n_groups >= 1 && n_groups <= RN_ADAM_MAX_GROUPS, "between 1 and RN_ADAM_MAX_GROUPS parameter groups");
        on_device = groups;
    } else {
        const int rc = fill_group_table(table, groups, n_groups);
        if (rc != RN_OK) return rc;
    }
    adam_tail_kernel<<<n_chunks, kThreads, 0, (cudaStream_t)stream>>>(tensors, n_tensors, table, on_device, grad_scale, found_inf,
                                                                        flags & RN_ADAM_ZERO_GRADS);
    int rc = finish_launch("rn_adam_step");
    if (rc != RN_OK) return rc;
    adam_advance_kernel<<<div_up(n_tensors, 128u), 128, 0, (cudaStream_t)stream>>>(tensors, n_tensors, found_inf);
    return finish_launch("rn_adam_step (advance)");
}

extern "C" int rn_adam_groups_store(const rn_adam_group* groups, uint32_t n_groups, rn_adam_group* groups_dev, void* stream) {
    RN_REQUIRE(groups_dev, "null pointer");
    GroupTable table = {};
    const int rc = fill_group_table(table, groups, n_groups);
    if (rc != RN_OK) return rc;
    adam_store_groups_kernel<<<1, RN_ADAM_MAX_GROUPS, 0, (cudaStream_t)stream>>>(table, n_groups, groups_dev);
    return finish_launch("rn_adam_groups_store");
}

extern "C" int rn_ema_update(const rn_adam_tensor* tensors, uint32_t n_tensors, uint32_t n_chunks, double decay, void* stream) {
    if (n_tensors == 0 || n_chunks == 0) return RN_OK;
    RN_REQUIRE(tensors, "null pointer");
    RN_REQUIRE(decay >= 0.0 && decay <= 1.0, "decay must lie in [0, 1]");
    ema_update_kernel<<<n_chunks, kThreads, 0, (cudaStream_t)stream>>>(tensors, n_tensors, (float)(1.0 - decay));
    return finish_launch("rn_ema_update");
}
