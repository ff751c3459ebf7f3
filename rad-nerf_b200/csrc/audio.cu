// audio.cu -- per-frame conditioning in ONE single-CTA kernel: AudioNet + AudioAttNet (nerf/network.py:10-67,170-185),
// the lip-smoothing EMA (nerf/renderer.py:190-194), and the per-frame-constant halves of the first MLP layers folded into
// fp32 bias vectors for the fused head / torso kernels ("hoisting": the audio code, eye value, individual codes and the
// encoded head pose are identical for every sample of a frame, so W[:, const part] @ const is computed once here).
//
// The reference runs ~25 tiny cuDNN / cuBLAS launches for this (~0.66 MMAC); here it is one launch.  Numerics follow
// fp16 autocast: weights and activations are rounded to fp16, sums are fp32, conv outputs are rounded before AND after
// the (fp16) bias add as PyTorch's cuDNN path does, Linear adds the bias before its single rounding, softmax and the
// attention-weighted sum are fp32.
#include "frame.cuh"

namespace rn {


namespace {

__device__ long long* g_audio_prof_dev = nullptr;
#define RN_STAMP(i) if (prof && threadIdx.x == 0) prof[i] = clock64();

__device__ __forceinline__ float h16(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.02f * x; }

// Conv1d(k=3, pad=1) + bias + LeakyReLU(0.02) for F frames; in [F][Cin][Lin] -> out [F][Cout][Lout].
// Work item o = (co, f, lo) with (f, lo) fastest: the lanes of a warp share the output channel, so weight loads are
// warp-uniform broadcasts.  Four input channels (12 taps) are fetched into registers before the FMAs -- written out
// explicitly because the compiler otherwise issues every load right before its use and exposes its latency 132 times.
__device__ void conv_layer(const float* __restrict__ in, float* __restrict__ out, const __half* __restrict__ w, const __half* __restrict__ b,
                           uint32_t F, uint32_t Cin, uint32_t Cout, uint32_t Lin, uint32_t stride) {
    const uint32_t Lout = (Lin + 2 - 3) / stride + 1;
    const uint32_t total = F * Cout * Lout;
    for (uint32_t o = threadIdx.x; o < total; o += blockDim.x) {
        const uint32_t lo = o % Lout, f = (o / Lout) % F, co = o / (Lout * F);
        const __half* wi = w + (size_t)co * Cin * 3;
        const float* xi = in + (size_t)f * Cin * Lin;
        const int l0 = (int)(lo * stride) - 1;
        const bool in0 = l0 >= 0, in2 = l0 + 2 < (int)Lin;
        float acc0 = 0.f, acc1 = 0.f;
        uint32_t ci = 0;
        if ((Cin & 3u) == 0 && (((uintptr_t)wi) & 7u) == 0) {
            for (; ci < Cin; ci += 4) {
                const uint2* wq = reinterpret_cast<const uint2*>(wi + ci * 3);  // 12 halves = 3 x 8 bytes
                const uint2 q0 = __ldg(wq), q1 = __ldg(wq + 1), q2 = __ldg(wq + 2);
                float xv[12];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* x = xi + (ci + c) * Lin + l0;
                    xv[3 * c] = in0 ? x[0] : 0.f;
                    xv[3 * c + 1] = x[1];
                    xv[3 * c + 2] = in2 ? x[2] : 0.f;
                }
                const __half2 h0 = *reinterpret_cast<const __half2*>(&q0.x), h1 = *reinterpret_cast<const __half2*>(&q0.y),
                              h2 = *reinterpret_cast<const __half2*>(&q1.x), h3 = *reinterpret_cast<const __half2*>(&q1.y),
                              h4 = *reinterpret_cast<const __half2*>(&q2.x), h5 = *reinterpret_cast<const __half2*>(&q2.y);
                acc0 = __fmaf_rn(__low2float(h0), xv[0], acc0);  acc1 = __fmaf_rn(__high2float(h0), xv[1], acc1);
                acc0 = __fmaf_rn(__low2float(h1), xv[2], acc0);  acc1 = __fmaf_rn(__high2float(h1), xv[3], acc1);
                acc0 = __fmaf_rn(__low2float(h2), xv[4], acc0);  acc1 = __fmaf_rn(__high2float(h2), xv[5], acc1);
                acc0 = __fmaf_rn(__low2float(h3), xv[6], acc0);  acc1 = __fmaf_rn(__high2float(h3), xv[7], acc1);
                acc0 = __fmaf_rn(__low2float(h4), xv[8], acc0);  acc1 = __fmaf_rn(__high2float(h4), xv[9], acc1);
                acc0 = __fmaf_rn(__low2float(h5), xv[10], acc0); acc1 = __fmaf_rn(__high2float(h5), xv[11], acc1);
            }
        }
        for (; ci < Cin; ++ci) {  // generic tail (tiny attention layers)
            const float* x = xi + ci * Lin + l0;
            if (in0) acc0 = __fmaf_rn(__half2float(__ldg(wi + ci * 3)), x[0], acc0);
            acc1 = __fmaf_rn(__half2float(__ldg(wi + ci * 3 + 1)), x[1], acc1);
            if (in2) acc0 = __fmaf_rn(__half2float(__ldg(wi + ci * 3 + 2)), x[2], acc0);
        }
        const float y = h16(h16(acc0 + acc1) + __half2float(__ldg(b + co)));
        out[((size_t)f * Cout + co) * Lout + lo] = h16(leaky(y));
    }
    __syncthreads();
}

// Register-tiled variant for the AudioNet convolutions (Cout % 8 == 0, Cin % 4 == 0): one thread produces EIGHT output
// channels of one (frame, position), so every input tap read from shared memory feeds 8 FMAs.  The one-channel-per-thread
// version above is shared-memory bound: 2048 outputs x 132 taps with 4-way bank conflicts = 34 k wavefronts for conv1
// alone (measured 40 k cycles).
__device__ void conv_layer_t8(const float* __restrict__ in, float* __restrict__ out, const __half* __restrict__ w,
                              const __half* __restrict__ b, uint32_t F, uint32_t Cin, uint32_t Cout, uint32_t Lin, uint32_t stride) {
    const uint32_t Lout = (Lin + 2 - 3) / stride + 1;
    const uint32_t FL = F * Lout, total = (Cout / 8) * FL;
    for (uint32_t o = threadIdx.x; o < total; o += blockDim.x) {
        const uint32_t lo = o % Lout, f = (o / Lout) % F, cg = o / FL;
        const float* xi = in + (size_t)f * Cin * Lin;
        const int l0 = (int)(lo * stride) - 1;
        const bool in0 = l0 >= 0, in2 = l0 + 2 < (int)Lin;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        for (uint32_t ci = 0; ci < Cin; ci += 4) {
            float xv[12];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float* x = xi + (ci + c) * Lin + l0;
                xv[3 * c] = in0 ? x[0] : 0.f;
                xv[3 * c + 1] = x[1];
                xv[3 * c + 2] = in2 ? x[2] : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint2* wq = reinterpret_cast<const uint2*>(w + ((size_t)(cg * 8 + c) * Cin + ci) * 3);  // 12 halves
                const uint2 q0 = __ldg(wq), q1 = __ldg(wq + 1), q2 = __ldg(wq + 2);
                const __half2 h0 = *reinterpret_cast<const __half2*>(&q0.x), h1 = *reinterpret_cast<const __half2*>(&q0.y),
                              h2 = *reinterpret_cast<const __half2*>(&q1.x), h3 = *reinterpret_cast<const __half2*>(&q1.y),
                              h4 = *reinterpret_cast<const __half2*>(&q2.x), h5 = *reinterpret_cast<const __half2*>(&q2.y);
                float a = acc[c];
                a = __fmaf_rn(__low2float(h0), xv[0], a);  a = __fmaf_rn(__high2float(h0), xv[1], a);
                a = __fmaf_rn(__low2float(h1), xv[2], a);  a = __fmaf_rn(__high2float(h1), xv[3], a);
                a = __fmaf_rn(__low2float(h2), xv[4], a);  a = __fmaf_rn(__high2float(h2), xv[5], a);
                a = __fmaf_rn(__low2float(h3), xv[6], a);  a = __fmaf_rn(__high2float(h3), xv[7], a);
                a = __fmaf_rn(__low2float(h4), xv[8], a);  a = __fmaf_rn(__high2float(h4), xv[9], a);
                a = __fmaf_rn(__low2float(h5), xv[10], a); a = __fmaf_rn(__high2float(h5), xv[11], a);
                acc[c] = a;
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint32_t co = cg * 8 + c;
            const float y = h16(h16(acc[c]) + __half2float(__ldg(b + co)));
            out[((size_t)f * Cout + co) * Lout + lo] = h16(leaky(y));
        }
    }
    __syncthreads();
}

// Linear(+bias) on R rows: in [R][K] -> out [R][N]; lanes run over the rows of one output feature (uniform weight loads);
// weights are fetched 8 at a time (one 16-byte load) when the row is aligned
__device__ void linear_layer(const float* __restrict__ in, float* __restrict__ out, const __half* __restrict__ w, const __half* __restrict__ b,
                             uint32_t R, uint32_t K, uint32_t N, bool act) {
    for (uint32_t o = threadIdx.x; o < R * N; o += blockDim.x) {
        const uint32_t r = o % R, n = o / R;
        const __half* wr = w + (size_t)n * K;
        const float* x = in + r * K;
        float acc0 = 0.f, acc1 = 0.f;
        uint32_t k = 0;
        if ((K & 7u) == 0 && (((uintptr_t)wr) & 15u) == 0) {
            for (; k < K; k += 8) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(wr + k));
                float xv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) xv[i] = x[k + i];
                const __half2 h0 = *reinterpret_cast<const __half2*>(&q.x), h1 = *reinterpret_cast<const __half2*>(&q.y),
                              h2 = *reinterpret_cast<const __half2*>(&q.z), h3 = *reinterpret_cast<const __half2*>(&q.w);
                acc0 = __fmaf_rn(__low2float(h0), xv[0], acc0); acc1 = __fmaf_rn(__high2float(h0), xv[1], acc1);
                acc0 = __fmaf_rn(__low2float(h1), xv[2], acc0); acc1 = __fmaf_rn(__high2float(h1), xv[3], acc1);
                acc0 = __fmaf_rn(__low2float(h2), xv[4], acc0); acc1 = __fmaf_rn(__high2float(h2), xv[5], acc1);
                acc0 = __fmaf_rn(__low2float(h3), xv[6], acc0); acc1 = __fmaf_rn(__high2float(h3), xv[7], acc1);
            }
        }
        for (; k < K; ++k) acc0 = __fmaf_rn(__half2float(__ldg(wr + k)), x[k], acc0);
        float y = h16(acc0 + acc1 + __half2float(__ldg(b + n)));
        if (act) y = h16(leaky(y));
        out[r * N + n] = y;
    }
    __syncthreads();
}

// bias[n] = sum_k W16[n, col0 + k] * v[k]   (fp32 accumulate; v already rounded to fp16 by the caller).  One warp per output row: lanes stride over k
// (coalesced row reads), then a shuffle reduction.
__device__ void hoist(const __half* __restrict__ W, uint32_t ld, uint32_t col0, const float* v, uint32_t K, uint32_t N, float* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t n = warp; n < N; n += nwarps) {
        float acc = 0.f;
        for (uint32_t k = lane; k < K; k += 32) acc = __fmaf_rn(__half2float(__ldg(W + (size_t)n * ld + col0 + k)), v[k], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) out[n] = acc;
    }
}

// The kernel is one CTA walking ~130 KB of fp16 weights layer by layer: without help every layer pays exposed L2 round
// trips on its first touch of each line.  All lines are requested up front (each thread a few prefetches) so that the
// layers find them in L1.
__device__ __forceinline__ void prefetch_l1(const __half* w, uint32_t n_halves) {
    if (!w) return;
    const char* base = reinterpret_cast<const char*>(w);
    const uint32_t bytes = n_halves * 2;
    for (uint32_t off = threadIdx.x * 128; off < bytes; off += blockDim.x * 128)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(base + off));
}

__global__ void __launch_bounds__(512)
audio_frame_kernel(AudioParams p) {
    __shared__ float bufA[8 * 44 * 16];
    __shared__ float bufB[8 * 32 * 8];
    __shared__ float s_enc[64];
    __shared__ float s_vec[64];
    const uint32_t F = p.F;
    long long* prof = g_audio_prof_dev;
    RN_STAMP(0)

    if (p.auds) {
        prefetch_l1(p.conv_w[0], 32 * p.Cin * 3); prefetch_l1(p.conv_w[1], 32 * 32 * 3);
        prefetch_l1(p.conv_w[2], 64 * 32 * 3);    prefetch_l1(p.conv_w[3], 64 * 64 * 3);
        prefetch_l1(p.fc_w[0], 64 * 64);          prefetch_l1(p.fc_w[1], 64 * 64);
        if (p.att > 0) prefetch_l1(p.att_w[0], 16 * 64 * 3);
        prefetch_l1(p.w_amb1, 64 * 96);
    }
    prefetch_l1(p.w_def1, 64 * 104); prefetch_l1(p.w_tor1, 32 * 136);
    if (p.auds) {
        for (uint32_t i = threadIdx.x; i < F * p.Cin * 16; i += blockDim.x) bufA[i] = h16(__ldg(p.auds + i));
        __syncthreads();
        RN_STAMP(1)
        if ((p.Cin & 3u) == 0) conv_layer_t8(bufA, bufB, p.conv_w[0], p.conv_b[0], F, p.Cin, 32, 16, 2); else conv_layer(bufA, bufB, p.conv_w[0], p.conv_b[0], F, p.Cin, 32, 16, 2);   // -> [F,32,8]
        RN_STAMP(2)
        conv_layer_t8(bufB, bufA, p.conv_w[1], p.conv_b[1], F, 32, 32, 8, 2);       // -> [F,32,4]
        conv_layer_t8(bufA, bufB, p.conv_w[2], p.conv_b[2], F, 32, 64, 4, 2);       // -> [F,64,2]
        conv_layer_t8(bufB, bufA, p.conv_w[3], p.conv_b[3], F, 64, 64, 2, 2);       // -> [F,64,1]
        RN_STAMP(3)
        linear_layer(bufA, bufB, p.fc_w[0], p.fc_b[0], F, 64, 64, true);
        linear_layer(bufB, bufA, p.fc_w[1], p.fc_b[1], F, 64, 64, false);        // x = bufA [F][64]  (fp16 values)
        RN_STAMP(4)
        if (p.att > 0) {
            // y = x^T as [1][64][F]; conv stack 64 -> 16 -> 8 -> 4 -> 2 -> 1 over the F frames
            float* xt = bufB;               // [64][F]
            float* tmp = bufB + 64 * 8;     // scratch
            for (uint32_t i = threadIdx.x; i < 64 * F; i += blockDim.x) xt[i] = bufA[(i % F) * 64 + i / F];
            __syncthreads();
            float* y0 = tmp;                // [16][F]
            float* y1 = tmp + 16 * 8;       // [8][F]
            conv_layer(xt, y0, p.att_w[0], p.att_b[0], 1, 64, 16, F, 1);
            conv_layer(y0, y1, p.att_w[1], p.att_b[1], 1, 16, 8, F, 1);
            conv_layer(y1, y0, p.att_w[2], p.att_b[2], 1, 8, 4, F, 1);
            conv_layer(y0, y1, p.att_w[3], p.att_b[3], 1, 4, 2, F, 1);
            conv_layer(y1, y0, p.att_w[4], p.att_b[4], 1, 2, 1, F, 1);          // y0[0..F)
            RN_STAMP(5)
            linear_layer(y0, y1, p.att_fc_w, p.att_fc_b, 1, F, F, false);       // y1[0..F)  fp16 logits
            if (threadIdx.x == 0) {  // softmax in fp32
                float m = -INFINITY, s = 0.f;
                for (uint32_t i = 0; i < F; ++i) m = fmaxf(m, y1[i]);
                for (uint32_t i = 0; i < F; ++i) { y0[i] = expf(y1[i] - m); s += y0[i]; }
                for (uint32_t i = 0; i < F; ++i) y0[i] = y0[i] / s;
            }
            __syncthreads();
            if (threadIdx.x < 64) {  // enc_a[c] = sum_f softmax[f] * x[f][c]   (fp32)
                float acc = 0.f;
                for (uint32_t f = 0; f < F; ++f) acc += y0[f] * bufA[f * 64 + threadIdx.x];
                s_enc[threadIdx.x] = acc;
            }
        } else if (threadIdx.x < 64) {
            s_enc[threadIdx.x] = bufA[threadIdx.x];
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            float e = s_enc[threadIdx.x];
            if (p.smooth) {
                if (p.enc_a_state[64] != 0.f) e = __fadd_rn(__fmul_rn(p.lambda, p.enc_a_state[threadIdx.x]), __fmul_rn(1.0f - p.lambda, e));
                p.enc_a_state[threadIdx.x] = e;
            }
            s_enc[threadIdx.x] = e;
        }
        __syncthreads();
        if (p.smooth && threadIdx.x == 0) p.enc_a_state[64] = 1.0f;  // every thread has read the flag before the barrier above
    }

    RN_STAMP(6)
    // ---- hoisted terms of the head
    if (p.auds) {
        if (threadIdx.x < 64) s_vec[threadIdx.x] = h16(s_enc[threadIdx.x]);
        __syncthreads();
        hoist(p.w_amb1, 96, 32, s_vec, 64, 64, p.head_consts);
        __syncthreads();
    }
    else if (threadIdx.x < 64) p.head_consts[threadIdx.x] = 0.f;
    if (threadIdx.x < 64) {
        const float e = p.eye ? h16(__ldg(p.eye)) : 0.f;
        p.head_consts[64 + threadIdx.x] = p.eye ? __half2float(__ldg(p.w_sig1 + (size_t)threadIdx.x * 65 + 64)) * e : 0.f;
    }
    if (p.ind_code) {
        if (threadIdx.x < 4) s_vec[threadIdx.x] = h16(__ldg(p.ind_code + threadIdx.x));
        __syncthreads();
        hoist(p.w_col1, 84, 80, s_vec, 4, 64, p.head_consts + 128);
    } else if (threadIdx.x < 64) {
        p.head_consts[128 + threadIdx.x] = 0.f;
    }
    __syncthreads();

    RN_STAMP(7)
    // ---- hoisted terms of the torso: [freq(pose6) (54) | individual code (8)]
    if (p.w_def1) {
        if (threadIdx.x < 54) {
            const uint32_t c = threadIdx.x;
            float v;
            if (c < 6) v = __ldg(p.pose6 + c);
            else {
                const uint32_t col = c / 6 - 1, d = c % 6, fr = col / 2;
                v = __sinf(scalbnf(__ldg(p.pose6 + d), (int)fr) + (float)(col % 2) * 1.5707963705062866f);
            }
            s_vec[c] = h16(v);
        } else if (threadIdx.x < 62) {
            s_vec[threadIdx.x] = p.ind_torso ? h16(__ldg(p.ind_torso + (threadIdx.x - 54))) : 0.f;
        }
        __syncthreads();
        const uint32_t K = p.ind_torso ? 62u : 54u;
        hoist(p.w_def1, 42 + K, 42, s_vec, K, 64, p.torso_consts);
        hoist(p.w_tor1, 32 + 42 + K, 74, s_vec, K, 32, p.torso_consts + 64);
    }
    __syncthreads();
    RN_STAMP(8)
}

}  // namespace

void set_audio_prof(void* p) { cudaMemcpyToSymbol(g_audio_prof_dev, &p, sizeof(p)); }

int launch_audio_frame(const AudioParams& p, cudaStream_t st) {
    audio_frame_kernel<<<1, 512, 0, st>>>(p);
    return finish_launch("audio_frame");
}

}  // namespace rn
