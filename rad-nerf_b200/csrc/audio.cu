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

__device__ __forceinline__ float h16(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.02f * x; }

// Conv1d(k=3, pad=1) + bias + LeakyReLU(0.02) for F frames; in [F][Cin][Lin] -> out [F][Cout][Lout].
// Work item o = (co, f, lo) with (f, lo) fastest: the lanes of a warp share the output channel, so every weight load is a
// warp-uniform broadcast (one L1 transaction) instead of 32 scattered rows, and the loads of consecutive taps pipeline.
__device__ void conv_layer(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w, const float* __restrict__ b,
                           uint32_t F, uint32_t Cin, uint32_t Cout, uint32_t Lin, uint32_t stride) {
    const uint32_t Lout = (Lin + 2 - 3) / stride + 1;
    const uint32_t total = F * Cout * Lout;
    for (uint32_t o = threadIdx.x; o < total; o += blockDim.x) {
        const uint32_t lo = o % Lout, f = (o / Lout) % F, co = o / (Lout * F);
        const float* wi = w + (size_t)co * Cin * 3;
        const float* xi = in + (size_t)f * Cin * Lin;
        const int l0 = (int)(lo * stride) - 1;
        const bool in0 = l0 >= 0, in2 = l0 + 2 < (int)Lin;
        float acc = 0.f;
#pragma unroll 4
        for (uint32_t ci = 0; ci < Cin; ++ci) {
            const float w0 = h16(__ldg(wi + ci * 3)), w1 = h16(__ldg(wi + ci * 3 + 1)), w2 = h16(__ldg(wi + ci * 3 + 2));
            const float* x = xi + ci * Lin + l0;
            if (in0) acc = __fmaf_rn(w0, x[0], acc);
            acc = __fmaf_rn(w1, x[1], acc);
            if (in2) acc = __fmaf_rn(w2, x[2], acc);
        }
        const float y = h16(h16(acc) + h16(__ldg(b + co)));
        out[((size_t)f * Cout + co) * Lout + lo] = h16(leaky(y));
    }
    __syncthreads();
}

// Linear(+bias) on R rows: in [R][K] -> out [R][N]; lanes run over the rows of one output feature (uniform weight loads)
__device__ void linear_layer(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w, const float* __restrict__ b,
                             uint32_t R, uint32_t K, uint32_t N, bool act) {
    for (uint32_t o = threadIdx.x; o < R * N; o += blockDim.x) {
        const uint32_t r = o % R, n = o / R;
        const float* wr = w + (size_t)n * K;
        float acc = 0.f;
#pragma unroll 8
        for (uint32_t k = 0; k < K; ++k) acc = __fmaf_rn(h16(__ldg(wr + k)), in[r * K + k], acc);
        float y = h16(acc + h16(__ldg(b + n)));
        if (act) y = h16(leaky(y));
        out[r * N + n] = y;
    }
    __syncthreads();
}

// bias[n] = sum_k fp16(W[n, col0 + k]) * fp16(v[k])   (fp32 accumulate).  One warp per output row: lanes stride over k
// (coalesced row reads), then a shuffle reduction.
__device__ void hoist(const float* __restrict__ W, uint32_t ld, uint32_t col0, const float* v, uint32_t K, uint32_t N, float* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t n = warp; n < N; n += nwarps) {
        float acc = 0.f;
        for (uint32_t k = lane; k < K; k += 32) acc = __fmaf_rn(h16(__ldg(W + (size_t)n * ld + col0 + k)), h16(v[k]), acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) out[n] = acc;
    }
}

__global__ void __launch_bounds__(512)
audio_frame_kernel(AudioParams p) {
    __shared__ float bufA[8 * 44 * 16];
    __shared__ float bufB[8 * 32 * 8];
    __shared__ float s_enc[64];
    __shared__ float s_vec[64];
    const uint32_t F = p.F;

    if (p.auds) {
        for (uint32_t i = threadIdx.x; i < F * p.Cin * 16; i += blockDim.x) bufA[i] = h16(__ldg(p.auds + i));
        __syncthreads();
        conv_layer(bufA, bufB, p.conv_w[0], p.conv_b[0], F, p.Cin, 32, 16, 2);   // -> [F,32,8]
        conv_layer(bufB, bufA, p.conv_w[1], p.conv_b[1], F, 32, 32, 8, 2);       // -> [F,32,4]
        conv_layer(bufA, bufB, p.conv_w[2], p.conv_b[2], F, 32, 64, 4, 2);       // -> [F,64,2]
        conv_layer(bufB, bufA, p.conv_w[3], p.conv_b[3], F, 64, 64, 2, 2);       // -> [F,64,1]
        linear_layer(bufA, bufB, p.fc_w[0], p.fc_b[0], F, 64, 64, true);
        linear_layer(bufB, bufA, p.fc_w[1], p.fc_b[1], F, 64, 64, false);        // x = bufA [F][64]  (fp16 values)
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

    // ---- hoisted terms of the head
    if (p.auds) hoist(p.w_amb1, 96, 32, s_enc, 64, 64, p.head_consts);
    else if (threadIdx.x < 64) p.head_consts[threadIdx.x] = 0.f;
    if (threadIdx.x < 64) {
        const float e = p.eye ? h16(__ldg(p.eye)) : 0.f;
        p.head_consts[64 + threadIdx.x] = p.eye ? h16(__ldg(p.w_sig1 + (size_t)threadIdx.x * 65 + 64)) * e : 0.f;
    }
    if (p.ind_code) {
        if (threadIdx.x < 4) s_vec[threadIdx.x] = __ldg(p.ind_code + threadIdx.x);
        __syncthreads();
        hoist(p.w_col1, 84, 80, s_vec, 4, 64, p.head_consts + 128);
    } else if (threadIdx.x < 64) {
        p.head_consts[128 + threadIdx.x] = 0.f;
    }
    __syncthreads();

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
            s_vec[c] = v;
        } else if (threadIdx.x < 62) {
            s_vec[threadIdx.x] = p.ind_torso ? __ldg(p.ind_torso + (threadIdx.x - 54)) : 0.f;
        }
        __syncthreads();
        const uint32_t K = p.ind_torso ? 62u : 54u;
        hoist(p.w_def1, 42 + K, 42, s_vec, K, 64, p.torso_consts);
        hoist(p.w_tor1, 32 + 42 + K, 74, s_vec, K, 32, p.torso_consts + 64);
    }
}

}  // namespace

int launch_audio_frame(const AudioParams& p, cudaStream_t st) {
    audio_frame_kernel<<<1, 512, 0, st>>>(p);
    return finish_launch("audio_frame");
}

}  // namespace rn
