// audio.cu -- per-frame conditioning in ONE single-CTA kernel: AudioNet + AudioAttNet (nerf/network.py:10-67,170-185),
// the lip-smoothing EMA (nerf/renderer.py:190-194), and the per-frame-constant halves of the first MLP layers folded into
// fp32 bias vectors for the fused head / torso kernels ("hoisting": the audio code, eye value, individual codes and the
// encoded head pose are identical for every sample of a frame, so W[:, const part] @ const is computed once here).
//
// The reference runs ~25 tiny cuDNN / cuBLAS launches for this (~0.66 MMAC); here it is one launch.  Numerics follow
// fp16 autocast: weights and activations are rounded to fp16, sums are fp32, conv outputs are rounded before AND after
// the (fp16) bias add as PyTorch's cuDNN path does, Linear adds the bias before its single rounding, softmax and the
// attention-weighted sum are fp32.
#include <algorithm>
#include "frame.cuh"
#include "umma.cuh"

namespace rn {


namespace {

__device__ long long* g_audio_prof_dev = nullptr;
#define RN_STAMP(i) if (prof && threadIdx.x == 0) prof[i] = clock64();

__device__ __forceinline__ float h16(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.02f * x; }

// ---- layer engine --------------------------------------------------------------------------------------------------------------
// Every layer is a small GEMM  out[m, n] = sum_k W[m, k] * X[k, n]  (m = output channel / feature, n = (frame, position) or row),
// at most 64 x 64 x 192.  One warp owns a 16 x 8 output tile and runs mma.sync.m16n8k16 (fp16 operands, fp32 accumulate).
// (tcgen05 needs M >= 64 and a TMEM round trip per layer -- not worth it at this size.)
//
//   * Weights stay in the reference's layout ([Cout][Cin][3] / [N][K]) and are staged once per launch into a shared-memory
//     arena by TMA bulk copies, one mbarrier per matrix.
//   * Activations live in shared memory POSITION-MAJOR: [frame][position (+ one zero border each side for k=3 convs)]
//     [channel (zero-padded to a multiple of 16)].  With K ordered (tap, channel) the B fragment of a convolution is two
//     aligned 32-bit loads from that buffer -- the im2col is implicit and branch-free; each layer's epilogue writes straight
//     into the layout its consumer wants, after the CTA has zeroed that region (borders and channel padding).
//   * The kernel is a single CTA that runs once per frame with a cold instruction cache and 4 warps per scheduler, so
//     instruction count is what it pays for.  Earlier versions (scalar FMA loops; one inlined specialised mma loop per
//     layer, 6000 instructions) took 45-90 us.  Hence ONE generic routine driven by a host-built layer table.
enum : uint16_t { ST_CONV = 0, ST_LINEAR = 1, ST_LINEAR_ACT = 2, ST_VEC = 3 };
enum : uint16_t { BUF_A = 0, BUF_B = 1, BUF_VEC = 2, BUF_HEAD = 3, BUF_TORSO = 4 };

struct AudioLayer {
    uint32_t w_off;                                      // byte offset of W in the arena
    uint16_t slot, ld, col0, M, N, Cin, taps, stride, Lout;   // W[m][col0 + ci*taps + tap]; n = (f, lo) with lo < Lout
    uint16_t in_buf, in_off, in_fs, in_cs;               // X element (f, position index q, ci) at in[in_off + f*in_fs + q*in_cs + ci], q = lo*stride + tap
    uint16_t out_buf, out_off, out_fs, out_cs, out_pad;  // result (m, f, lo) at out[out_off + f*out_fs + (lo + out_pad)*out_cs + m]
    uint16_t zero_halves;                                // halves of the output region cleared before the layer (multiple of 2)
    uint16_t mode, bias_row, stamp;
};
constexpr int N_SLOTS = 15, MAX_LAYERS = 16;
struct AudioSlot { const __half* src; uint32_t bytes, off; uint16_t rows, ld_src, ld_dst, tma; };   // tma = 0: CTA copies rows x ld_src -> row stride ld_dst
struct AudioProgram {
    AudioSlot slots[N_SLOTS];
    AudioLayer layers[MAX_LAYERS];
    const __half* bias_src[12];
    uint16_t bias_n[12];
    uint16_t n_net, n_head, n_torso, cp0;   // layers [0,n_net) audio+attention, then n_head head hoists, then n_torso torso hoists; cp0 = padded Cin
};

constexpr uint32_t ACT_A_HALVES = 8 * 18 * 48, ACT_B_HALVES = 8 * 10 * 32;
constexpr uint32_t ARENA_MAX_BYTES = 128 * 1024;
constexpr uint32_t AUDIO_SMEM_BYTES = ARENA_MAX_BYTES + (ACT_A_HALVES + ACT_B_HALVES) * 2;

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct AudioCtx {
    uint8_t* smem; __half* hA; __half* hB; __half* vec; const float* bias; uint64_t* mbar; const AudioLayer* layers; const uint16_t* slot_tma;
    float* head_consts; float* torso_consts; long long* prof;
};

// weight (row, channel ci) of the current tap; channels past Cin (K padding) read as zero, branch-free
__device__ __forceinline__ uint32_t ld_a(const __half* w, uint32_t ci, uint32_t Cin, uint32_t taps) {
    const uint32_t v = __half_as_ushort(w[min(ci, Cin - 1) * taps]);
    return ci < Cin ? v : 0u;
}

__device__ __noinline__ void run_layers(const AudioCtx cx, uint32_t first, uint32_t count) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint32_t g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (uint32_t li = first; li < first + count; ++li) {
        const AudioLayer L = cx.layers[li];
        const uint32_t M = L.M, N = L.N, Cin = L.Cin, taps = L.taps, Lout = L.Lout, mode = L.mode;
        const __half* in = (L.in_buf == BUF_A ? cx.hA : L.in_buf == BUF_B ? cx.hB : cx.vec) + L.in_off;
        __half* out_h = (L.out_buf == BUF_A ? cx.hA : cx.hB) + L.out_off;
        float* out_f = (L.out_buf == BUF_HEAD ? cx.head_consts : cx.torso_consts) + L.out_off;
        const float* bias = cx.bias + L.bias_row * 64;
        const __half* W = reinterpret_cast<const __half*>(cx.smem + L.w_off) + L.col0;
        if (L.zero_halves) {
            uint32_t* z = reinterpret_cast<uint32_t*>(out_h);
            for (uint32_t i = tid; i < L.zero_halves / 2u; i += blockDim.x) z[i] = 0u;
            __syncthreads();
        }
        if (cx.slot_tma[L.slot]) umma::mbar_wait(&cx.mbar[L.slot], 0);
        const uint32_t nt = (N + 7) / 8, tiles = ((M + 15) / 16) * nt;
        for (uint32_t tile = warp; tile < tiles; tile += nwarps) {
            const uint32_t m0 = (tile / nt) * 16, n0 = (tile % nt) * 8;
            const uint32_t r0 = m0 + g, r1 = r0 + 8;
            // rows / columns past the edge are clamped: they compute garbage that is never stored
            const __half* w0 = W + min(r0, M - 1) * (uint32_t)L.ld;
            const __half* w1 = W + min(r1, M - 1) * (uint32_t)L.ld;
            const uint32_t n = min(n0 + g, N - 1), f = n / Lout, lo = n - f * Lout;
            const __half* xq = in + f * L.in_fs + lo * L.stride * L.in_cs + 2 * t;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (uint32_t tap = 0; tap < taps; ++tap, xq += L.in_cs) {
#pragma unroll 1
                for (uint32_t c0 = 0; c0 < Cin; c0 += 16) {
                    const uint32_t ca = c0 + 2 * t, cb = ca + 8;
                    const uint32_t a0 = ld_a(w0 + tap, ca, Cin, taps) | (ld_a(w0 + tap, ca + 1, Cin, taps) << 16);
                    const uint32_t a1 = ld_a(w1 + tap, ca, Cin, taps) | (ld_a(w1 + tap, ca + 1, Cin, taps) << 16);
                    const uint32_t a2 = ld_a(w0 + tap, cb, Cin, taps) | (ld_a(w0 + tap, cb + 1, Cin, taps) << 16);
                    const uint32_t a3 = ld_a(w1 + tap, cb, Cin, taps) | (ld_a(w1 + tap, cb + 1, Cin, taps) << 16);
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(xq + c0);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(xq + c0 + 8);
                    mma16816(c, a0, a1, a2, a3, b0, b1);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t m = (i & 2) ? r1 : r0, nn = n0 + 2 * t + (i & 1);
                if (m >= M || nn >= N) continue;
                const float acc = c[i];
                if (mode == ST_VEC) { out_f[m] = acc; continue; }
                const float bv = bias[m];
                float y;
                if (mode == ST_CONV) y = leaky(h16(h16(acc) + bv));   // round, add the fp16 bias, round, LeakyReLU, round
                else {                                                // Linear: bias added before the single rounding
                    y = h16(acc + bv);
                    if (mode == ST_LINEAR_ACT) y = leaky(y);
                }
                const uint32_t ff = nn / Lout, l2 = nn - ff * Lout;
                out_h[ff * L.out_fs + (l2 + L.out_pad) * L.out_cs + m] = __float2half_rn(y);
            }
        }
        __syncthreads();
        if (L.stamp && cx.prof && tid == 0) cx.prof[L.stamp] = clock64();
    }
}

__global__ void __launch_bounds__(512)
audio_frame_kernel(const __grid_constant__ AudioParams p, const __grid_constant__ AudioProgram pg) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar[N_SLOTS];
    __shared__ __align__(16) AudioLayer s_layers[MAX_LAYERS];
    __shared__ uint16_t s_tma[N_SLOTS + 1];
    __shared__ float s_bias[12][64];   // conv 0-3, fc 4-5, attention convs 6-10, attention fc 11
    __shared__ float s_enc[64];
    __shared__ __align__(16) __half s_vec[64];
    __shared__ float s_soft[8];
    __half* hA = reinterpret_cast<__half*>(smem + ARENA_MAX_BYTES);
    __half* hB = hA + ACT_A_HALVES;
    const uint32_t F = p.F, tid = threadIdx.x;
    long long* prof = g_audio_prof_dev;
    RN_STAMP(0)

    // ---- stage every weight matrix of the frame in shared memory: one TMA bulk copy + one mbarrier per matrix, each issued
    //      by its own thread, so a layer only waits for its own weights
    // phase (AudioParams::reserved): 0 = the whole conditioning; 1 = FEATURES only -- AudioNet + attention of this frame, the raw
    // (unsmoothed) code parked in head_consts[0..63]; 2 = SMOOTH + HOIST only -- picks the raw code up there.  Frames in flight on
    // several lanes are coupled only through the lip-smoothing EMA: with the split, the ~35 us of network evaluation run concurrently
    // on the lanes' own streams and only the ~5 us tail is serialised in frame order (one 40 us kernel per frame in frame order
    // capped a GPU at ~21 k frames/s and put 20 x 46 us in front of a 20-frame window on 8 GPUs).
    const uint32_t phase = p.reserved;
    auto staged = [&](uint32_t slot) { return phase == 0 || ((phase == 1) == (slot < 12)); };
    if (tid < N_SLOTS) {
        const AudioSlot& sl = pg.slots[tid];
        umma::mbar_init(&mbar[tid], 1);
        umma::fence_mbar_init();
        s_tma[tid] = sl.tma;
        if (sl.tma && staged(tid)) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(&mbar[tid])), "r"(sl.bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(smem + sl.off)),
                         "l"(sl.src), "r"(sl.bytes), "r"(umma::smem_u32(&mbar[tid]))
                         : "memory");
        }
    }
    {   // layer table -> shared memory (the interpreter indexes it dynamically)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(pg.layers);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_layers);
        for (uint32_t i = tid; i < sizeof(pg.layers) / 4; i += blockDim.x) dst[i] = src[i];
    }
    if (p.auds && phase != 2) {
        const uint32_t cp0 = pg.cp0;
        uint32_t* z = reinterpret_cast<uint32_t*>(hA);
        for (uint32_t i = tid; i < F * 18 * cp0 / 2; i += blockDim.x) z[i] = 0u;
        const uint32_t row = tid >> 5;   // 16 warps >= 12 bias rows
        if (row < 12)
            for (uint32_t c = tid & 31; c < pg.bias_n[row]; c += 32) s_bias[row][c] = __half2float(__ldg(pg.bias_src[row] + c));
    }
#pragma unroll 1
    for (int i = 0; i < N_SLOTS; ++i) {   // odd-sized / unaligned matrices: plain copy, re-strided to an even row pitch
        const AudioSlot& sl = pg.slots[i];
        if (sl.bytes == 0 || sl.tma || !staged((uint32_t)i)) continue;
        __half* dst = reinterpret_cast<__half*>(smem + sl.off);
        for (uint32_t j = tid; j < (uint32_t)sl.rows * sl.ld_src; j += blockDim.x) {
            const uint32_t r = j / sl.ld_src, cc = j - r * sl.ld_src;
            dst[r * sl.ld_dst + cc] = __ldg(sl.src + j);
        }
    }
    __syncthreads();
    if (p.auds && phase != 2) {   // [F][Cin][16] fp32 -> [F][1 + 16 + 1][cp0] fp16
        const uint32_t Cin = p.Cin, cp0 = pg.cp0;
        for (uint32_t i = tid; i < F * Cin * 16; i += blockDim.x) {
            const uint32_t l = i & 15, fc = i >> 4, f = fc / Cin, ci = fc - f * Cin;
            hA[(f * 18 + l + 1) * cp0 + ci] = __float2half_rn(__ldg(p.auds + i));
        }
    }
    __syncthreads();   // barriers initialised, table / inputs / biases / fallback copies visible
    RN_STAMP(1)

    AudioCtx cx;
    cx.smem = smem; cx.hA = hA; cx.hB = hB; cx.vec = s_vec; cx.bias = &s_bias[0][0]; cx.mbar = mbar; cx.layers = s_layers; cx.slot_tma = s_tma;
    cx.head_consts = p.head_consts; cx.torso_consts = p.torso_consts; cx.prof = prof;

    if (p.auds && phase == 2) {
        if (tid < 64) s_enc[tid] = p.head_consts[tid];      // the raw code of this frame, left there by the phase-1 launch
        __syncthreads();
    }
    if (p.auds && phase != 2) {
        run_layers(cx, 0, pg.n_net);   // AudioNet convs + fcs -> x, attention convs + fc -> logits hB[256 .. 256+F)
        const __half* x = hA + (p.att > 0 ? 64 : 0);   // rows 1..F of the attention input layout
        if (p.att > 0) {
            const __half* logits = hB + 256;
            if (tid == 0) {  // softmax in fp32
                float m = -INFINITY, s = 0.f;
                for (uint32_t i = 0; i < F; ++i) m = fmaxf(m, __half2float(logits[i]));
                for (uint32_t i = 0; i < F; ++i) { const float e = expf(__half2float(logits[i]) - m); s_soft[i] = e; s += e; }
                for (uint32_t i = 0; i < F; ++i) s_soft[i] = s_soft[i] / s;
            }
            __syncthreads();
            if (tid < 64) {  // enc_a[c] = sum_f softmax[f] * x[f][c]   (fp32)
                float acc = 0.f;
                for (uint32_t f = 0; f < F; ++f) acc += s_soft[f] * __half2float(x[f * 64 + tid]);
                s_enc[tid] = acc;
            }
        } else if (tid < 64) {
            s_enc[tid] = __half2float(x[tid]);
        }
        __syncthreads();
        if (phase == 1) {
            if (tid < 64) p.head_consts[tid] = s_enc[tid];
            return;
        }
    }
    if (p.auds) {
        if (tid < 64) {
            float e = s_enc[tid];
            if (p.smooth) {
                if (p.enc_a_state[64] != 0.f) e = __fadd_rn(__fmul_rn(p.lambda, p.enc_a_state[tid]), __fmul_rn(1.0f - p.lambda, e));
                p.enc_a_state[tid] = e;
            }
            s_vec[tid] = __float2half_rn(e);
        }
        __syncthreads();
        if (p.smooth && tid == 0) p.enc_a_state[64] = 1.0f;  // every thread has read the flag before the barrier above
    }

    RN_STAMP(6)
    // ---- hoisted terms of the head: audio code through the ambient layer, eye through the sigma layer, individual code
    //      through the colour layer
    if (p.auds) run_layers(cx, pg.n_net, pg.n_head);
    else if (tid < 64) p.head_consts[tid] = 0.f;
    if (tid < 64) {
        const float e = p.eye ? h16(__ldg(p.eye)) : 0.f;
        p.head_consts[64 + tid] = p.eye ? __half2float(__ldg(p.w_sig1 + (size_t)tid * 65 + 64)) * e : 0.f;
        float acc = 0.f;
        if (p.ind_code)
            for (uint32_t k = 0; k < 4; ++k) acc = __fmaf_rn(__half2float(__ldg(p.w_col1 + (size_t)tid * 84 + 80 + k)), h16(__ldg(p.ind_code + k)), acc);
        p.head_consts[128 + tid] = acc;
    }
    __syncthreads();

    RN_STAMP(7)
    // ---- hoisted terms of the torso: [freq(pose6) (54) | individual code (8)], zero-padded to 64
    if (p.w_def1) {
        // the head pose as (XYZ Euler angles, translation): taken from the caller, or derived here from the 4x4 cam2world
        // (convert_poses -> matrix_to_euler_angles(R, 'XYZ'), nerf/utils.py:113-170, 230-237: for R = Rx(a) Ry(b) Rz(c),
        //  a = atan2(-R12, R22), b = asin(R02), c = atan2(-R01, R00))
        __shared__ float s_pose6[6];
        if (tid < 6) {
            float v;
            if (p.pose44) {
                const float* P = p.pose44;
                if (tid == 0) v = atan2f(-__ldg(P + 6), __ldg(P + 10));
                else if (tid == 1) v = asinf(__ldg(P + 2));
                else if (tid == 2) v = atan2f(-__ldg(P + 1), __ldg(P + 0));
                else v = __ldg(P + 4 * (tid - 3) + 3);
            } else {
                v = __ldg(p.pose6 + tid);
            }
            s_pose6[tid] = v;
            if (p.pose6_out) p.pose6_out[tid] = v;
        }
        __syncthreads();
        if (tid < 54) {
            const uint32_t c = tid;
            float v;
            if (c < 6) v = s_pose6[c];
            else {
                const uint32_t col = c / 6 - 1, d = c % 6, fr = col / 2;
                v = __sinf(scalbnf(s_pose6[d], (int)fr) + (float)(col % 2) * 1.5707963705062866f);
            }
            s_vec[c] = __float2half_rn(v);
        } else if (tid < 64) {
            s_vec[tid] = __float2half_rn((tid < 62 && p.ind_torso) ? __ldg(p.ind_torso + (tid - 54)) : 0.f);
        }
        __syncthreads();
        run_layers(cx, pg.n_net + pg.n_head, pg.n_torso);
    }
    RN_STAMP(8)
}

}  // namespace

// ---- host side: the layer table ------------------------------------------------------------------------------------------------
namespace {

inline uint32_t pad16(uint32_t c) { return (c + 15u) & ~15u; }

struct ProgramBuilder {
    AudioProgram pg{};
    uint32_t arena = 0, nl = 0;
    // weight matrix [rows][ld] halves -> arena slot i; an odd row pitch (only Cin = 29: 87) is re-strided to an even one
    void add_slot(int i, const __half* src, uint32_t rows, uint32_t ld) {
        AudioSlot& s = pg.slots[i];
        const uint32_t bytes = rows * ld * 2;
        const bool tma = (ld % 2 == 0) && (bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
        s.src = src; s.off = arena; s.rows = (uint16_t)rows; s.ld_src = (uint16_t)ld; s.ld_dst = (uint16_t)(ld + (ld & 1)); s.tma = tma ? 1 : 0;
        s.bytes = tma ? bytes : rows * s.ld_dst * 2;
        arena += (s.bytes + 15u) & ~15u;
    }
    AudioLayer& layer(int slot, uint32_t col0, uint32_t M, uint32_t N, uint32_t Cin) {
        AudioLayer& L = pg.layers[nl++];
        L = AudioLayer{};
        L.w_off = pg.slots[slot].off; L.slot = (uint16_t)slot; L.ld = pg.slots[slot].ld_dst; L.col0 = (uint16_t)col0;
        L.M = (uint16_t)M; L.N = (uint16_t)N; L.Cin = (uint16_t)Cin; L.taps = 1; L.stride = 1; L.Lout = 1;
        return L;
    }
    // Conv1d(k=3, pad=1, stride) + LeakyReLU: input [F][Lin + 2][in_cs] at in_buf/in_off, output [F][Lout + 2*out_pad][out_cs]
    void conv(int slot, uint32_t Cin, uint32_t Cout, uint32_t F, uint32_t Lin, uint32_t stride, uint16_t in_buf, uint16_t in_off, uint32_t in_cs,
              uint16_t out_buf, uint16_t out_off, uint32_t out_cs, uint32_t out_pad, uint16_t bias_row, uint16_t stamp) {
        const uint32_t Lout = (Lin + 2 - 3) / stride + 1;
        AudioLayer& L = layer(slot, 0, Cout, F * Lout, Cin);
        L.taps = 3; L.stride = (uint16_t)stride; L.Lout = (uint16_t)Lout;
        L.in_buf = in_buf; L.in_off = in_off; L.in_fs = (uint16_t)((Lin + 2) * in_cs); L.in_cs = (uint16_t)in_cs;
        L.out_buf = out_buf; L.out_off = out_off; L.out_fs = (uint16_t)((Lout + 2 * out_pad) * out_cs); L.out_cs = (uint16_t)out_cs; L.out_pad = (uint16_t)out_pad;
        L.zero_halves = (uint16_t)((out_pad || out_cs != Cout) ? std::max<uint32_t>(F * L.out_fs, 16u) : 0u);
        L.mode = ST_CONV; L.bias_row = bias_row; L.stamp = stamp;
    }
    // Linear on R rows: in[r*K + k] -> out[(r + out_pad)*Nout + m]
    void linear(int slot, uint32_t K, uint32_t Nout, uint32_t R, uint16_t in_buf, uint16_t in_off, uint16_t out_buf, uint16_t out_off,
                uint32_t out_pad, bool act, uint16_t bias_row, uint16_t stamp) {
        AudioLayer& L = layer(slot, 0, Nout, R, K);
        L.in_buf = in_buf; L.in_off = in_off; L.in_fs = (uint16_t)K; L.in_cs = (uint16_t)K;
        L.out_buf = out_buf; L.out_off = out_off; L.out_fs = (uint16_t)Nout; L.out_cs = (uint16_t)Nout; L.out_pad = (uint16_t)out_pad;
        L.zero_halves = (uint16_t)(out_pad ? (R + 2 * out_pad) * Nout : 0u);
        L.mode = act ? ST_LINEAR_ACT : ST_LINEAR; L.bias_row = bias_row; L.stamp = stamp;
    }
    // out[m] = sum_k W[m, col0 + k] * vec[k]   (fp32 result in global memory; vec zero-padded to a multiple of 16)
    void hoist(int slot, uint32_t col0, uint32_t M, uint32_t K, uint16_t out_buf, uint16_t out_off) {
        AudioLayer& L = layer(slot, col0, M, 1, K);
        L.in_buf = BUF_VEC; L.out_buf = out_buf; L.out_off = out_off; L.mode = ST_VEC;
    }
};

}  // namespace

void set_audio_prof(void* p) { cudaMemcpyToSymbol(g_audio_prof_dev, &p, sizeof(p)); }

int launch_audio_frame(const AudioParams& p, cudaStream_t st) {
    static bool attr_set[64] = {};           // the opt-in is a per-device function attribute: remembered per device, not per process
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        RN_REQUIRE(cudaFuncSetAttribute(audio_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AUDIO_SMEM_BYTES) == cudaSuccess,
                   "shared memory opt-in failed");
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    ProgramBuilder b;
    const uint32_t F = p.F, Cin = p.Cin;
    if (p.auds) {
        const __half* bias[12] = {p.conv_b[0], p.conv_b[1], p.conv_b[2], p.conv_b[3], p.fc_b[0], p.fc_b[1],
                                  p.att_b[0], p.att_b[1], p.att_b[2], p.att_b[3], p.att_b[4], p.att_fc_b};
        const uint16_t bias_n[12] = {32, 32, 64, 64, 64, 64, 16, 8, 4, 2, 1, (uint16_t)F};
        for (int i = 0; i < 12; ++i) { b.pg.bias_src[i] = bias[i]; b.pg.bias_n[i] = (i < 6 || p.att > 0) ? bias_n[i] : 0; }
        const uint32_t cp0 = pad16(Cin), att_pad = p.att > 0 ? 1u : 0u;
        b.pg.cp0 = (uint16_t)cp0;
        b.add_slot(0, p.conv_w[0], 32, Cin * 3); b.add_slot(1, p.conv_w[1], 32, 96);
        b.add_slot(2, p.conv_w[2], 64, 96);      b.add_slot(3, p.conv_w[3], 64, 192);
        b.add_slot(4, p.fc_w[0], 64, 64);        b.add_slot(5, p.fc_w[1], 64, 64);
        // AudioNet (nerf/network.py:36-67): [F,Cin,16] -> [F,32,8] -> [F,32,4] -> [F,64,2] -> [F,64,1] -> fc 64 -> 64
        b.conv(0, Cin, 32, F, 16, 2, BUF_A, 0, cp0, BUF_B, 0, 32, 1, 0, 2);
        b.conv(1, 32, 32, F, 8, 2, BUF_B, 0, 32, BUF_A, 0, 32, 1, 1, 0);
        b.conv(2, 32, 64, F, 4, 2, BUF_A, 0, 32, BUF_B, 0, 64, 1, 2, 0);
        b.conv(3, 64, 64, F, 2, 2, BUF_B, 0, 64, BUF_A, 0, 64, 0, 3, 3);
        b.linear(4, 64, 64, F, BUF_A, 0, BUF_B, 0, 0, true, 4, 0);
        b.linear(5, 64, 64, F, BUF_B, 0, BUF_A, 0, att_pad, false, 5, 4);   // x = hA rows att_pad .. att_pad+F of [.][64]
        if (p.att > 0) {
            // AudioAttNet (nerf/network.py:10-33): Conv1d stack 64 -> 16 -> 8 -> 4 -> 2 -> 1 along the F frames (x is already
            // position-major for it: frames are the positions), then Linear(F, F).  Temporaries ping-pong in hB at 0 / 256.
            b.add_slot(6, p.att_w[0], 16, 192); b.add_slot(7, p.att_w[1], 8, 48); b.add_slot(8, p.att_w[2], 4, 24);
            b.add_slot(9, p.att_w[3], 2, 12);   b.add_slot(10, p.att_w[4], 1, 6); b.add_slot(11, p.att_fc_w, F, F);
            b.conv(6, 64, 16, 1, F, 1, BUF_A, 0, 64, BUF_B, 0, 16, 1, 6, 0);
            b.conv(7, 16, 8, 1, F, 1, BUF_B, 0, 16, BUF_B, 256, 16, 1, 7, 0);
            b.conv(8, 8, 4, 1, F, 1, BUF_B, 256, 16, BUF_B, 0, 16, 1, 8, 0);
            b.conv(9, 4, 2, 1, F, 1, BUF_B, 0, 16, BUF_B, 256, 16, 1, 9, 0);
            b.conv(10, 2, 1, 1, F, 1, BUF_B, 256, 16, BUF_B, 0, 1, 0, 10, 5);   // y[0..F) ...
            b.pg.layers[b.nl - 1].zero_halves = 16;                             // ... inside 16 cleared halves: the Linear reads K padded to 16
            b.linear(11, F, F, 1, BUF_B, 0, BUF_B, 256, 0, false, 11, 0);      // logits
        }
        b.pg.n_net = (uint16_t)b.nl;
        b.add_slot(12, p.w_amb1, 64, 96);
        b.hoist(12, 32, 64, 64, BUF_HEAD, 0);
        b.pg.n_head = 1;
    }
    if (p.w_def1) {
        const uint32_t K = p.ind_torso ? 62u : 54u;
        b.add_slot(13, p.w_def1, 64, 42 + K); b.add_slot(14, p.w_tor1, 32, 74 + K);
        b.hoist(13, 42, 64, K, BUF_TORSO, 0);
        b.hoist(14, 74, 32, K, BUF_TORSO, 64);
        b.pg.n_torso = 2;
    }
    RN_REQUIRE(b.arena <= ARENA_MAX_BYTES, "weight arena overflow");
    audio_frame_kernel<<<1, 512, AUDIO_SMEM_BYTES, st>>>(p, b.pg);
    return finish_launch("audio_frame");
}

}  // namespace rn
