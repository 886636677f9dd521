// EXPERIMENTAL (round 1: compiles, NOT yet validated on a GPU; off unless NFLOWS_B200_TRUNK_KERNEL=1 -- see DESIGN.md section 8).
//
// The square layers of a conditioner trunk (ResidualNet blocks, nn/nets/resnet.py:9-55 in the reference) as ONE persistent
// kernel: a CTA keeps the activation pair of a 128-row tile RESIDENT in shared memory as the K-major A operand of the next
// layer, streams only the weights (L2-resident, multicast across a 2-CTA cluster), and hands the last layer's pair to the
// consumer with TMA stores straight out of that buffer.  Motivation (profiles/ncu_summary_r1.md): run one launch per
// layer, the 256-wide layers are bound by per-tile epilogue latency and output round trips (tensor pipe 10 %, DRAM 31 %),
// not by MMAs or HBM bandwidth.
//
//   h_{l+1} = post_l( pre_l(h_l) W_l^T + b_l ) [+ skip]          l = 0 .. L-1,  every W_l is H x H,  H <= 256, H % 32 == 0
//
// Same arithmetic as nfk_linear_tc.cu: fp16 (hi, lo) split pairs, three kind::f16 MMAs per K-step, partial sums over two
// K-slabs drained from TMEM into registers (round-to-nearest), exact power-of-two scaling.  Shared memory: resident pair
// 8 slabs x (hi 8 KB | lo 8 KB) = 128 KB, weight ring 3 x (W hi 16 KB | W lo 16 KB) = 96 KB.  Skip connections read / write
// fp32 tensors in global memory (L2-resident at this tile rate).  Within a tile the layers are sequential (layer l+1
// multiplies what layer l's epilogue wrote); only the drains overlap with MMAs.
#include <stdlib.h>

#include "tc_common.cuh"

namespace nfk {
namespace tc {

constexpr int TRUNK_MAX_LAYERS = 8;
constexpr int TRUNK_W_STAGES = 3;
constexpr int TRUNK_SLAB_BYTES = 2 * A_BYTES;                         // one K-slab of the resident pair: hi | lo
constexpr int TRUNK_A_BYTES = (BN_MAX / BK) * TRUNK_SLAB_BYTES;      // 128 KB
constexpr int TRUNK_W_STAGE_BYTES = 2 * B_BYTES;                      // 32 KB
constexpr int TRUNK_BAR_OFF = TRUNK_A_BYTES + TRUNK_W_STAGES * TRUNK_W_STAGE_BYTES;
constexpr int TRUNK_SMEM_BYTES = TRUNK_BAR_OFF + 1024 /*alignment slack*/ + 256 /*barriers*/;
static_assert(TRUNK_SMEM_BYTES <= 232448, "trunk kernel shared memory");

// layer_flags bits
constexpr int TL_RELU_OUT = 1;     // relu on (acc + bias)
constexpr int TL_ADD_SKIP = 2;     // + current skip tensor (never combined with TL_RELU_OUT)
constexpr int TL_SAVE_SKIP = 4;    // the fp32 result is the skip tensor of a later layer: written to skip_buf
constexpr int TL_SPLIT_RELU = 8;   // the consumer of this layer's output applies relu to its input

struct TrunkParams {
    const float* bias;        // [L * H]
    const float* skip_in;     // fp32 [n_rows, ld_skip]: skip tensor of the first TL_ADD_SKIP layer
    float* skip_buf;          // fp32 [n_rows, ld_skip] scratch for TL_SAVE_SKIP results (may be null if no layer saves)
    int64_t ld_skip;
    int32_t* flags;
    int64_t n_rows;
    int H, num_layers, num_m_tiles;
    int layer_flags[TRUNK_MAX_LAYERS];
    float acc_scale[TRUNK_MAX_LAYERS];       // 2^(e_act + e_w[l]): the accumulators of layer l hold (A W^T) * acc_scale
    float inv_acc_scale[TRUNK_MAX_LAYERS];
    float out_scale;                         // 2^e_act: exponent of every activation pair (input, intermediate, output)
};

__device__ __forceinline__ void trunk_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}

template <int CL>
__global__ void __launch_bounds__(THREADS, 1)
residual_trunk_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                      const __grid_constant__ CUtensorMap map_y_hi, const __grid_constant__ CUtensorMap map_y_lo,
                      const TrunkParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t w_ring = smem_base + TRUNK_A_BYTES;
    const uint32_t bars = smem_base + TRUNK_BAR_OFF;
    const uint32_t bar_wfull = bars, bar_wempty = bars + 8 * TRUNK_W_STAGES;
    const uint32_t bar_tfull = bars + 16 * TRUNK_W_STAGES, bar_tempty = bar_tfull + 16;
    const uint32_t bar_afull = bar_tempty + 16;      // input pair of a tile has landed in the resident buffer
    const uint32_t bar_aready = bar_afull + 8;       // the epilogue warps have written the next layer's operand
    const uint32_t bar_outready = bar_aready + 8;    // ... the last layer's pair: ready for the TMA stores
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + TRUNK_BAR_OFF + 16 * TRUNK_W_STAGES + 64);

    uint32_t tid_x;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
    const int warp = tid_x >> 5, lane = tid_x & 31;
    const int num_k = p.H / BK;                                  // K-slabs per layer (H is a multiple of 32)
    const int num_groups = (num_k + DRAIN_SLABS_LINEAR - 1) / DRAIN_SLABS_LINEAR;

    if (tid_x == 0) {
        for (int s = 0; s < TRUNK_W_STAGES; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, CL); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 8); }
        mbar_init(bar_afull, 1); mbar_init(bar_aready, 8); mbar_init(bar_outready, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_hi); prefetch_tmap(&map_w_lo);
        prefetch_tmap(&map_y_hi); prefetch_tmap(&map_y_lo);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int cta_rank = CL > 1 ? (int)cluster_ctarank() : 0;
    constexpr uint16_t cl_mask = (uint16_t)((1u << CL) - 1);
    const int units = (p.num_m_tiles + CL - 1) / CL;             // groups of CL neighbouring 128-row tiles
    const int first = blockIdx.x / CL, step = gridDim.x / CL;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
        if (warp == 0) {
            // ================================================= TMA producer (one elected thread): input pair, weight stream, output stores
            if (elect_one()) {
                const uint32_t w_tx = 2u * (uint32_t)p.H * ROW_BYTES;
                const int wrows = p.H / CL;
                int stage = 0; uint32_t phase = 0, out_phase = 0;
                for (int u = first; u < units; u += step) {
                    const int m0 = (u * CL + cta_rank) * BM;
                    mbar_expect_tx(bar_afull, (uint32_t)num_k * TRUNK_SLAB_BYTES);
                    for (int ks = 0; ks < num_k; ++ks) {
                        tma_load_2d(smem_base + ks * TRUNK_SLAB_BYTES, &map_a_hi, bar_afull, ks * BK, m0);
                        tma_load_2d(smem_base + ks * TRUNK_SLAB_BYTES + A_BYTES, &map_a_lo, bar_afull, ks * BK, m0);
                    }
                    for (int l = 0; l < p.num_layers; ++l) {
                        for (int ks = 0; ks < num_k; ++ks) {
                            mbar_wait(bar_wempty + 8 * stage, phase ^ 1);
                            const uint32_t full = bar_wfull + 8 * stage;
                            const uint32_t sw = w_ring + stage * TRUNK_W_STAGE_BYTES;
                            mbar_expect_tx(full, w_tx);
                            if (CL == 1) {
                                tma_load_2d(sw, &map_w_hi, full, ks * BK, l * p.H);
                                tma_load_2d(sw + B_BYTES, &map_w_lo, full, ks * BK, l * p.H);
                            } else {
                                const uint32_t off = (uint32_t)(cta_rank * wrows) * ROW_BYTES;
                                tma_load_2d_multicast(sw + off, &map_w_hi, full, ks * BK, l * p.H + cta_rank * wrows, cl_mask);
                                tma_load_2d_multicast(sw + B_BYTES + off, &map_w_lo, full, ks * BK, l * p.H + cta_rank * wrows, cl_mask);
                            }
                            if (++stage == TRUNK_W_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    // the epilogue warps have left the last layer's pair in the resident buffer: send it out, and do not load
                    // the next tile over it before the stores have read it
                    mbar_wait(bar_outready, out_phase);
                    out_phase ^= 1;
                    for (int ks = 0; ks < num_k; ++ks) {
                        trunk_store_2d(&map_y_hi, smem_base + ks * TRUNK_SLAB_BYTES, ks * BK, m0);
                        trunk_store_2d(&map_y_lo, smem_base + ks * TRUNK_SLAB_BYTES + A_BYTES, ks * BK, m0);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            }
        } else if (warp == 1) {
            // ================================================= MMA issuer (one elected thread: uniform datapath, tc_common.cuh)
            if (elect_one()) {
            const bool leader = true;
            const uint32_t idesc = make_idesc(p.H);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t afull_phase = 0, aready_phase = 0;
            for (int u = first; u < units; u += step) {
                for (int l = 0; l < p.num_layers; ++l) {
                    if (l == 0) { mbar_wait(bar_afull, afull_phase); afull_phase ^= 1; }
                    else { mbar_wait(bar_aready, aready_phase); aready_phase ^= 1; }
                    tc_fence_after();
                    for (int g = 0; g < num_groups; ++g) {
                        const int slabs = min(DRAIN_SLABS_LINEAR, num_k - g * DRAIN_SLABS_LINEAR);
                        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                        const uint32_t d_tmem = tmem_base + acc * BN_MAX;
                        int st = stage; uint32_t ph = phase;
                        for (int j = 0; j < slabs; ++j) {
                            mbar_wait(bar_wfull + 8 * st, ph);
                            if (++st == TRUNK_W_STAGES) { st = 0; ph ^= 1; }
                        }
                        tc_fence_after();
                        st = stage;
                        for (int j = 0; j < slabs; ++j) {              // cross terms of the group first (small magnitudes)
                            const uint32_t sa = smem_base + (g * DRAIN_SLABS_LINEAR + j) * TRUNK_SLAB_BYTES;
                            const uint32_t sw = w_ring + st * TRUNK_W_STAGE_BYTES;
                            const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + A_BYTES);
                            const uint64_t w_hi = make_smem_desc(sw), w_lo = make_smem_desc(sw + B_BYTES);
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk) {
                                const uint64_t adv = (uint64_t)(kk * 2);
                                if (leader) umma_f16(d_tmem, a_lo + adv, w_hi + adv, idesc, (j | kk) != 0);
                                if (leader) umma_f16(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
                            }
                            if (++st == TRUNK_W_STAGES) st = 0;
                        }
                        for (int j = 0; j < slabs; ++j) {              // then the main products; each releases its weight slab
                            const uint32_t sa = smem_base + (g * DRAIN_SLABS_LINEAR + j) * TRUNK_SLAB_BYTES;
                            const uint32_t sw = w_ring + stage * TRUNK_W_STAGE_BYTES;
                            const uint64_t a_hi = make_smem_desc(sa), w_hi = make_smem_desc(sw);
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk) {
                                const uint64_t adv = (uint64_t)(kk * 2);
                                if (leader) umma_f16(d_tmem, a_hi + adv, w_hi + adv, idesc, 1);
                            }
                            if (!leader) {} else if (CL == 1) umma_commit(bar_wempty + 8 * stage);
                            else umma_commit_multicast(bar_wempty + 8 * stage, cl_mask);
                            if (++stage == TRUNK_W_STAGES) { stage = 0; phase ^= 1; }
                        }
                        if (leader) umma_commit(bar_tfull + 8 * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                }
            }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
        // ================================================= accumulate + layer epilogue: 8 warps
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int r_tile = q * 32 + lane;                         // row inside the 128-row tile
        const int n0 = half * HALF;                               // first column this thread owns
        int acc = 0; uint32_t acc_phase = 0;
        int flag = 0;
        for (int u = first; u < units; u += step) {
            const int64_t row = (int64_t)(u * CL + cta_rank) * BM + r_tile;
            const bool row_ok = row < p.n_rows;
            const float* skip = p.skip_in;
            for (int l = 0; l < p.num_layers; ++l) {
                const int lf = p.layer_flags[l];
                const bool last = l == p.num_layers - 1;
                float sum[HALF];
                // running sums start from the skip tensor (loads issued first, back to back) and the bias, in the accumulators'
                // power-of-two scaled domain
#pragma unroll
                for (int c = 0; c < HALF; c += 4) {
                    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((lf & TL_ADD_SKIP) && row_ok && n0 + c < p.H)
                        r4 = *reinterpret_cast<const float4*>(skip + row * p.ld_skip + n0 + c);
                    sum[c] = r4.x; sum[c + 1] = r4.y; sum[c + 2] = r4.z; sum[c + 3] = r4.w;
                }
#pragma unroll
                for (int c = 0; c < HALF; c += 4) {
                    if (n0 + c < p.H) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + l * p.H + n0 + c));
                        sum[c] += b4.x; sum[c + 1] += b4.y; sum[c + 2] += b4.z; sum[c + 3] += b4.w;
                    }
                }
                const float as = p.acc_scale[l], ias = p.inv_acc_scale[l];
#pragma unroll
                for (int c = 0; c < HALF; ++c) sum[c] *= as;
                for (int g = 0; g < num_groups; ++g) {
                    mbar_wait(bar_tfull + 8 * acc, acc_phase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN_MAX + half * HALF;
#pragma unroll
                    for (int c = 0; c < HALF; c += 64) {
                        uint32_t raw[2][32];
                        tmem_ld32(taddr + c, raw[0]);
                        tmem_ld32(taddr + c + 32, raw[1]);
                        tmem_ld_wait();
#pragma unroll
                        for (int v = 0; v < 2; ++v) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float2 r2 = __fadd2_rn(make_float2(sum[c + 32 * v + j], sum[c + 32 * v + j + 1]),
                                                             make_float2(__uint_as_float(raw[v][j]), __uint_as_float(raw[v][j + 1])));
                                sum[c + 32 * v + j] = r2.x;
                                sum[c + 32 * v + j + 1] = r2.y;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                // ---- layer epilogue.  Every MMA of this layer has retired (the last partial sum was drained), so the resident
                // buffer may be overwritten with the operand of the next layer.
#pragma unroll
                for (int c = 0; c < HALF; ++c) {
                    float x = sum[c] * ias;
                    if (lf & TL_RELU_OUT) x = fmaxf(x, 0.0f);
                    sum[c] = x;
                }
                if ((lf & TL_SAVE_SKIP) && row_ok) {
#pragma unroll
                    for (int c = 0; c < HALF; c += 4)
                        if (n0 + c < p.H)
                            *reinterpret_cast<float4*>(p.skip_buf + row * p.ld_skip + n0 + c) =
                                make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
                }
                float amax = 0.0f;
#pragma unroll
                for (int s = 0; s < HALF / BK; ++s) {                 // 4 K-slabs of 32 columns per thread
                    const int slab = half * (HALF / BK) + s;
                    if (slab < num_k) {
                        uint8_t* base = smem_gen + slab * TRUNK_SLAB_BYTES + r_tile * ROW_BYTES;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {                 // 16-byte pieces of the 64-byte row, SWIZZLE_64B placement
                            __half2 h2[4], l2[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float x0 = sum[32 * s + 8 * c + 2 * e], x1 = sum[32 * s + 8 * c + 2 * e + 1];
                                if (lf & TL_SPLIT_RELU) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); }
                                x0 *= p.out_scale; x1 *= p.out_scale;
                                amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
                                h2[e] = __floats2half2_rn(x0, x1);
                                const float2 hf = __half22float2(h2[e]);
                                l2[e] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                            }
                            const int piece = (c ^ ((r_tile >> 1) & 3)) * 16;
                            *reinterpret_cast<uint4*>(base + piece) = *reinterpret_cast<const uint4*>(h2);
                            *reinterpret_cast<uint4*>(base + A_BYTES + piece) = *reinterpret_cast<const uint4*>(l2);
                        }
                    }
                }
                if (row_ok && !(amax <= 65000.0f)) flag |= 4;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tcgen05.mma / TMA store reads
                __syncwarp();
                if (lane == 0) mbar_arrive(last ? bar_outready : bar_aready);
                if (lf & TL_SAVE_SKIP) skip = p.skip_buf;
            }
        }
        if (flag && p.flags) atomicOr(p.flags, flag);
    }

    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace nfk

using namespace nfk;

extern "C" int nfk_residual_trunk_f16x3_supported(int32_t hidden_features, int32_t num_layers, int64_t lda, int64_t ldw) {
    return (hidden_features >= 32 && hidden_features <= tc::BN_MAX && hidden_features % 32 == 0 && num_layers >= 1 &&
            num_layers <= tc::TRUNK_MAX_LAYERS && lda % 8 == 0 && ldw % 8 == 0) ? 1 : 0;
}

extern "C" int nfk_residual_trunk_f16x3(const void* a_hi_, const void* a_lo_, int64_t lda, int32_t act_exp, const void* w_hi_,
                                        const void* w_lo_, int64_t ldw, const int32_t* w_exps, const float* bias,
                                        const int32_t* layer_flags, int32_t num_layers, const float* skip_in, float* skip_buf,
                                        int64_t ld_skip, void* y_hi_, void* y_lo_, int64_t lds, int64_t n_rows,
                                        int32_t hidden_features, int32_t* flags, void* stream) {
    const __half* a_hi = (const __half*)a_hi_; const __half* a_lo = (const __half*)a_lo_;
    const __half* w_hi = (const __half*)w_hi_; const __half* w_lo = (const __half*)w_lo_;
    __half* y_hi = (__half*)y_hi_; __half* y_lo = (__half*)y_lo_;
    NFK_REQUIRE(n_rows >= 0, "bad sizes");
    if (n_rows == 0) return NFK_OK;
    NFK_REQUIRE(a_hi && a_lo && w_hi && w_lo && bias && y_hi && y_lo && w_exps && layer_flags, "NULL pointer");
    NFK_REQUIRE(nfk_residual_trunk_f16x3_supported(hidden_features, num_layers, lda, ldw) && lds % 8 == 0,
                "trunk kernel takes hidden_features in {32, 64, ..., 256}, up to %d layers, row pitches multiples of 8",
                tc::TRUNK_MAX_LAYERS);
    NFK_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(w_hi) && aligned16(w_lo) && aligned16(y_hi) && aligned16(y_lo) &&
                    aligned16(bias),
                "operands must be 16-byte aligned");
    NFK_REQUIRE(n_rows < (1ll << 31), "n_rows too large for one launch");
    tc::TrunkParams p;
    bool any_skip = false, any_save = false;
    for (int l = 0; l < num_layers; ++l) {
        const int lf = layer_flags[l];
        NFK_REQUIRE(!((lf & tc::TL_ADD_SKIP) && (lf & tc::TL_RELU_OUT)), "layer %d: skip add after a relu output is not supported", l);
        NFK_REQUIRE(act_exp + w_exps[l] >= -60 && act_exp + w_exps[l] <= 60, "scale exponent out of range");
        p.layer_flags[l] = lf;
        p.acc_scale[l] = ldexpf(1.0f, act_exp + w_exps[l]);
        p.inv_acc_scale[l] = ldexpf(1.0f, -(act_exp + w_exps[l]));
        any_skip |= (lf & tc::TL_ADD_SKIP) != 0;
        any_save |= (lf & tc::TL_SAVE_SKIP) != 0;
    }
    NFK_REQUIRE(!any_skip || (skip_in && ld_skip % 4 == 0 && aligned16(skip_in)), "skip tensor missing or not 16-byte aligned");
    NFK_REQUIRE(!any_save || (skip_buf && ld_skip % 4 == 0 && aligned16(skip_buf)), "skip scratch missing or not 16-byte aligned");
    p.bias = bias; p.skip_in = skip_in; p.skip_buf = skip_buf; p.ld_skip = ld_skip; p.flags = flags; p.n_rows = n_rows;
    p.H = hidden_features; p.num_layers = num_layers; p.out_scale = ldexpf(1.0f, act_exp);
    p.num_m_tiles = (int)((n_rows + tc::BM - 1) / tc::BM);

    static int cluster_pref = 0;
    if (!cluster_pref) {
        const char* e = getenv("NFK_CLUSTER");
        cluster_pref = (e && e[0] == '1') ? 1 : 2;
    }
    const int CL = (cluster_pref == 2 && p.num_m_tiles >= 2 && hidden_features % 16 == 0) ? 2 : 1;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo, my_hi, my_lo;
    int rc;
    if ((rc = tc::make_map(&ma_hi, a_hi, n_rows, hidden_features, lda, tc::BM))) return rc;
    if ((rc = tc::make_map(&ma_lo, a_lo, n_rows, hidden_features, lda, tc::BM))) return rc;
    if ((rc = tc::make_map(&mw_hi, w_hi, (int64_t)num_layers * hidden_features, hidden_features, ldw, hidden_features / CL))) return rc;
    if ((rc = tc::make_map(&mw_lo, w_lo, (int64_t)num_layers * hidden_features, hidden_features, ldw, hidden_features / CL))) return rc;
    if ((rc = tc::make_map(&my_hi, y_hi, n_rows, hidden_features, lds, tc::BM))) return rc;
    if ((rc = tc::make_map(&my_lo, y_lo, n_rows, hidden_features, lds, tc::BM))) return rc;

    static DeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(tc::residual_trunk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TRUNK_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::residual_trunk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TRUNK_SMEM_BYTES);
        if (e != cudaSuccess) return fail(NFK_E_CUDA, "cudaFuncSetAttribute(smem=%d): %s", tc::TRUNK_SMEM_BYTES, cudaGetErrorString(e));
        attr_once.mark(attr_dev);
    }
    const int units = (p.num_m_tiles + CL - 1) / CL;
    const int max_clusters = tc::sm_count() / CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CL * (units < max_clusters ? units : max_clusters)));
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = tc::TRUNK_SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = (CL == 2) ? cudaLaunchKernelEx(&cfg, tc::residual_trunk_kernel<2>, ma_hi, ma_lo, mw_hi, mw_lo, my_hi, my_lo, p)
                               : cudaLaunchKernelEx(&cfg, tc::residual_trunk_kernel<1>, ma_hi, ma_lo, mw_hi, mw_lo, my_hi, my_lo, p);
    if (le != cudaSuccess) return fail(NFK_E_CUDA, "cudaLaunchKernelEx(residual_trunk_kernel, cluster %d): %s", CL, cudaGetErrorString(le));
    return check_launch("residual_trunk_kernel");
}
