// qttt_kernels.cu -- sm_100a kernels and the C ABI (include/qttt_b200.h) of libqttt_b200.so.
//
// All kernels are one-thread-per-game integer programs over the packed 16-byte state of
// qttt_core.cuh: one coalesced 128-bit load and store per game, lookup tables staged in
// shared memory, no tensor cores (nothing here is a contraction).  No torch headers: the
// Python side passes raw device pointers and a stream.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/qttt_b200.h"
#include "qttt_core.cuh"
#include "qttt_mcts.cuh"

namespace qttt {

__device__ __align__(128) const LutImage g_lut = make_lut_image();

// Tables of the observation kernel only (k_observe): byte selectors that compact the codes of
// the uncollapsed move slots of one player into q_states_p1 / q_states_p2 (env.py:73-78).
struct ObsLutImage {
    uint32_t q1sel[32][4];   // which of the slots 0,2,4,6,8 are uncollapsed -> selA0, selB0, selA1, selB1
    uint32_t q2sel[16][4];   // the same for slots 1,3,5,7: two selectors and two padding masks
};
constexpr int kObsLutBytes = (int)sizeof(ObsLutImage);
static_assert(kObsLutBytes == 768, "ObsLutImage layout");
constexpr ObsLutImage make_obs_lut() {
    ObsLutImage t{};
    // q_states_p1: five candidate codes k0..k4 live in bytes 0..9 of (S0, S1, S2) and bytes 10..11
    // of S2 hold the (-1, -1) padding.  Word w of the list is PRMT(PRMT(S0, S1, selA), S2, selB).
    for (uint32_t m = 0; m < 32; ++m) {
        int idx[5] = {5, 5, 5, 5, 5}, cnt = 0;            // 5 = the padding entry
        for (int j = 0; j < 5; ++j)
            if (m >> j & 1u) idx[cnt++] = j;
        for (int w = 0; w < 2; ++w) {
            uint32_t selA = 0, selB = 0;
            for (int p = 0; p < 4; ++p) {
                const int src = 2 * idx[2 * w + p / 2] + (p & 1);          // byte 0..11
                if (src < 8) { selA |= (uint32_t)src << (4 * p); selB |= (uint32_t)p << (4 * p); }
                else selB |= (uint32_t)(src - 8 + 4) << (4 * p);
            }
            t.q1sel[m][2 * w] = selA;
            t.q1sel[m][2 * w + 1] = selB;
        }
    }
    // q_states_p2: four candidates in (S0, S1); word w = PRMT(S0, S1, sel[w]) | pad[w]
    for (uint32_t m = 0; m < 16; ++m) {
        int idx[4] = {4, 4, 4, 4}, cnt = 0;
        for (int j = 0; j < 4; ++j)
            if (m >> j & 1u) idx[cnt++] = j;
        for (int w = 0; w < 2; ++w) {
            uint32_t sel = 0, pad = 0;
            for (int p = 0; p < 4; ++p) {
                const int e = idx[2 * w + p / 2];
                if (e < 4) sel |= (uint32_t)(2 * e + (p & 1)) << (4 * p);
                else pad |= 0xFFu << (8 * p);
            }
            t.q2sel[m][w] = sel;
            t.q2sel[m][2 + w] = pad;
        }
    }
    return t;
}
__device__ __align__(128) const ObsLutImage g_obs_lut = make_obs_lut();

constexpr int kThreads = 256;

// Stage the first `bytes` of the table image into shared memory: ONE bulk asynchronous copy
// (cp.async.bulk, the 1-D TMA path: global -> shared, completion counted in bytes on an
// mbarrier) issued by thread 0, instead of every thread looping over 16-byte loads and stores.
// With blocks that live for only a few chunks of games the staging is paid often, and as a
// loop it was ~90 instructions per thread per block.
__device__ __forceinline__ void stage_luts(uint8_t* smem, int bytes, uint8_t* smem2 = nullptr,
                                           const void* src2 = nullptr, int bytes2 = 0) {
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    const uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes + bytes2) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_a), "l"(&g_lut), "r"(bytes), "r"(bar_a) : "memory");
        if (smem2) {      // a second table image on the same barrier
            const uint32_t dst2_a = (uint32_t)__cvta_generic_to_shared(smem2);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst2_a), "l"(src2), "r"(bytes2), "r"(bar_a) : "memory");
        }
    }
    __syncthreads();                  // the barrier object is initialised for everybody
    uint32_t done = 0;                // every thread observes the completion itself: that is what
    while (!done) {                   // makes the bytes written by the copy engine visible to it
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a) : "memory");
    }
}

__device__ __forceinline__ State load_state(const qttt_state* p, int64_t i) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + i);
    return State{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ void store_state(qttt_state* p, int64_t i, const State& s) {
    *reinterpret_cast<uint4*>(p + i) = make_uint4(s.x, s.y, s.z, s.w);
}

// ------------------------------------------------------------------------------ work distribution
// Every batch kernel below cuts its games into chunks of kThreads and gives each block a short
// run of consecutive chunks (`iters`), with many more blocks than fit the machine at once.  The
// hardware block scheduler then refills an SM as soon as one of its blocks retires.  (Round 1
// launched exactly one resident wave of persistent grid-stride blocks: ncu showed the SMs
// running out of warps long before the kernel ended -- 46 of 64 warps resident on average at
// ply 8 -- because the warp arbiter is not fair and nothing replaces a warp that finishes early.)
constexpr int kStepIters = 8;       // chunks per block of the step kernels (tables restaged per block)
constexpr int kZcMaxIters = 8;      // chunks per block of the mapped-memory step kernel (its staging is sized for it)
constexpr int kSweepIters = 32;     // 8192 self-play games per block of the sweep (19 KB of tables per block)

static int chunk_grid(int64_t n, int iters) {
    const int64_t per_block = (int64_t)kThreads * iters;
    const int64_t g = (n + per_block - 1) / per_block;
    return (int)(g < 1 ? 1 : (g > 0x7FFFFFFF ? 0x7FFFFFFF : g));     // grid-stride loops take any grid
}

// ------------------------------------------------------------------------------ K2 reset
// Board.__init__ / Env.reset: empty games; optionally the outputs a step would have produced
// for them (legal mask of the empty board, reward -0.0f, not terminated, status ok).
__global__ void __launch_bounds__(kThreads) k_reset(qttt_state* state, uint64_t* mask, float* reward,
                                                    uint8_t* done, uint8_t* status, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        *reinterpret_cast<uint4*>(state + i) = make_uint4(0u, 0u, 0u, 0u);
        if (mask) mask[i] = 0xFFFFFFFFFull;   // all 36 pairs legal on the empty board
        if (reward) reinterpret_cast<uint32_t*>(reward)[i] = 0x80000000u;   // -0.0f: no line (env.py:49)
        if (done) done[i] = 0;
        if (status) status[i] = 0;
    }
}

// ------------------------------------------------------------------------------ K1 step
struct StepArgs {
    qttt_state* state;
    const uint8_t* action;
    const uint8_t* coin;
    uint64_t seed, game_base;
    uint32_t dword;             // Philox counter word 3: domain 0 | epoch << 8
    float* reward;
    uint8_t* done;
    uint64_t* mask;
    uint8_t* status;
    uint8_t* action_out;
    uint8_t* coin_out;
    uint32_t n;                 // <= 2^31 games per launch (32-bit indices)
    int iters;                  // chunks of kThreads games per block
};

struct StepIn { uint4 sv; uint32_t act, coin; };

// kFmt: QTTT_ACT_INDEX / QTTT_ACT_PAIR; kRandom: Philox policy instead of given actions;
// kFull: reward, done, mask and status are all requested and coins are given (no NULL tests);
// kMode: kStepPlain / kStepFresh / kStepAuto / kStepAutoNext (qttt_core.cuh).
// (Loading the next chunk's inputs before computing the current one was measured: 39 registers,
// 6 blocks per SM, 2 % slower overall.)
template <int kFmt, bool kRandom, bool kFull, int kMode>
__global__ void __launch_bounds__(kThreads, 8) k_step(const StepArgs a) {   // 8 blocks/SM: 32 registers
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    // Programmatic dependent launch (when the host launched with it, see launch_pdl): the NEXT
    // kernel in the stream may start placing blocks as soon as every block of this one has got
    // here, and this kernel stages its tables -- constants, independent of earlier kernels --
    // before it waits for the previous kernel to finish and its writes to be visible.  Back-to-back
    // steps overlap one kernel's tail with the next one's prologue.  (No-ops in a plain launch.)
    asm volatile("griddepcontrol.launch_dependents;");
    stage_luts(smem, kLutStepBytes);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const Luts L = luts_from_image(smem);

    auto load = [&](uint32_t i) {
        StepIn in;
        in.sv = make_uint4(0u, 0u, 0u, 0u);
        in.act = 255u;
        in.coin = 0u;
        if (i < a.n) {
            if (kMode != kStepFresh) in.sv = *reinterpret_cast<const uint4*>(a.state + i);
            if (!kRandom) {
                if (kFmt == QTTT_ACT_INDEX) {
                    in.act = a.action[i];
                } else {
                    const uchar2 ab = reinterpret_cast<const uchar2*>(a.action)[i];
                    in.act = (uint32_t)ab.x | ((uint32_t)ab.y << 8);
                }
                if (kFull || a.coin) in.coin = a.coin[i];
            }
        }
        return in;
    };
    // Lanes past the end of the batch run the step on an empty game with an illegal action and
    // store nothing, so that all 32 lanes of every warp execute step_game together (full-mask
    // warp votes, no __activemask query).
    auto run = [&](uint32_t i, const StepIn& in) {
        const bool valid = i < a.n;
        State s{in.sv.x, in.sv.y, in.sv.z, in.sv.w};
        uint32_t enew = 0u;
        if (!kRandom)
            enew = kFmt == QTTT_ACT_INDEX ? (uint32_t)L.pair[in.act] : pair_to_edge(in.act & 255u, in.act >> 8);
        const StepOut o = step_game<kRandom, kMode, true>(s, enew, kFull || a.coin != nullptr, in.coin & 1u, a.seed,
                                                          a.game_base + (uint64_t)i, a.dword, L);
        if (!valid) return;
        if (o.write_state) *reinterpret_cast<uint4*>(a.state + i) = make_uint4(s.x, s.y, s.z, s.w);
        // the reward is stored through an integer pointer: its two values differ only in bit
        // patterns (-0.0f / -1.0f), and a float-typed select gets "simplified" by the compiler
        // into an int->float conversion that loses the sign of zero
        if (kFull || a.reward) reinterpret_cast<uint32_t*>(a.reward)[i] = reward_bits(o.win);   // env.py:49
        if (kFull || a.done) a.done[i] = (uint8_t)o.done;                                       // env.py:51
        if (kFull || a.mask) a.mask[i] = L.legal[~o.classical & M9];                            // mcts.py:87-91
        if (kFull || a.status) a.status[i] = (uint8_t)o.status;
        if (kRandom) {
            if (a.action_out) a.action_out[i] = (uint8_t)o.action;
            if (a.coin_out) a.coin_out[i] = (uint8_t)o.coin;
        }
    };

    const uint32_t first = (blockIdx.x * (uint32_t)a.iters) * kThreads + threadIdx.x;
    if constexpr (kMode == kStepFresh && !kRandom) {
        // The fused reset + first step reads two bytes per game and writes 30: nothing to overlap
        // the load latency with inside one chunk, so the next chunk's bytes are fetched first.
        StepIn cur = load(first);
        for (int it = 0; it < a.iters; ++it) {
            const uint32_t i = first + (uint32_t)it * kThreads;
            StepIn nxt = cur;
            if (it + 1 < a.iters) nxt = load(i + kThreads);
            run(i, cur);
            cur = nxt;
        }
    } else {
        for (int it = 0; it < a.iters; ++it) {
            const uint32_t i = first + (uint32_t)it * kThreads;
            run(i, load(i));
        }
    }
}

// K1 with compact I/O (one byte in, one 16-bit word out per game): see qttt_step_packed.
// obs (optional): the post-step packed state is ALSO written there (the observation of a caller
// whose buffers are not the state array itself, e.g. mapped host memory).
// kPack12: results leave bit-packed, four games in three 16-bit words (see qttt_step_packed12_mapped):
// lane 4g+k (k < 3) writes word k of its group, with nibble k of lane 4g+3's result on top.
template <bool kPack12>
__global__ void __launch_bounds__(kThreads)
k_step_packed(qttt_state* __restrict__ state, const uint8_t* __restrict__ action_coin,
              uint16_t* __restrict__ result, qttt_state* __restrict__ obs, uint32_t n, int iters) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    stage_luts(smem, kLutStepBytes);
    const Luts L = luts_from_image(smem);
    const uint32_t first = (blockIdx.x * (uint32_t)iters) * kThreads + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const uint32_t i = first + (uint32_t)it * kThreads;
        if (!kPack12 && i >= n) break;
        uint32_t word = 0u;
        if (i < n) {
            uint4* sp = reinterpret_cast<uint4*>(state + i);
            const uint4 sv = *sp;
            State s{sv.x, sv.y, sv.z, sv.w};
            const uint32_t ac = action_coin[i];
            const StepResult r = step_core(s, (uint32_t)L.pair[ac & 63u], ac >> 7, L);
            const uint4 out = make_uint4(s.x, s.y, s.z, s.w);
            if (!r.illegal) *sp = out;
            if (obs) *reinterpret_cast<uint4*>(obs + i) = out;
            const uint32_t win = any_line(s, r.classical, L) != 0u;
            const uint32_t term = win | (uint32_t)(r.n > 8u);
            word = (~r.classical & M9) | (term << 9) | (win << 10) | (r.illegal << 11);
        }
        if (kPack12) {
            // (a warp's lanes are in or out of step_core together only by accident, but nothing in it
            // votes: step_core queries the active mask)
            const uint32_t k = threadIdx.x & 3u;
            const uint32_t last = __shfl_sync(0xFFFFFFFFu, word, (threadIdx.x & 31u) | 3u);
            if (k < 3u && (i & ~3u) < n)
                result[(i >> 2) * 3u + k] = (uint16_t)(word | (((last >> (4u * k)) & 15u) << 12));
        } else {
            result[i] = (uint16_t)word;
        }
    }
}

// The same step for buffers in MAPPED PINNED HOST memory, read and written by the kernel itself
// (no copy engine, no staging buffers): a block moves its chunk's 256 action bytes in as sixteen
// 16-byte loads and its 256 result words out as thirty-two 16-byte stores, so every PCIe
// transaction is a full 512-byte burst; the observation (16 B per game) is written by each
// thread directly (512 contiguous bytes per warp).
// Env.step for a host-resident caller that wants the OBSERVATION back: the step, then the env.py
// observation (env.py:68-85) and the step's flags as one 12-byte record per game instead of the
// 16-byte packed state plus a result word -- what crosses PCIe is what bounds that caller.
//   word 0: classical squares 0..7, 4 bits each: owning move index + 1, 0 = free
//   word 1: square 8 (bits 0-3) | slots 0..3, 6 bits each from bit 4 | terminated 28 | line 29 | illegal 30
//   word 2: slots 4..8, 6 bits each
// A slot holds the action index (mcts.py:339-349) of an UNCOLLAPSED move, 63 otherwise: exactly the
// moves env.py lists in q_states_p1 (even slots) / q_states_p2 (odd slots); turn = len(moves) % 2
// with len(moves) = classical squares + uncollapsed moves.
__device__ __forceinline__ uint32_t live_move_code(uint32_t E, uint32_t free_sq) {
    if ((E & free_sq) == 0u) return 63u;
    const uint32_t a = (uint32_t)ctz32(E), b = (uint32_t)flo32(E);
    return (15u * a - a * a + 2u * b - 2u) >> 1;                          // move2ind, mcts.py:345-349
}
__global__ void __launch_bounds__(kThreads)
k_step_packed_obs12(qttt_state* __restrict__ state, const uint8_t* __restrict__ action_coin,
                    uint32_t* __restrict__ obs12, uint32_t n, int iters) {
    __shared__ __align__(16) uint8_t smem[kLutQevalBytes];
    stage_luts(smem, kLutQevalBytes);
    const Luts L = luts_from_image(smem);
    const uint32_t first = (blockIdx.x * (uint32_t)iters) * kThreads + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const uint32_t i = first + (uint32_t)it * kThreads;
        if (i >= n) break;
        uint4* sp = reinterpret_cast<uint4*>(state + i);
        const uint4 sv = *sp;
        State s{sv.x, sv.y, sv.z, sv.w};
        const uint32_t ac = action_coin[i];
        const StepResult r = step_core(s, (uint32_t)L.pair[ac & 63u], ac >> 7, L);
        if (!r.illegal) *sp = make_uint4(s.x, s.y, s.z, s.w);
        const uint32_t win = any_line(s, r.classical, L) != 0u;
        const uint32_t term = win | (uint32_t)(r.n > 8u);
        const uint64_t nib = board_nibbles(s, L);
        const uint32_t free_sq = ~r.classical & M9;
        uint32_t w1 = (uint32_t)(nib >> 32) & 15u, w2 = 0u;
        w1 |= live_move_code(edge<0>(s), free_sq) << 4;
        w1 |= live_move_code(edge<1>(s), free_sq) << 10;
        w1 |= live_move_code(edge<2>(s), free_sq) << 16;
        w1 |= live_move_code(edge<3>(s), free_sq) << 22;
        w1 |= (term << 28) | (win << 29) | (r.illegal << 30);
        w2 |= live_move_code(edge<4>(s), free_sq);
        w2 |= live_move_code(edge<5>(s), free_sq) << 6;
        w2 |= live_move_code(edge<6>(s), free_sq) << 12;
        w2 |= live_move_code(edge<7>(s), free_sq) << 18;
        w2 |= live_move_code(edge<8>(s), free_sq) << 24;
        uint32_t* o = obs12 + 3ull * i;
        o[0] = (uint32_t)nib;
        o[1] = w1;
        o[2] = w2;
    }
}

// kPack12: the result words carry 12 bits each (free-square set, terminated, line, illegal), so
// four games' results leave as THREE 16-bit words -- word k of a group holds game k's result in
// its low 12 bits and nibble k of game 3's result on top: 1.5 bytes per game cross the link.
template <bool kPack12>
__global__ void __launch_bounds__(kThreads)
k_step_packed_zc(qttt_state* __restrict__ state, const uint8_t* __restrict__ action_coin_host,
                 uint16_t* __restrict__ result_host, qttt_state* __restrict__ obs_host, uint32_t n, int iters) {
    // A block owns `iters` (<= kZcMaxIters) consecutive chunks: their input bytes are one contiguous
    // run in host memory and their result words another.  The run is fetched across PCIe with ONE
    // round of 16-byte loads before any game is stepped, and the results leave in one burst of
    // 16-byte stores at the end -- two link round trips per block instead of two per chunk.
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    __shared__ __align__(16) uint8_t sh_in[kThreads * kZcMaxIters];
    __shared__ __align__(16) uint16_t sh_out[kThreads * kZcMaxIters];
    __shared__ __align__(16) uint16_t sh_pk[kPack12 ? kThreads * kZcMaxIters * 3 / 4 : 8];
    const uint32_t t = threadIdx.x;
    const uint32_t base0 = blockIdx.x * (uint32_t)iters * kThreads;
    if (base0 >= n) return;
    const uint32_t span = n - base0 < (uint32_t)iters * kThreads ? n - base0 : (uint32_t)iters * kThreads;
    uint16_t* out_host = kPack12 ? result_host + (base0 / 4u) * 3u : result_host + base0;
    const bool wide = span == (uint32_t)iters * kThreads &&
                      ((reinterpret_cast<uintptr_t>(action_coin_host + base0) |
                        reinterpret_cast<uintptr_t>(out_host)) & 15u) == 0u;
    if (wide) {
        for (uint32_t v = t; v < span / 16u; v += kThreads)
            reinterpret_cast<uint4*>(sh_in)[v] = reinterpret_cast<const uint4*>(action_coin_host + base0)[v];
    } else {
        for (uint32_t v = t; v < span; v += kThreads) sh_in[v] = action_coin_host[base0 + v];
    }
    stage_luts(smem, kLutStepBytes);          // (has the block-wide barrier that also covers sh_in)
    const Luts L = luts_from_image(smem);
    for (int it = 0; it < iters; ++it) {
        const uint32_t off = (uint32_t)it * kThreads + t;
        if (off >= span) break;
        const uint32_t i = base0 + off;
        uint4* sp = reinterpret_cast<uint4*>(state + i);
        const uint4 sv = *sp;
        State s{sv.x, sv.y, sv.z, sv.w};
        const uint32_t ac = sh_in[off];
        const StepResult r = step_core(s, (uint32_t)L.pair[ac & 63u], ac >> 7, L);
        const uint4 out = make_uint4(s.x, s.y, s.z, s.w);
        if (!r.illegal) *sp = out;
        if (obs_host) *reinterpret_cast<uint4*>(obs_host + i) = out;
        const uint32_t win = any_line(s, r.classical, L) != 0u;
        const uint32_t term = win | (uint32_t)(r.n > 8u);
        sh_out[off] = (uint16_t)((~r.classical & M9) | (term << 9) | (win << 10) | (r.illegal << 11));
    }
    __syncthreads();
    if (kPack12) {
        const uint32_t groups = (span + 3u) / 4u;
        for (uint32_t g = t; g < groups; g += kThreads) {
            uint32_t r[4];
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) r[k] = 4u * g + k < span ? (uint32_t)sh_out[4u * g + k] : 0u;
            sh_pk[3u * g + 0u] = (uint16_t)(r[0] | ((r[3] & 15u) << 12));
            sh_pk[3u * g + 1u] = (uint16_t)(r[1] | (((r[3] >> 4) & 15u) << 12));
            sh_pk[3u * g + 2u] = (uint16_t)(r[2] | (((r[3] >> 8) & 15u) << 12));
        }
        __syncthreads();
        const uint32_t words = 3u * groups;
        if (wide) {
            for (uint32_t v = t; v < words / 8u; v += kThreads)
                reinterpret_cast<uint4*>(out_host)[v] = reinterpret_cast<const uint4*>(sh_pk)[v];
        } else {
            for (uint32_t v = t; v < words; v += kThreads) out_host[v] = sh_pk[v];
        }
        return;
    }
    if (wide) {
        for (uint32_t v = t; v < span / 8u; v += kThreads)
            reinterpret_cast<uint4*>(out_host)[v] = reinterpret_cast<const uint4*>(sh_out)[v];
    } else {
        for (uint32_t v = t; v < span; v += kThreads) out_host[v] = sh_out[v];
    }
}

// ------------------------------------------------------------------------------ observe / pack
// Copies one block's worth of a staged output array (kBytesPerGame bytes per game, `valid`
// games) from shared memory to its place in the global array, as 16-byte vectors when the
// destination allows it.
template <int kBytesPerGame>
__device__ __forceinline__ void copy_out(const uint8_t* sm, void* dst, int64_t block_start, int valid) {
    if (!dst) return;
    uint8_t* g = static_cast<uint8_t*>(dst) + block_start * kBytesPerGame;
    if (valid == kThreads && (reinterpret_cast<uintptr_t>(g) & 15u) == 0u) {
        const uint4* s16 = reinterpret_cast<const uint4*>(sm);
        uint4* g16 = reinterpret_cast<uint4*>(g);
        for (int v = threadIdx.x; v < kThreads * kBytesPerGame / 16; v += kThreads) g16[v] = s16[v];
    } else {
        for (int v = threadIdx.x; v < valid * kBytesPerGame; v += kThreads) g[v] = sm[v];
    }
}

// Env._observation & co. (env.py:68-85, mcts.py:52-65,87-91) for one game.  The kernel moves a
// few odd-sized rows per game through shared memory, so shared-memory wavefronts are its scarce
// resource (ncu: the LSU data pipe at 90 % with table-driven decoding): the decode below is
// arithmetic, the only tables are two tiny selector tables.
//   classical      the four 8-bit planes are transposed into eight 4-bit values with two delta
//                  swaps, widened to bytes, and 1 is subtracted bytewise (free -> -1)
//   moves          per slot a 16-bit code a | b << 8 from two find-first-set instructions
//                  (both give -1 on an empty slot; the autofill entry has a == b)
//   q_states_p1/2  the codes of the uncollapsed slots of each player, compacted by PRMT with
//                  selectors looked up from the 5-bit / 4-bit "which slots are live" mask
// Outputs whose per-game size keeps them aligned (q_states_p2, rounds, reward, the one-byte ones)
// go straight to global memory, coalesced; the odd-sized rows (9, 18, 10, 36 bytes) are staged in
// shared memory and written out with 16-byte stores.  A slot at or beyond len(moves) is empty in
// every state this library produces, so len(moves) is not consulted per slot.  (The host emulation
// keeps the plain form, observe_game; the GPU tests diff this one against the oracle on every
// output after every ply.)
struct ObsOut {
    int8_t* classical; int8_t* moves; uint8_t* nmoves; int8_t* q1; int8_t* q2; uint8_t* turn;
    int8_t* rounds; float* reward_p1; uint8_t* winner; uint8_t* mask_bool;
};
// Which outputs a launch writes, known at compile time for the two common requests so that no
// pointer is tested per game: the env.py observation (classical, q_states_p1/2, turn), everything,
// or (kObsAny) whatever is non-null.
enum { kObsAny = 0, kObsEnv = 1, kObsAll = 2 };
template <int kSet> __device__ __forceinline__ bool obs_has(const void* p, bool in_env_set) {
    return kSet == kObsAny ? p != nullptr : (kSet == kObsAll || in_env_set);
}
__device__ __forceinline__ uint32_t bfind32(uint32_t v) {      // index of the highest set bit, 0xFFFFFFFF for 0
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {   // PRMT, selector used as is
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// bytes (a, b, ?, ?) of one move slot: a = lowest, b = highest square of E, both -1 when E == 0
__device__ __forceinline__ uint32_t move_code(uint32_t E) {
    return prmt(bfind32(E & (0u - E)), bfind32(E), 0x0040u);
}
// m += bit when the slot still has a free square, i.e. the move is uncollapsed (env.py:73-78):
// one LOP3 with a predicate output and one predicated add
__device__ __forceinline__ void add_if_live(uint32_t& m, uint32_t E, uint32_t free_sq, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 h;\n\t"
        "and.b32 h, %1, %2;\n\t"
        "setp.ne.u32 p, h, 0;\n\t"
        "@p add.u32 %0, %0, %3;\n\t}"
        : "+r"(m) : "r"(E), "r"(free_sq), "r"(bit));
}
template <int kSet>
__device__ __forceinline__ void observe_row(const State& s, const Luts& L, const ObsLutImage* O, const ObsOut& o,
                                            int64_t i, int t, uint8_t* st_cl, uint8_t* st_moves,
                                            uint8_t* st_q1, uint8_t* st_mask) {
    const uint32_t P3 = plane3(s);
    const uint32_t C = plane0(s) | plane1(s) | plane2(s) | P3;
    const uint32_t nm = n_moves(s);
    if (obs_has<kSet>(o.nmoves, false)) o.nmoves[i] = (uint8_t)nm;
    if (obs_has<kSet>(o.turn, true)) o.turn[i] = (uint8_t)(nm & 1u);          // env.py:83
    if (obs_has<kSet>(o.classical, true)) {
        // M: byte r = squares 0..7 of plane r.  Swapping index bits (0 <-> 3) and (1 <-> 4) leaves
        // square k in the low nibble of byte k (k < 4) and square k + 4 in its high nibble.
        uint32_t M = prmt(prmt(s.w, s.w >> 9, 0x0040u), prmt(s.w >> 18, P3, 0x0040u), 0x5410u);
        uint32_t d = ((M >> 7) ^ M) & 0x00AA00AAu;
        M ^= d ^ (d << 7);
        d = ((M >> 14) ^ M) & 0x0000CCCCu;
        M ^= d ^ (d << 14);
        // value = move index + 1, 0 = free; byte = value - 1: add 0x7F and flip the top bit (no carries)
        const uint32_t w0 = ((M & 0x0F0F0F0Fu) + 0x7F7F7F7Fu) ^ 0x80808080u;
        const uint32_t w1 = (((M >> 4) & 0x0F0F0F0Fu) + 0x7F7F7F7Fu) ^ 0x80808080u;
        const uint32_t v8 = ((s.w >> 8) & 1u) | ((s.w >> 16) & 2u) | ((s.w >> 24) & 4u) | ((P3 >> 5) & 8u);
        uint8_t* r = st_cl + 9 * t;
        r[0] = (uint8_t)w0; r[1] = (uint8_t)(w0 >> 8); r[2] = (uint8_t)(w0 >> 16); r[3] = (uint8_t)(w0 >> 24);
        r[4] = (uint8_t)w1; r[5] = (uint8_t)(w1 >> 8); r[6] = (uint8_t)(w1 >> 16); r[7] = (uint8_t)(w1 >> 24);
        r[8] = (uint8_t)(v8 - 1u);
    }
    const bool has_moves = obs_has<kSet>(o.moves, false), has_q1 = obs_has<kSet>(o.q1, true);
    const bool has_q2 = obs_has<kSet>(o.q2, true);
    if (has_moves || has_q1 || has_q2) {
        const uint32_t free_sq = ~C;
        uint32_t k[9], m1 = 0u, m2 = 0u;
#define QTTT_OBS_SLOT(T)                                                                 \
        {                                                                                \
            const uint32_t E = slot<T>(s.x, s.y, s.z);                                   \
            k[T] = move_code(E);                                                         \
            add_if_live(((T) & 1) ? m2 : m1, E, free_sq, 1u << ((T) >> 1));              \
        }
        QTTT_OBS_SLOT(0) QTTT_OBS_SLOT(1) QTTT_OBS_SLOT(2) QTTT_OBS_SLOT(3) QTTT_OBS_SLOT(4)
        QTTT_OBS_SLOT(5) QTTT_OBS_SLOT(6) QTTT_OBS_SLOT(7) QTTT_OBS_SLOT(8)
#undef QTTT_OBS_SLOT
        if (has_moves) {
            uint16_t* r = reinterpret_cast<uint16_t*>(st_moves + 18 * t);
#pragma unroll
            for (int T = 0; T < 9; ++T) r[T] = (uint16_t)k[T];
        }
        if (has_q1) {
            const uint32_t S0 = prmt(k[0], k[2], 0x5410u), S1 = prmt(k[4], k[6], 0x5410u);
            const uint32_t S2 = prmt(k[8], 0xFFFFFFFFu, 0x7610u);               // the last code, then the padding
            const uint4 sel = *reinterpret_cast<const uint4*>(O->q1sel[m1]);
            const uint32_t w0 = prmt(prmt(S0, S1, sel.x), S2, sel.y);
            const uint32_t w1 = prmt(prmt(S0, S1, sel.z), S2, sel.w);
            uint16_t* r = reinterpret_cast<uint16_t*>(st_q1 + 10 * t);
            r[0] = (uint16_t)w0; r[1] = (uint16_t)(w0 >> 16);
            r[2] = (uint16_t)w1; r[3] = (uint16_t)(w1 >> 16);
            r[4] = (uint16_t)(m1 == 31u ? k[8] : 0xFFFFu);
        }
        if (has_q2) {
            const uint32_t S0 = prmt(k[1], k[3], 0x5410u), S1 = prmt(k[5], k[7], 0x5410u);
            const uint4 sel = *reinterpret_cast<const uint4*>(O->q2sel[m2]);
            reinterpret_cast<uint2*>(o.q2)[i] =
                make_uint2(prmt(S0, S1, sel.x) | sel.z, prmt(S0, S1, sel.y) | sel.w);
        }
    }
    const bool has_rounds = obs_has<kSet>(o.rounds, false), has_reward = obs_has<kSet>(o.reward_p1, false);
    const bool has_winner = obs_has<kSet>(o.winner, false);
    if (has_rounds || has_reward || has_winner) {
        int px, po;
        win_rounds(s, L, px, po);
        if (has_rounds) reinterpret_cast<uint16_t*>(o.rounds)[i] = (uint16_t)((uint32_t)(px & 255) | ((uint32_t)(po & 255) << 8));
        if (has_reward) {                                                       // env.py:87-112
            const int a = px < 0 ? 10 : px, b = po < 0 ? 10 : po;
            o.reward_p1[i] = a < b ? 1.0f : (b < a ? -1.0f : 0.0f);
        }
        if (has_winner) o.winner[i] = (uint8_t)winner_of(px, po);               // mcts.py:52-65
    }
    if (obs_has<kSet>(o.mask_bool, false)) {                                    // mcts.py:87-91
        const uint64_t lm = L.legal[~C & M9];
        uint32_t* w = reinterpret_cast<uint32_t*>(st_mask + 36 * t);
#pragma unroll
        for (int q = 0; q < 9; ++q)       // 4 mask bits -> 4 bool bytes: (bits * 0x204081) & 0x01010101
            w[q] = (((uint32_t)(lm >> (4 * q)) & 15u) * 0x00204081u) & 0x01010101u;
    }
}

// A full chunk of 256 staged rows to global memory, 16 bytes per store (the destination of a
// full chunk is 16-byte aligned when the array is: 256 * kBytesPerGame is a multiple of 16).
template <int kBytesPerGame>
__device__ __forceinline__ void copy_out_full(const uint8_t* sm, void* dst, int64_t block_start) {
    const uint4* s16 = reinterpret_cast<const uint4*>(sm);
    uint4* g16 = reinterpret_cast<uint4*>(static_cast<uint8_t*>(dst) + block_start * kBytesPerGame);
#pragma unroll
    for (int v = 0; v < (kBytesPerGame + 15) / 16; ++v) {
        const int at = v * kThreads + (int)threadIdx.x;
        if (kBytesPerGame % 16 == 0 || at < kThreads * kBytesPerGame / 16) g16[at] = s16[at];
    }
}

template <int kSet>
__global__ void __launch_bounds__(kThreads)
k_observe(const qttt_state* __restrict__ state, const ObsOut o, int64_t n, bool aligned16) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    __shared__ __align__(16) uint8_t smem_obs[kObsLutBytes];
    __shared__ __align__(16) uint8_t st_classical[kThreads * 9];
    __shared__ __align__(16) uint8_t st_moves[kSet == kObsEnv ? 16 : kThreads * 18];
    __shared__ __align__(16) uint8_t st_q1[kThreads * 10];
    __shared__ __align__(16) uint8_t st_mask[kSet == kObsEnv ? 16 : kThreads * 36];
    stage_luts(smem, kLutStepBytes, smem_obs, &g_obs_lut, kObsLutBytes);
    const Luts L = luts_from_image(smem);
    const ObsLutImage* O = reinterpret_cast<const ObsLutImage*>(smem_obs);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t block_start = (int64_t)blockIdx.x * kThreads; block_start < n; block_start += stride) {
        const int valid = (int)((n - block_start) < kThreads ? (n - block_start) : kThreads);
        const int t = threadIdx.x;
        if (t < valid)
            observe_row<kSet>(load_state(state, block_start + t), L, O, o, block_start + t, t,
                              st_classical, st_moves, st_q1, st_mask);
        __syncthreads();
        if (valid == kThreads && aligned16) {
            if (obs_has<kSet>(o.classical, true)) copy_out_full<9>(st_classical, o.classical, block_start);
            if (obs_has<kSet>(o.moves, false)) copy_out_full<18>(st_moves, o.moves, block_start);
            if (obs_has<kSet>(o.q1, true)) copy_out_full<10>(st_q1, o.q1, block_start);
            if (obs_has<kSet>(o.mask_bool, false)) copy_out_full<36>(st_mask, o.mask_bool, block_start);
        } else {
            copy_out<9>(st_classical, o.classical, block_start, valid);
            if (kSet != kObsEnv) copy_out<18>(st_moves, o.moves, block_start, valid);
            copy_out<10>(st_q1, o.q1, block_start, valid);
            if (kSet != kObsEnv) copy_out<36>(st_mask, o.mask_bool, block_start, valid);
        }
        __syncthreads();
    }
}

// Env.step returning the observation (env.py:34-53 with env.py:68-85) as ONE launch: K1 followed
// by the env.py observation of the state it just produced, which is still in registers -- the
// separate observe launch would read the 16-byte state back and pay a second launch.
struct StepObsArgs {
    StepArgs st;
    ObsOut obs;            // classical, q1, q2, turn (all required); the rest unused
    bool aligned16;        // classical and q1 start on 16-byte boundaries
};
template <int kFmt, int kMode>
__global__ void __launch_bounds__(kThreads, 8) k_step_obs(const StepObsArgs fa) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    __shared__ __align__(16) uint8_t smem_obs[kObsLutBytes];
    __shared__ __align__(16) uint8_t st_classical[kThreads * 9];
    __shared__ __align__(16) uint8_t st_q1[kThreads * 10];
    stage_luts(smem, kLutStepBytes, smem_obs, &g_obs_lut, kObsLutBytes);
    const Luts L = luts_from_image(smem);
    const ObsLutImage* O = reinterpret_cast<const ObsLutImage*>(smem_obs);
    const StepArgs& a = fa.st;
    const uint32_t first_chunk = blockIdx.x * (uint32_t)a.iters;
    for (int it = 0; it < a.iters; ++it) {
        const uint32_t chunk_start = (first_chunk + (uint32_t)it) * kThreads;
        if (chunk_start >= a.n) break;                              // the same for the whole block
        const uint32_t i = chunk_start + threadIdx.x;
        const bool valid = i < a.n;
        // lanes past the end step an empty game with an illegal action (full-warp votes inside)
        State s = empty_state();
        uint32_t act = 255u, coin = 0u;
        if (valid) {
            if (kMode != kStepFresh) s = load_state(a.state, i);
            if (kFmt == QTTT_ACT_INDEX) {
                act = a.action[i];
            } else {
                const uchar2 ab = reinterpret_cast<const uchar2*>(a.action)[i];
                act = (uint32_t)ab.x | ((uint32_t)ab.y << 8);
            }
            if (a.coin) coin = a.coin[i];
        }
        const uint32_t enew = kFmt == QTTT_ACT_INDEX ? (uint32_t)L.pair[act] : pair_to_edge(act & 255u, act >> 8);
        const StepOut o = step_game<false, kMode, true>(s, enew, a.coin != nullptr, coin & 1u, a.seed,
                                                        a.game_base + (uint64_t)i, a.dword, L);
        if (valid) {
            if (o.write_state) store_state(a.state, i, s);
            if (a.reward) reinterpret_cast<uint32_t*>(a.reward)[i] = reward_bits(o.win);   // env.py:49
            if (a.done) a.done[i] = (uint8_t)o.done;                                       // env.py:51
            if (a.mask) a.mask[i] = L.legal[~o.classical & M9];                            // mcts.py:87-91
            if (a.status) a.status[i] = (uint8_t)o.status;
            observe_row<kObsEnv>(s, L, O, fa.obs, (int64_t)i, (int)threadIdx.x, st_classical, nullptr, st_q1, nullptr);
        }
        __syncthreads();
        const uint32_t left = a.n - chunk_start;
        if (left >= (uint32_t)kThreads && fa.aligned16) {
            copy_out_full<9>(st_classical, fa.obs.classical, chunk_start);
            copy_out_full<10>(st_q1, fa.obs.q1, chunk_start);
        } else {
            const int nvalid = left < (uint32_t)kThreads ? (int)left : kThreads;
            copy_out<9>(st_classical, fa.obs.classical, chunk_start, nvalid);
            copy_out<10>(st_q1, fa.obs.q1, chunk_start, nvalid);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads)
k_pack(qttt_state* __restrict__ state, const int8_t* __restrict__ classical_in,
       const int8_t* __restrict__ moves, const uint8_t* __restrict__ nmoves, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
        store_state(state, i, pack_game(classical_in, moves, nmoves, i));
}

// ------------------------------------------------------------------------------ features
// GameState.to_vector for n games -> float[n][18][10].  720 B per game are written and only
// ~30 of the 180 values are non-zero, so: every lane zero-fills its own game's slot in shared
// memory, scatters the non-zeros (one-hot of the board, 1/sqrt(9) marks of every move, the
// "no live mark" column), and the warp then streams its 32 games (23 KB, contiguous in the
// output) to HBM as lane-contiguous float4 vectors.
constexpr int kFeatThreads = 64;            // 2 warps x 32 games x 188 floats = 48,128 B of shared memory
constexpr int kFeatStride = 188;            // floats per game slot: 180 used; 188 = 4 (mod 32) x 7 keeps float4 accesses of 8 lanes on distinct banks

template <int T>
__device__ __forceinline__ void feature_move(float* slot, const State& s, uint32_t nm) {
    if ((uint32_t)T < nm) {
        const uint32_t E = edge<T>(s);
        if (E) {
            slot[(9 + ctz32(E)) * 10 + T] = 1.0f / 3.0f;
            slot[(9 + flo32(E)) * 10 + T] = 1.0f / 3.0f;    // the same cell for the autofill entry (s, s, t)
        }
    }
}

// One warp's 32 games -> features: every lane zero-fills its own slot of `warp_buf` and scatters
// its game's non-zeros, then the warp streams the `games_here` slots (contiguous in the output) to
// HBM as lane-contiguous float4 vectors.  `mine_valid`: this lane holds a game.
__device__ __forceinline__ void emit_features(float* warp_buf, int lane, const State& s, bool mine_valid,
                                              float* __restrict__ out_base, int games_here) {
    float* mine = warp_buf + lane * kFeatStride;
#pragma unroll
    for (int k = 0; k < 45; ++k) reinterpret_cast<float4*>(mine)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mine_valid) {
        const uint32_t P0 = plane0(s), P1 = plane1(s), P2 = plane2(s), P3 = plane3(s);
        const uint32_t C = P0 | P1 | P2 | P3, nm = n_moves(s);
#pragma unroll
        for (int sq = 0; sq < 9; ++sq) {                       // rows 0..8: one-hot of board[sq]
            const int b = board_value(P0, P1, P2, P3, sq);
            mine[sq * 10 + (b < 0 ? 9 : b)] = 1.0f;
        }
        feature_move<0>(mine, s, nm); feature_move<1>(mine, s, nm); feature_move<2>(mine, s, nm);
        feature_move<3>(mine, s, nm); feature_move<4>(mine, s, nm); feature_move<5>(mine, s, nm);
        feature_move<6>(mine, s, nm); feature_move<7>(mine, s, nm); feature_move<8>(mine, s, nm);
        uint32_t live = 0u;                                    // squares with an uncollapsed mark
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const uint32_t E = edge_dyn(s, (uint32_t)t);
            live |= ((uint32_t)t < nm && !(E & C)) ? E : 0u;
        }
#pragma unroll
        for (int sq = 0; sq < 9; ++sq)
            if (!(live >> sq & 1u)) mine[(9 + sq) * 10 + 9] = 1.0f;
    }
    __syncwarp();
    float4* dst = reinterpret_cast<float4*>(out_base);
    for (int q = lane; q < games_here * 45; q += 32) {
        const int owner = q / 45, r = q - owner * 45;
        dst[q] = reinterpret_cast<const float4*>(warp_buf + owner * kFeatStride)[r];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kFeatThreads)
k_features(const qttt_state* __restrict__ state, float* __restrict__ out, int64_t n) {
    __shared__ __align__(16) float buf[kFeatThreads / 32][32 * kFeatStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * kFeatThreads + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * kFeatThreads) >> 5;
    for (int64_t base = warp0 * 32; base < n; base += n_warps * 32) {
        const int64_t g = base + lane;
        const int games_here = (int)((n - base) < 32 ? (n - base) : 32);
        State s = empty_state();
        if (g < n) s = load_state(state, g);
        emit_features(buf[warp], lane, s, g < n, out + base * 180, games_here);
    }
}

// nn.Model.get_mask's illegal-action mask as 36 bool bytes (9 words) of one game: the complement
// of the legal mask (nn.py:44-61: occupied[i] or occupied[j]).
__device__ __forceinline__ void store_illegal_mask(uint8_t* __restrict__ dst, uint64_t legal) {
    uint32_t* w = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
    for (int k = 0; k < 9; ++k)
        w[k] = ((((uint32_t)(legal >> (4 * k)) & 15u) * 0x00204081u) & 0x01010101u) ^ 0x01010101u;
}

// Env.step fused with the policy/value net's input encoding: the step of k_step followed, in the
// same thread, by GameState.to_vector of the NEW state (and optionally nn.Model.get_mask), so the
// features cost no second pass over the state array.  Same block shape and staging as k_features
// (the 720 B of features per game are the traffic; the step itself is 48 B).
struct StepFeatArgs {
    StepArgs st;
    float* features;            // [n][18][10], 16-byte aligned
    uint8_t* illegal_mask;      // [n][36] bool, optional, 4-byte aligned
};

template <int kFmt, int kMode>
__global__ void __launch_bounds__(kFeatThreads) k_step_features(const StepFeatArgs fa) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    uint8_t* lut = dyn_smem;
    float* buf = reinterpret_cast<float*>(dyn_smem + kLutStepBytes);
    stage_luts(lut, kLutStepBytes);
    const Luts L = luts_from_image(lut);
    const StepArgs& a = fa.st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* warp_buf = buf + warp * 32 * kFeatStride;
    const uint32_t warp0 = (blockIdx.x * kFeatThreads + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * kFeatThreads) >> 5;
    for (uint32_t base = warp0 * 32u; base < a.n; base += n_warps * 32u) {
        const uint32_t i = base + lane;
        const bool valid = i < a.n;
        const int games_here = (int)((a.n - base) < 32u ? (a.n - base) : 32u);
        State s = empty_state();
        uint64_t legal_of_mine = 0ull;
        if (valid) {
            if (kMode != kStepFresh) {
                const uint4 sv = *reinterpret_cast<const uint4*>(a.state + i);
                s = State{sv.x, sv.y, sv.z, sv.w};
            }
            uint32_t enew;
            if (kFmt == QTTT_ACT_INDEX) {
                enew = L.pair[a.action[i]];
            } else {
                const uchar2 ab = reinterpret_cast<const uchar2*>(a.action)[i];
                enew = pair_to_edge(ab.x, ab.y);
            }
            const uint32_t coin = a.coin ? (a.coin[i] & 1u) : 0u;
            const StepOut o = step_game<false, kMode>(s, enew, a.coin != nullptr, coin, a.seed,
                                                      a.game_base + (uint64_t)i, a.dword, L);
            if (o.write_state) *reinterpret_cast<uint4*>(a.state + i) = make_uint4(s.x, s.y, s.z, s.w);
            const uint64_t legal = L.legal[~o.classical & M9];
            if (a.reward) reinterpret_cast<uint32_t*>(a.reward)[i] = reward_bits(o.win);
            if (a.done) a.done[i] = (uint8_t)o.done;
            if (a.mask) a.mask[i] = legal;
            if (a.status) a.status[i] = (uint8_t)o.status;
            legal_of_mine = legal;
        }
        emit_features(warp_buf, lane, s, valid, fa.features + (size_t)base * 180, games_here);
        if (fa.illegal_mask) {
            // the warp's 32 masks (1,152 contiguous bytes of the output) go through the feature
            // buffer, which is free again, and out as lane-contiguous words
            uint8_t* mbuf = reinterpret_cast<uint8_t*>(warp_buf);
            store_illegal_mask(mbuf + 36 * lane, legal_of_mine);
            __syncwarp();
            uint32_t* dst = reinterpret_cast<uint32_t*>(fa.illegal_mask + 36ull * base);
            for (int q = lane; q < games_here * 9; q += 32) dst[q] = reinterpret_cast<const uint32_t*>(mbuf)[q];
            __syncwarp();
        }
    }
}

// nn.Model.get_mask for packed states.  36 bytes per game at a 36-byte stride: each block decodes
// its 256 games into shared memory and writes the 9,216 bytes out as coalesced 16-byte stores
// (per-thread word stores at that stride ran at 18 % of the DRAM peak: 569 us per 2^24 games).
__global__ void __launch_bounds__(kThreads)
k_get_mask(const qttt_state* __restrict__ state, uint8_t* __restrict__ illegal_mask, int64_t n) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    __shared__ __align__(16) uint8_t st_mask[kThreads * 36];
    stage_luts(smem, kLutStepBytes);
    const Luts L = luts_from_image(smem);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t block_start = (int64_t)blockIdx.x * kThreads; block_start < n; block_start += stride) {
        const int valid = (int)((n - block_start) < kThreads ? (n - block_start) : kThreads);
        const int t = threadIdx.x;
        if (t < valid) {
            const State s = load_state(state, block_start + t);
            store_illegal_mask(st_mask + 36 * t, L.legal[~classical(s) & M9]);
        }
        __syncthreads();
        copy_out<36>(st_mask, illegal_mask, block_start, valid);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ single env
// qtttgym.Env for ONE game with the lowest possible latency: one launch steps the game (the
// action travels as kernel arguments: no host->device copy), decodes everything Env.step /
// observ / turn / _reward report into a 128-byte record and writes that record straight into
// MAPPED PINNED HOST memory; the host spins on the sequence word at its end (written last, after
// a system-scope fence).  No copy engine, no stream synchronisation.
//   record: state[0:16] mask[16:24] reward[24:28] done[28] status[29] turn[30] n_moves[31]
//           classical[32:41] q_p1[48:58] q_p2[58:66] rounds[66:68] winner[68] reward_p1[72:76]
//           moves[80:98] ... seq[124:128]
__global__ void __launch_bounds__(32)
k_env1(qttt_state* __restrict__ state, int a, int b, int coin, int op, uint64_t seed, uint32_t dword,
       uint8_t* __restrict__ rec_host, uint32_t seq) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    __shared__ __align__(16) uint8_t rec[128];
    stage_luts(smem, kLutStepBytes);
    const Luts L = luts_from_image(smem);
    if (threadIdx.x < 8) reinterpret_cast<uint4*>(rec)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (threadIdx.x == 0) {
        State s = empty_state();
        if (op != 1) s = load_state(state, 0);                   // op 1: reset (Board.__init__)
        StepOut o;
        if (op == 0) {                                            // op 0: Env.step
            // anything outside 0..8 (negative values included: outside the action domain) is illegal
            const uint32_t enew = pair_to_edge((uint32_t)a, (uint32_t)b);
            o = step_game<false, kStepPlain>(s, enew, coin >= 0, (uint32_t)coin & 1u, seed, 0ull, dword, L);
        } else {                                                  // reset / refresh: outputs of the state as it is
            const uint32_t C = classical(s);
            o.win = any_line(s, C, L);
            o.done = (o.win != 0u) | (n_moves(s) > 8u);
            o.classical = C;
            o.status = 0u;
        }
        store_state(state, 0, s);
        *reinterpret_cast<uint4*>(rec) = make_uint4(s.x, s.y, s.z, s.w);
        *reinterpret_cast<uint64_t*>(rec + 16) = L.legal[~o.classical & M9];
        *reinterpret_cast<uint32_t*>(rec + 24) = reward_bits(o.win);
        rec[28] = (uint8_t)o.done;
        rec[29] = (uint8_t)o.status;
        observe_game(s, L, reinterpret_cast<int8_t*>(rec + 32), reinterpret_cast<int8_t*>(rec + 80), rec + 31,
                     reinterpret_cast<int8_t*>(rec + 48), reinterpret_cast<int8_t*>(rec + 58), rec + 30,
                     reinterpret_cast<int8_t*>(rec + 66), reinterpret_cast<float*>(rec + 72), rec + 68, nullptr, 0);
    }
    __syncwarp();
    if (threadIdx.x < 7) reinterpret_cast<uint4*>(rec_host)[threadIdx.x] = reinterpret_cast<const uint4*>(rec)[threadIdx.x];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile uint32_t*>(rec_host + 124) = seq;
    }
}

// QEvalClassic.eval for ONE measurement at minimum latency (the plugin seam, board.py:51): the
// component's moves travel as a packed state in the kernel arguments, both outcomes' per-move
// squares come back through mapped pinned host memory (32-byte record: sq0[9], sq1[9], closes,
// ..., seq at byte 28), the host spins on the sequence word.
__global__ void __launch_bounds__(32)
k_qeval1(uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint32_t action, uint8_t* __restrict__ rec_host,
         uint32_t seq) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    __shared__ __align__(16) int8_t rec[32];
    stage_luts(smem, kLutStepBytes);
    const Luts L = luts_from_image(smem);
    if (threadIdx.x == 0) {
        const State s{x, y, z, w};
        uint8_t closes = 0;
        qeval_game<true>(s, action, L, nullptr, nullptr, nullptr, nullptr, rec, rec + 9, &closes, nullptr, 0);
        rec[18] = (int8_t)closes;
        reinterpret_cast<uint4*>(rec_host)[0] = reinterpret_cast<const uint4*>(rec)[0];
        reinterpret_cast<uint2*>(rec_host)[2] = reinterpret_cast<const uint2*>(rec)[2];
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(rec_host + 28) = seq;
    }
}

// ------------------------------------------------------------------------------ K3 qeval
template <bool kSquares>
__global__ void __launch_bounds__(kThreads, kSquares ? 4 : 8)
k_qeval_both(const qttt_state* __restrict__ state, const uint8_t* __restrict__ action,
             qttt_state* __restrict__ next0, qttt_state* __restrict__ next1,
             uint64_t* __restrict__ board0, uint64_t* __restrict__ board1,
             int8_t* __restrict__ sq0, int8_t* __restrict__ sq1, uint8_t* __restrict__ closes,
             float* __restrict__ result_prob, int64_t n) {
    __shared__ __align__(16) uint8_t smem[kLutQevalBytes];
    stage_luts(smem, kLutQevalBytes);
    const Luts L = luts_from_image(smem);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
        qeval_game<kSquares>(load_state(state, i), action[i], L, reinterpret_cast<State*>(next0),
                   reinterpret_cast<State*>(next1), board0, board1, sq0, sq1, closes, result_prob, i);
}

// The config-3 shape of K3 -- both outcome boards and the closes flag, nothing else: the sweep's
// plane accumulators go straight to the nibble boards (boards_both), the successor states are
// never assembled.
__global__ void __launch_bounds__(kThreads, 8)
k_qeval_boards(const qttt_state* __restrict__ state, const uint8_t* __restrict__ action,
               uint64_t* __restrict__ board0, uint64_t* __restrict__ board1, uint8_t* __restrict__ closes,
               uint32_t n, int iters) {
    __shared__ __align__(16) uint8_t smem[kLutQevalBytes];
    stage_luts(smem, kLutQevalBytes);
    const Luts L = luts_from_image(smem);
    // a block owns `iters` consecutive chunks of 256 boards (32-bit indices: the host splits larger batches)
    const uint32_t first = (blockIdx.x * (uint32_t)iters) * kThreads + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const uint32_t i = first + (uint32_t)it * kThreads;
        const bool valid = i < n;
        // out-of-range lanes run on an empty game so that the warp votes inside stay full
        const State s = valid ? load_state(state, i) : State{0u, 0u, 0u, 0u};
        const uint32_t act = valid ? (uint32_t)action[i] : 255u;
        const BoardsBoth r = boards_both<true>(s, (uint32_t)L.pair[act], L);
        if (valid) {
            board0[i] = r.board0;
            board1[i] = r.board1;
            closes[i] = (uint8_t)r.collapsed;
        }
    }
}

// ------------------------------------------------------------------------------ playouts
// K4: a block works on one root at a time (rollouts strided over its threads).
__global__ void __launch_bounds__(kThreads)
k_rollout(const qttt_state* __restrict__ roots, int64_t n_roots, int32_t n_rollouts, uint64_t seed,
          int32_t* __restrict__ tallies, float* __restrict__ value,
          unsigned long long* __restrict__ steps_total) {
    __shared__ __align__(16) uint8_t smem[kLutPolicyBytes];
    __shared__ int sh_tally[3];
    __shared__ unsigned long long sh_steps;
    stage_luts(smem, kLutPolicyBytes);
    const Luts L = luts_from_image(smem);
    unsigned long long block_steps = 0ull;
    // persistent blocks: the tables are staged once per block, roots are taken grid-stride
    for (int64_t root = blockIdx.x; root < n_roots; root += gridDim.x) {
        if (threadIdx.x < 3) sh_tally[threadIdx.x] = 0;
        if (threadIdx.x == 0) sh_steps = 0ull;
        __syncthreads();
        const State s0 = load_state(roots, root);
        int xw = 0, ow = 0, dr = 0;
        uint32_t steps = 0, cols = 0;
        for (int32_t j = threadIdx.x; j < n_rollouts; j += kThreads) {
            const uint64_t game = (uint64_t)root * (uint64_t)n_rollouts + (uint64_t)j;
            const uint32_t w = playout_game(s0, seed, game, 1u, L, steps, cols);
            xw += w == 1u; ow += w == 2u; dr += w == 0u;
        }
        xw = __reduce_add_sync(0xFFFFFFFFu, xw);
        ow = __reduce_add_sync(0xFFFFFFFFu, ow);
        dr = __reduce_add_sync(0xFFFFFFFFu, dr);
        steps = __reduce_add_sync(0xFFFFFFFFu, steps);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&sh_tally[0], xw); atomicAdd(&sh_tally[1], ow); atomicAdd(&sh_tally[2], dr);
            atomicAdd(&sh_steps, (unsigned long long)steps);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (tallies) { tallies[3 * root] = sh_tally[0]; tallies[3 * root + 1] = sh_tally[1]; tallies[3 * root + 2] = sh_tally[2]; }
            if (value) {
                // mcts.py:171,173: sum(r if leaf.turn else -r) / num_simulations
                const float r = (float)(sh_tally[0] - sh_tally[1]) / (float)n_rollouts;
                value[root] = (plies_of(s0) & 1u) ? -r : r;
            }
            block_steps += sh_steps;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && steps_total && block_steps) atomicAdd(steps_total, block_steps);
}

// K5: self-play sweep from the empty board, entirely in registers.  The 32 games of a warp are
// played in lock-step (ply p of all of them together), so len(moves) == p in every active lane:
// the nine plies are nine compile-time specialisations (playout_ply_fixed<p>: immediates instead
// of table rows, sweep<p> called directly, the even / odd halves of the shared Philox block
// resolved at compile time).  Lanes whose game ended earlier (mean 8.29 of 9 plies) idle until
// the warp's last game ends.
template <int PLY>
__device__ __forceinline__ void sweep_from_ply(State& s, uint32_t& C, bool& active, uint32_t& len, uint32_t& co,
                                               uint64_t seed, uint64_t g, const Luts& L, DrawCache& cache) {
    if constexpr (PLY < 9) {
        if (!__any_sync(0xFFFFFFFFu, active)) return;
        if (active) {
            const StepResult r = playout_ply_fixed<PLY>(s, C, seed, g, 0u, L, cache);
            C = r.classical;
            co += r.collapsed;
            len = PLY + 1;
            active = !((any_line(s, C, L) != 0u) | (r.n >= 9u));              // mcts.py:52-65
        }
        sweep_from_ply<PLY + 1>(s, C, active, len, co, seed, g, L, cache);
    }
}

__global__ void __launch_bounds__(kThreads, 6)      // 39 registers; 7 or 8 blocks per SM spill and measured 2 % slower
k_sweep(int64_t game_lo, int64_t game_hi, uint64_t seed, unsigned long long* __restrict__ stats) {
    __shared__ __align__(16) uint8_t smem[kLutPolicyBytes];
    __shared__ unsigned long long sh[16];
    stage_luts(smem, kLutPolicyBytes);
    const Luts L = luts_from_image(smem);
    if (threadIdx.x < 16) sh[threadIdx.x] = 0ull;
    __syncthreads();

    const int64_t stride = (int64_t)gridDim.x * kThreads;
    uint32_t xw = 0, ow = 0, dr = 0, st = 0, co = 0, games = 0;
    uint32_t h5 = 0, h6 = 0, h7 = 0, h8 = 0, h9 = 0;       // games by length (a game has 5..9 plies)
    for (int64_t g = game_lo + (int64_t)blockIdx.x * kThreads + threadIdx.x;; g += stride) {
        bool active = g < game_hi;
        if (!__any_sync(0xFFFFFFFFu, active)) break;
        State s = empty_state();
        uint32_t C = 0u, len = 0u;
        DrawCache cache = empty_draw_cache();
        const bool mine = active;
        games += active;
        sweep_from_ply<0>(s, C, active, len, co, seed, (uint64_t)g, L, cache);
        // every lane's game is over: who has the earlier line (mcts.py:52-65), once, undiverged
        bool t2;
        const uint32_t w = finished_winner(s, L, t2);
        xw += mine & (w == 1u); ow += mine & (w == 2u); dr += mine & (w == 0u);
        st += len;
        h5 += len == 5u; h6 += len == 6u; h7 += len == 7u; h8 += len == 8u; h9 += len == 9u;
    }
    uint32_t vals[11] = {xw, ow, dr, st, co, games, h5, h6, h7, h8, h9};
    const int slot[11] = {0, 1, 2, 3, 4, 5, 11, 12, 13, 14, 15};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, vals[k]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[slot[k]], (unsigned long long)v);
    }
    __syncthreads();
    if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&stats[threadIdx.x], sh[threadIdx.x]);
}

// ------------------------------------------------------------------------------ MCTS
__global__ void __launch_bounds__(kThreads)
k_mcts_init(MctsNode* __restrict__ pool, int64_t capacity, int32_t* __restrict__ meta,
            const qttt_state* __restrict__ roots, int64_t n_roots) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    stage_luts(smem, kLutStepBytes);
    const Luts L = luts_from_image(smem);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n_roots; r += stride)
        mcts_init_root(pool + r * capacity, meta + r * kMetaStride, load_state(roots, r), L);
}

// ---- warp-cooperative forms of mcts_uct_select / mcts_expand / mcts_select (qttt_mcts.cuh):
// the same arithmetic in the same order per action, spread over the 32 lanes of warp 0.
__device__ __forceinline__ int warp_uct_select(const MctsNode& nd, double c_puct, int lane) {
    const uint64_t legal = nd.legal;
    const int m = popc32((uint32_t)legal) + popc32((uint32_t)(legal >> 32));
    const double ps = d_mul(d_div(1.0, (double)m), d_sqrt((double)nd.ntot));
    int best = 64;
    double best_v = 0.0;
    for (int a = lane; a < 36; a += 32) {
        if (!(legal >> a & 1ull)) continue;
        const uint32_t na = nd.n[a];
        const double u = d_div(ps, (double)(1u + na));
        const double q = na ? d_div(nd.w[a], (double)na) : 0.0;
        const double v = d_add(q, d_mul(c_puct, u));
        if (best == 64 || v > best_v) { best = a; best_v = v; }
    }
    // first maximum in ascending action order == (greater value) or (equal value, smaller action)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_xor_sync(0xFFFFFFFFu, best_v, off);
        const int oa = __shfl_xor_sync(0xFFFFFFFFu, best, off);
        if (oa != 64 && (best == 64 || ov > best_v || (ov == best_v && oa < best))) { best = oa; best_v = ov; }
    }
    return best;
}

__device__ __forceinline__ void warp_init_node(MctsNode& nd, const State& s, bool turn, const Luts& L, int lane) {
    for (int a = lane; a < 36; a += 32) { nd.n[a] = 0u; nd.child[a][0] = -1; nd.child[a][1] = -1; nd.w[a] = 0.0; }
    if (lane == 0) {
        nd.state = s;
        nd.ntot = 0u;
        nd.has_p = 0;
        bool terminal;
        nd.winner = (uint8_t)finished_winner(s, L, terminal);
        nd.terminal = terminal ? 1 : 0;
        nd.turn = turn ? 1 : 0;
        nd.legal = L.legal[~classical(s) & M9];
    }
}

__device__ __forceinline__ bool warp_expand(MctsNode* tree, int32_t* meta, int64_t capacity, int node, int a,
                                            const Luts& L, int lane) {
    const uint32_t enew = L.pair[a];
    State s0 = tree[node].state, s1 = s0;                    // every lane computes the same children
    const StepResult r0 = step_core(s0, enew, 0u, L);
    const int need = r0.collapsed ? 2 : 1;
    if (mcts_nodes_available(meta, capacity) < need) {
        if (lane == 0) meta[kMetaError] |= kMctsErrPoolFull;
        return false;
    }
    // lane 0 takes the nodes (free list first, mcts_alloc); everybody initialises them
    int c0 = 0, c1 = -1;
    if (lane == 0) {
        c0 = mcts_alloc(tree, meta);
        if (r0.collapsed) c1 = mcts_alloc(tree, meta);
    }
    c0 = __shfl_sync(0xFFFFFFFFu, c0, 0);
    c1 = __shfl_sync(0xFFFFFFFFu, c1, 0);
    const bool turn = !tree[node].turn;
    warp_init_node(tree[c0], s0, turn, L, lane);
    if (r0.collapsed) {
        step_core(s1, enew, 1u, L);
        warp_init_node(tree[c1], s1, turn, L, lane);
    }
    __syncwarp();
    if (lane == 0) {
        tree[node].child[a][0] = c0;
        if (r0.collapsed) tree[node].child[a][1] = c1;
    }
    __syncwarp();
    return true;
}

// A rollout is: one warp walks the tree (PUCT select, expanding one (node, action) when needed),
// the leaf's num_simulations playouts run in parallel on the lanes, a reduction gives r_tot, one
// thread backs the value up.  Trees of different roots advance independently.
//   kWarpPerRoot = false: one block per root, its threads share the playouts (num_simulations > 32).
//   kWarpPerRoot = true:  num_simulations <= 32 -- a root needs one warp only, and a block carries
//       kMctsWarps of them on one staged copy of the tables.  (As one-warp blocks the 19 KB of tables
//       per block capped an SM at 11 trees; searching 32,768 roots ran at 116 M rollouts/s.)
constexpr int kMctsWarps = 8;
template <bool kWarpPerRoot>
__global__ void __launch_bounds__(kThreads)
k_mcts_run(MctsNode* __restrict__ pool, int64_t capacity, int32_t* __restrict__ meta, int64_t n_roots,
           int32_t n_rollouts, int32_t num_sims, double c_puct, uint64_t seed, uint64_t root_base) {
    __shared__ __align__(16) uint8_t smem[kLutPolicyBytes];
    __shared__ int sh_path_node_all[kWarpPerRoot ? kMctsWarps : 1][12], sh_path_act_all[kWarpPerRoot ? kMctsWarps : 1][12];
    __shared__ int sh_depth_all[kWarpPerRoot ? kMctsWarps : 1], sh_leaf_all[kWarpPerRoot ? kMctsWarps : 1];
    __shared__ int sh_rtot_all[kWarpPerRoot ? kMctsWarps : 1];
    stage_luts(smem, kLutPolicyBytes);
    const Luts L = luts_from_image(smem);
    const int lane = threadIdx.x & 31, warp = kWarpPerRoot ? (int)(threadIdx.x >> 5) : 0;
    const int64_t root = kWarpPerRoot ? (int64_t)blockIdx.x * (blockDim.x >> 5) + warp : (int64_t)blockIdx.x;
    if (root >= n_roots) return;                       // whole warps only (kWarpPerRoot); no block barrier below then
    int* sh_path_node = sh_path_node_all[warp];
    int* sh_path_act = sh_path_act_all[warp];
    int& sh_depth = sh_depth_all[warp];
    int& sh_leaf = sh_leaf_all[warp];
    int& sh_rtot = sh_rtot_all[warp];
    MctsNode* tree = pool + root * capacity;
    int32_t* m = meta + root * kMetaStride;
    const uint64_t root_id = (root_base + (uint64_t)root) << 32;
    const int first = m[kMetaRollouts];
    const int tid = kWarpPerRoot ? lane : (int)threadIdx.x;          // index within the root's thread group
    const int group = kWarpPerRoot ? 32 : (int)blockDim.x;
    auto group_sync = [&]() { if (kWarpPerRoot) __syncwarp(); else __syncthreads(); };
    for (int it = 0; it < n_rollouts; ++it) {
        const uint64_t base = root_id + (uint64_t)(uint32_t)(first + it);
        if (tid < 32) {                                                    // mcts.py:269-277
            int node = m[kMetaRoot], depth = 0;
            while (tree[node].has_p && !tree[node].terminal) {
                const int a = warp_uct_select(tree[node], c_puct, lane);
                if (tree[node].child[a][0] < 0 && !warp_expand(tree, m, capacity, node, a, L, lane)) break;
                if (lane == 0) { sh_path_node[depth] = node; sh_path_act[depth] = a; }
                uint32_t c0 = (uint32_t)base, c1 = (uint32_t)(base >> 32), c2 = (uint32_t)depth, c3 = kDomainSelect;
                philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
                const int second = tree[node].child[a][1];
                node = (second >= 0 && (c1 & 1u)) ? second : tree[node].child[a][0];
                ++depth;
            }
            if (lane == 0) { sh_leaf = node; sh_depth = depth; sh_rtot = 0; }
        }
        group_sync();
        const MctsNode& leaf = tree[sh_leaf];
        int r = 0;
        for (int sim = tid; sim < num_sims; sim += group)
            r += mcts_sim_reward(leaf, seed, base, (uint32_t)sim, L);
        r = __reduce_add_sync(0xFFFFFFFFu, r);
        if (lane == 0 && r) atomicAdd(&sh_rtot, r);
        group_sync();
        if (tid == 0) {
            if (!tree[sh_leaf].terminal) tree[sh_leaf].has_p = 1;             // mcts.py:189-191
            mcts_backprop(tree, sh_path_node, sh_path_act, sh_depth, sh_rtot, num_sims);
        }
        group_sync();
    }
    if (tid == 0) m[kMetaRollouts] = first + n_rollouts;
}

__global__ void __launch_bounds__(kThreads)
k_mcts_stats(const MctsNode* __restrict__ pool, int64_t capacity, const int32_t* __restrict__ meta,
             int32_t* __restrict__ n_out, double* __restrict__ q_out, int32_t* __restrict__ ntot_out,
             uint8_t* __restrict__ choose_out, int64_t n_roots) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n_roots; r += stride) {
        const MctsNode& root = pool[r * capacity + meta[r * kMetaStride + kMetaRoot]];
        for (int a = 0; a < 36; ++a) {
            if (n_out) n_out[36 * r + a] = (int32_t)root.n[a];
            if (q_out) q_out[36 * r + a] = root.n[a] ? d_div(root.w[a], (double)root.n[a]) : 0.0;
        }
        if (ntot_out) ntot_out[r] = (int32_t)root.ntot;
        if (choose_out) choose_out[r] = (uint8_t)mcts_choose(root);
    }
}

__global__ void __launch_bounds__(kThreads)
k_mcts_sync(MctsNode* __restrict__ pool, int64_t capacity, int32_t* __restrict__ meta,
            const uint8_t* __restrict__ action, const qttt_state* __restrict__ now, int64_t n_roots) {
    __shared__ __align__(16) uint8_t smem[kLutStepBytes];
    stage_luts(smem, kLutStepBytes);
    const Luts L = luts_from_image(smem);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n_roots; r += stride)
        mcts_sync(pool + r * capacity, meta + r * kMetaStride, capacity, (int)action[r], load_state(now, r), L);
}

// ------------------------------------------------------------------------------ launch helpers
// Resident blocks of a kernel on the current device (SM count x occupancy), for the kernels
// that still run as one persistent wave.  Cached per (kernel, device).
static int resident_blocks(const void* kernel, int threads) {
    struct Entry { const void* k; int dev, threads, blocks; };
    static Entry cache[64];
    static int used = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    for (int i = 0; i < used; ++i)
        if (cache[i].k == kernel && cache[i].dev == dev && cache[i].threads == threads) return cache[i].blocks;
    int sms = 148, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    const int blocks = sms * per_sm;
    if (used < 64) cache[used++] = Entry{kernel, dev, threads, blocks};   // benign race: worst case a recompute
    return blocks;
}
template <class Kernel>
static int grid_for(Kernel kernel, int64_t n) {
    const int resident = resident_blocks(reinterpret_cast<const void*>(kernel), kThreads);
    const int64_t want = (n + kThreads - 1) / kThreads;
    return (int)(want < resident ? (want < 1 ? 1 : want) : resident);
}
static int check_launch() {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QTTT_OK : -(1000 + (int)e);
}
static bool misaligned(const void* p, uintptr_t a) { return p && (reinterpret_cast<uintptr_t>(p) & (a - 1)); }

// QTTT_NO_DEVICE unless the current device can run the sm_100a kernels of this library.
static int device_ok() {
    static int verdict[64];                 // 0 unknown, 1 ok, 2 not usable
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return QTTT_ERR_NO_DEVICE; }
    if (dev < 0 || dev >= 64) return QTTT_OK;
    if (verdict[dev] == 0) {
        int major = 0;
        const bool ok = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess && major == 10;
        verdict[dev] = ok ? 1 : 2;
    }
    return verdict[dev] == 1 ? QTTT_OK : QTTT_ERR_NO_DEVICE;
}

// Tuning knob of the step kernels, read once from the environment (experiments only; the default
// is what the numbers in DESIGN.md were measured with).
static int step_iters() {
    static const int v = []() {
        const char* e = getenv("QTTT_STEP_ITERS");
        const int x = e ? atoi(e) : kStepIters;
        return x < 1 ? 1 : (x > 64 ? 64 : x);
    }();
    return v;
}
// Chunks per block for a batch of n games: the tuned value when the batch is large enough to give
// every SM many blocks, fewer for small batches (4096 envs must not end up in two blocks).
static int iters_for(int64_t n, int tuned) {
    const int64_t chunks = (n + kThreads - 1) / kThreads;
    const int64_t it = chunks / (148 * 16);
    return (int)(it < 1 ? 1 : (it > tuned ? tuned : it));
}
}  // namespace qttt

using namespace qttt;

extern "C" {

int qttt_abi_version(void) { return QTTT_ABI_VERSION; }

const char* qttt_strerror(int rc) {
    switch (rc) {
        case QTTT_OK: return "ok";
        case QTTT_ERR_ARG: return "qttt: invalid argument (NULL buffer, negative size or bad enum)";
        case QTTT_ERR_ALIGN: return "qttt: buffer not aligned for its element type";
        case QTTT_ERR_NO_DEVICE: return "qttt: no usable CUDA device (the kernels are built for sm_100a only)";
        default: break;
    }
    if (rc <= -1000) return cudaGetErrorString((cudaError_t)(-rc - 1000));
    return "qttt: unknown error code";
}

int qttt_reset(qttt_state* state, uint64_t* mask, int64_t n, void* stream) {
    return qttt_reset_all(state, mask, nullptr, nullptr, nullptr, n, stream);
}

int qttt_reset_all(qttt_state* state, uint64_t* mask, float* reward, uint8_t* done, uint8_t* status,
                   int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(mask, 8) || misaligned(reward, 4)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    k_reset<<<grid_for(k_reset, n), kThreads, 0, (cudaStream_t)stream>>>(state, mask, reward, done, status, n);
    return check_launch();
}

}  // extern "C"

// Kernel launch that allows programmatic dependent launch on the stream (the kernel must order
// itself after its predecessor with griddepcontrol.wait before touching anything the predecessor
// writes -- k_step does).  QTTT_PDL=0 in the environment turns it off.
static bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("QTTT_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), int grid, int threads, cudaStream_t st, Args... args) {
    if (!pdl_enabled()) {
        kernel<<<grid, threads, 0, st>>>(args...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

// Launches k_step over [0, n) in slices of at most 2^31 games (32-bit indices in the kernel).
template <int kFmt, bool kRandom, int kMode>
static int launch_step_mode(StepArgs a, int64_t n, cudaStream_t st) {
    const int64_t kSlice = 1ll << 31;
    const bool full = !kRandom && a.coin && a.reward && a.done && a.mask && a.status;
    const int act_bytes = kFmt == QTTT_ACT_PAIR ? 2 : 1;
    for (int64_t lo = 0; lo < n; lo += kSlice) {
        const int64_t m = n - lo < kSlice ? n - lo : kSlice;
        StepArgs b = a;
        b.state = a.state + lo;
        b.action = a.action ? a.action + act_bytes * lo : nullptr;
        b.coin = a.coin ? a.coin + lo : nullptr;
        b.game_base = a.game_base + (uint64_t)lo;
        b.reward = a.reward ? a.reward + lo : nullptr;
        b.done = a.done ? a.done + lo : nullptr;
        b.mask = a.mask ? a.mask + lo : nullptr;
        b.status = a.status ? a.status + lo : nullptr;
        b.action_out = a.action_out ? a.action_out + lo : nullptr;
        b.coin_out = a.coin_out ? a.coin_out + lo : nullptr;
        const int iters = iters_for(m, step_iters());
        b.n = (uint32_t)m;
        b.iters = iters;
        const int grid = chunk_grid(m, iters);
        bool launched = false;
        if constexpr (kFmt == QTTT_ACT_INDEX && !kRandom && kMode != kStepAutoNext) {
            if (full) {                           // the headline shape gets the specialised variants
                launch_pdl(k_step<kFmt, false, true, kMode>, grid, kThreads, st, b);
                launched = true;
            }
        }
        if (!launched) launch_pdl(k_step<kFmt, kRandom, false, kMode>, grid, kThreads, st, b);
        const int rc = check_launch();
        if (rc != QTTT_OK) return rc;
    }
    return QTTT_OK;
}

template <int kFmt, bool kRandom>
static int launch_step(const StepArgs& a, uint32_t flags, int64_t n, cudaStream_t st) {
    if (flags & QTTT_STEP_FRESH) return launch_step_mode<kFmt, kRandom, kStepFresh>(a, n, st);
    if (flags & QTTT_STEP_AUTORESET) return launch_step_mode<kFmt, kRandom, kStepAuto>(a, n, st);
    if (flags & QTTT_STEP_AUTORESET_NEXT) return launch_step_mode<kFmt, kRandom, kStepAutoNext>(a, n, st);
    return launch_step_mode<kFmt, kRandom, kStepPlain>(a, n, st);
}

static int step_entry(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
                      uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags, float* reward,
                      uint8_t* done, uint64_t* mask, uint8_t* status, int64_t n, void* stream) {
    if (action_format != QTTT_ACT_INDEX && action_format != QTTT_ACT_PAIR) return QTTT_ERR_ARG;
    const uint32_t modes = flags & (QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT);
    if ((flags & ~(QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT)) || (modes & (modes - 1)))
        return QTTT_ERR_ARG;                                   // unknown flag, or two modes at once
    if (n == 0) return QTTT_OK;
    if (!state || !action || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(mask, 8) || misaligned(reward, 4)) return QTTT_ERR_ALIGN;
    if (action_format == QTTT_ACT_PAIR && misaligned(action, 2)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    StepArgs a{};
    a.state = state;
    a.action = static_cast<const uint8_t*>(action);
    a.coin = coin;
    a.seed = seed;
    a.game_base = game_base;
    a.dword = domain_word(0u, epoch);
    a.reward = reward;
    a.done = done;
    a.mask = mask;
    a.status = status;
    if (action_format == QTTT_ACT_INDEX) return launch_step<QTTT_ACT_INDEX, false>(a, flags, n, (cudaStream_t)stream);
    return launch_step<QTTT_ACT_PAIR, false>(a, flags, n, (cudaStream_t)stream);
}

extern "C" {

int qttt_step(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
              uint64_t seed, uint64_t game_base, float* reward, uint8_t* done, uint64_t* mask,
              uint8_t* status, int64_t n, void* stream) {
    return step_entry(state, action, action_format, coin, seed, game_base, 0, 0, reward, done, mask, status, n, stream);
}

int qttt_reset_step(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
                    uint64_t seed, uint64_t game_base, float* reward, uint8_t* done, uint64_t* mask,
                    uint8_t* status, int64_t n, void* stream) {
    return step_entry(state, action, action_format, coin, seed, game_base, 0, QTTT_STEP_FRESH, reward, done, mask,
                      status, n, stream);
}

int qttt_step_ex(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
                 uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags, float* reward,
                 uint8_t* done, uint64_t* mask, uint8_t* status, int64_t n, void* stream) {
    return step_entry(state, action, action_format, coin, seed, game_base, epoch, flags, reward, done, mask, status,
                      n, stream);
}

static int packed_entry(qttt_state* state, const uint8_t* action_coin, uint16_t* result, qttt_state* obs,
                        int64_t n, bool zero_copy, cudaStream_t st) {
    const int64_t kSlice = 1ll << 31;
    for (int64_t lo = 0; lo < n; lo += kSlice) {
        const int64_t m = n - lo < kSlice ? n - lo : kSlice;
        const int iters = iters_for(m, step_iters());
        qttt_state* o = obs ? obs + lo : nullptr;
        const int zc_iters = iters < kZcMaxIters ? iters : kZcMaxIters;
        if (zero_copy)
            k_step_packed_zc<false><<<chunk_grid(m, zc_iters), kThreads, 0, st>>>(state + lo, action_coin + lo, result + lo,
                                                                                   o, (uint32_t)m, zc_iters);
        else
            k_step_packed<false><<<chunk_grid(m, iters), kThreads, 0, st>>>(state + lo, action_coin + lo, result + lo, o,
                                                                             (uint32_t)m, iters);
        const int rc = check_launch();
        if (rc != QTTT_OK) return rc;
    }
    return QTTT_OK;
}

int qttt_step_packed(qttt_state* state, const uint8_t* action_coin, uint16_t* result, int64_t n,
                     void* stream) {
    return qttt_step_packed_obs(state, action_coin, result, nullptr, n, stream);
}

int qttt_step_packed_obs(qttt_state* state, const uint8_t* action_coin, uint16_t* result, qttt_state* obs,
                         int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || !action_coin || !result || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(result, 2) || misaligned(obs, 16)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    return packed_entry(state, action_coin, result, obs, n, false, (cudaStream_t)stream);
}

int qttt_step_packed_mapped(qttt_state* state, const uint8_t* action_coin_host, uint16_t* result_host,
                            qttt_state* obs_host, int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || !action_coin_host || !result_host || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(result_host, 2) || misaligned(obs_host, 16)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    return packed_entry(state, action_coin_host, result_host, obs_host, n, true, (cudaStream_t)stream);
}

int qttt_step_packed12_mapped(qttt_state* state, const uint8_t* action_coin_host, uint16_t* result12_host,
                              int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || !action_coin_host || !result12_host || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(result12_host, 2)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    const int64_t kSlice = 1ll << 31;          // a multiple of 4: every slice starts on a group boundary
    for (int64_t lo = 0; lo < n; lo += kSlice) {
        const int64_t m = n - lo < kSlice ? n - lo : kSlice;
        int iters = iters_for(m, step_iters());
        iters = iters < kZcMaxIters ? iters : kZcMaxIters;
        k_step_packed_zc<true><<<chunk_grid(m, iters), kThreads, 0, (cudaStream_t)stream>>>(
            state + lo, action_coin_host + lo, result12_host + (lo / 4) * 3, nullptr, (uint32_t)m, iters);
        const int rc = check_launch();
        if (rc != QTTT_OK) return rc;
    }
    return QTTT_OK;
}

int qttt_step_packed_host(qttt_state* state, const uint8_t* action_coin_host, uint16_t* result_host,
                          uint8_t* in_dev, uint16_t* out_dev, int64_t n, int64_t slice,
                          void* const* streams, int n_streams) {
    return qttt_step_packed_host_obs(state, action_coin_host, result_host, nullptr, in_dev, out_dev, n, slice,
                                     streams, n_streams);
}

int qttt_step_packed_host_obs(qttt_state* state, const uint8_t* action_coin_host, uint16_t* result_host,
                              qttt_state* obs_host, uint8_t* in_dev, uint16_t* out_dev, int64_t n,
                              int64_t slice, void* const* streams, int n_streams) {
    if (n == 0) return QTTT_OK;
    if (!state || !action_coin_host || !result_host || !in_dev || !out_dev || !streams || n < 0 ||
        slice < 1 || slice > (1ll << 31) || n_streams < 1)
        return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(out_dev, 2) || misaligned(result_host, 2) || misaligned(obs_host, 16))
        return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    int k = 0;
    for (int64_t lo = 0; lo < n; lo += slice, ++k) {
        const int64_t m = n - lo < slice ? n - lo : slice;
        cudaStream_t st = (cudaStream_t)streams[k % n_streams];
        cudaError_t e = cudaMemcpyAsync(in_dev + lo, action_coin_host + lo, (size_t)m, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return -(1000 + (int)e);
        const int rc = packed_entry(state + lo, in_dev + lo, out_dev + lo, nullptr, m, false, st);
        if (rc != QTTT_OK) return rc;
        e = cudaMemcpyAsync(result_host + lo, out_dev + lo, (size_t)m * 2, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return -(1000 + (int)e);
        if (obs_host) {     // the observation: the packed states themselves, straight from the state array
            e = cudaMemcpyAsync(obs_host + lo, state + lo, (size_t)m * sizeof(qttt_state), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return -(1000 + (int)e);
        }
    }
    return QTTT_OK;
}

int qttt_step_packed12_host(qttt_state* state, const uint8_t* action_coin_host, uint16_t* result12_host,
                            uint8_t* in_dev, uint16_t* out12_dev, int64_t n, int64_t slice,
                            void* const* streams, int n_streams) {
    if (n == 0) return QTTT_OK;
    if (!state || !action_coin_host || !result12_host || !in_dev || !out12_dev || !streams || n < 0 ||
        slice < 4 || (slice & 3) || slice > (1ll << 31) || n_streams < 1)
        return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(out12_dev, 2) || misaligned(result12_host, 2)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    int k = 0;
    for (int64_t lo = 0; lo < n; lo += slice, ++k) {
        const int64_t m = n - lo < slice ? n - lo : slice;
        const int64_t w0 = (lo / 4) * 3, words = ((m + 3) / 4) * 3;       // slices start on group boundaries
        cudaStream_t st = (cudaStream_t)streams[k % n_streams];
        cudaError_t e = cudaMemcpyAsync(in_dev + lo, action_coin_host + lo, (size_t)m, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return -(1000 + (int)e);
        const int iters = iters_for(m, step_iters());
        k_step_packed<true><<<chunk_grid(m, iters), kThreads, 0, st>>>(state + lo, in_dev + lo, out12_dev + w0, nullptr,
                                                                       (uint32_t)m, iters);
        const int rc = check_launch();
        if (rc != QTTT_OK) return rc;
        e = cudaMemcpyAsync(result12_host + w0, out12_dev + w0, (size_t)words * 2, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return -(1000 + (int)e);
    }
    return QTTT_OK;
}

int qttt_step_packed_host_obs12(qttt_state* state, const uint8_t* action_coin_host, uint32_t* obs12_host,
                                uint8_t* in_dev, uint32_t* obs12_dev, int64_t n, int64_t slice,
                                void* const* streams, int n_streams) {
    if (n == 0) return QTTT_OK;
    if (!state || !action_coin_host || !obs12_host || !in_dev || !obs12_dev || !streams || n < 0 ||
        slice < 1 || slice > (1ll << 31) || n_streams < 1)
        return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(obs12_dev, 4) || misaligned(obs12_host, 4)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    int k = 0;
    for (int64_t lo = 0; lo < n; lo += slice, ++k) {
        const int64_t m = n - lo < slice ? n - lo : slice;
        cudaStream_t st = (cudaStream_t)streams[k % n_streams];
        cudaError_t e = cudaMemcpyAsync(in_dev + lo, action_coin_host + lo, (size_t)m, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return -(1000 + (int)e);
        const int iters = iters_for(m, step_iters());
        k_step_packed_obs12<<<chunk_grid(m, iters), kThreads, 0, st>>>(state + lo, in_dev + lo, obs12_dev + 3 * lo,
                                                                       (uint32_t)m, iters);
        const int rc = check_launch();
        if (rc != QTTT_OK) return rc;
        e = cudaMemcpyAsync(obs12_host + 3 * lo, obs12_dev + 3 * lo, (size_t)m * 12, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return -(1000 + (int)e);
    }
    return QTTT_OK;
}

int qttt_step_random(qttt_state* state, uint64_t seed, uint64_t game_base, uint8_t* action_out,
                     uint8_t* coin_out, float* reward, uint8_t* done, uint64_t* mask,
                     uint8_t* status, int64_t n, void* stream) {
    return qttt_step_random_ex(state, seed, game_base, 0, 0, action_out, coin_out, reward, done, mask, status, n, stream);
}

int qttt_step_random_ex(qttt_state* state, uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags,
                        uint8_t* action_out, uint8_t* coin_out, float* reward, uint8_t* done,
                        uint64_t* mask, uint8_t* status, int64_t n, void* stream) {
    const uint32_t modes = flags & (QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT);
    if ((flags & ~(QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT)) || (modes & (modes - 1)))
        return QTTT_ERR_ARG;
    if (n == 0) return QTTT_OK;
    if (!state || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(mask, 8) || misaligned(reward, 4)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    StepArgs a{};
    a.state = state;
    a.seed = seed;
    a.game_base = game_base;
    a.dword = domain_word(0u, epoch);
    a.reward = reward;
    a.done = done;
    a.mask = mask;
    a.status = status;
    a.action_out = action_out;
    a.coin_out = coin_out;
    return launch_step<QTTT_ACT_INDEX, true>(a, flags, n, (cudaStream_t)stream);
}

int qttt_observe(const qttt_state* state, int8_t* classical, int8_t* moves, uint8_t* n_moves,
                 int8_t* q_p1, int8_t* q_p2, uint8_t* turn, int8_t* rounds, float* reward_p1,
                 uint8_t* winner, uint8_t* mask_bool, int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(reward_p1, 4) || misaligned(q_p2, 8) || misaligned(rounds, 2))
        return QTTT_ERR_ALIGN;
    const ObsOut o{classical, moves, n_moves, q_p1, q_p2, turn, rounds, reward_p1, winner, mask_bool};
    // the staged rows leave with 16-byte stores when every staged array starts on a 16-byte boundary
    const bool aligned16 = !(misaligned(classical, 16) || misaligned(moves, 16) || misaligned(q_p1, 16) ||
                             misaligned(mask_bool, 16));
    const bool env_set = classical && q_p1 && q_p2 && turn;
    const bool extras = moves || n_moves || rounds || reward_p1 || winner || mask_bool;
    const bool all_set = env_set && moves && n_moves && rounds && reward_p1 && winner && mask_bool;
    const int grid = chunk_grid(n, iters_for(n, 4));
    if (all_set)
        k_observe<kObsAll><<<grid, kThreads, 0, (cudaStream_t)stream>>>(state, o, n, aligned16);
    else if (env_set && !extras)
        k_observe<kObsEnv><<<grid, kThreads, 0, (cudaStream_t)stream>>>(state, o, n, aligned16);
    else
        k_observe<kObsAny><<<grid, kThreads, 0, (cudaStream_t)stream>>>(state, o, n, aligned16);
    return check_launch();
}

int qttt_features(const qttt_state* state, float* features, int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || !features || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(features, 16)) return QTTT_ERR_ALIGN;
    {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t want = (n + kFeatThreads - 1) / kFeatThreads;
        const int64_t cap = (int64_t)sms * 4;                     // 4 blocks of 47 KB fit one SM
        k_features<<<(int)(want < cap ? want : cap), kFeatThreads, 0, (cudaStream_t)stream>>>(state, features, n);
    }
    return check_launch();
}

int qttt_env1(qttt_state* state, int op, int a, int b, int coin, uint64_t seed, uint64_t epoch,
              void* record_host, uint32_t seq, void* stream) {
    if (!state || !record_host || op < 0 || op > 2) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(record_host, 16)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    k_env1<<<1, 32, 0, (cudaStream_t)stream>>>(state, a, b, coin, op, seed, domain_word(0u, epoch),
                                               static_cast<uint8_t*>(record_host), seq);
    return check_launch();
}

int qttt_qeval1(const qttt_state* state_host, int action, void* record_host, uint32_t seq, void* stream) {
    if (!state_host || !record_host || action < 0 || action > 255) return QTTT_ERR_ARG;
    if (misaligned(record_host, 16)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    k_qeval1<<<1, 32, 0, (cudaStream_t)stream>>>(state_host->w[0], state_host->w[1], state_host->w[2], state_host->w[3],
                                                 (uint32_t)action, static_cast<uint8_t*>(record_host), seq);
    return check_launch();
}

int qttt_get_mask(const qttt_state* state, uint8_t* illegal_mask, int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || !illegal_mask || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(illegal_mask, 4)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    k_get_mask<<<chunk_grid(n, iters_for(n, 4)), kThreads, 0, (cudaStream_t)stream>>>(state, illegal_mask, n);
    return check_launch();
}

}  // extern "C"

template <int kFmt, int kMode>
static int launch_step_features(const StepFeatArgs& fa, int64_t n, cudaStream_t st) {
    constexpr int kSmem = kLutStepBytes + (kFeatThreads / 32) * 32 * kFeatStride * (int)sizeof(float);
    static bool configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(k_step_features<kFmt, kMode>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return -(1000 + (int)e);
        configured[dev] = true;
    }
    const int64_t kSlice = 1ll << 31;
    const int act_bytes = kFmt == QTTT_ACT_PAIR ? 2 : 1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int64_t lo = 0; lo < n; lo += kSlice) {
        const int64_t m = n - lo < kSlice ? n - lo : kSlice;
        StepFeatArgs b = fa;
        b.st.state = fa.st.state + lo;
        b.st.action = fa.st.action + act_bytes * lo;
        b.st.coin = fa.st.coin ? fa.st.coin + lo : nullptr;
        b.st.game_base = fa.st.game_base + (uint64_t)lo;
        b.st.reward = fa.st.reward ? fa.st.reward + lo : nullptr;
        b.st.done = fa.st.done ? fa.st.done + lo : nullptr;
        b.st.mask = fa.st.mask ? fa.st.mask + lo : nullptr;
        b.st.status = fa.st.status ? fa.st.status + lo : nullptr;
        b.st.n = (uint32_t)m;
        b.features = fa.features + 180 * lo;
        b.illegal_mask = fa.illegal_mask ? fa.illegal_mask + 36 * lo : nullptr;
        const int64_t want = (m + kFeatThreads - 1) / kFeatThreads;
        const int64_t cap = (int64_t)sms * 4 * 8;                 // 4 blocks of 53 KB per SM, 8 waves
        k_step_features<kFmt, kMode><<<(int)(want < cap ? want : cap), kFeatThreads, kSmem, st>>>(b);
        const int rc = check_launch();
        if (rc != QTTT_OK) return rc;
    }
    return QTTT_OK;
}

template <int kFmt>
static int launch_step_features_fmt(const StepFeatArgs& fa, uint32_t flags, int64_t n, cudaStream_t st) {
    if (flags & QTTT_STEP_FRESH) return launch_step_features<kFmt, kStepFresh>(fa, n, st);
    if (flags & QTTT_STEP_AUTORESET) return launch_step_features<kFmt, kStepAuto>(fa, n, st);
    if (flags & QTTT_STEP_AUTORESET_NEXT) return launch_step_features<kFmt, kStepAutoNext>(fa, n, st);
    return launch_step_features<kFmt, kStepPlain>(fa, n, st);
}

extern "C" {

int qttt_step_features(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
                       uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags, float* reward,
                       uint8_t* done, uint64_t* mask, uint8_t* status, float* features,
                       uint8_t* illegal_mask, int64_t n, void* stream) {
    if (action_format != QTTT_ACT_INDEX && action_format != QTTT_ACT_PAIR) return QTTT_ERR_ARG;
    const uint32_t modes = flags & (QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT);
    if ((flags & ~(QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT)) || (modes & (modes - 1)))
        return QTTT_ERR_ARG;
    if (n == 0) return QTTT_OK;
    if (!state || !action || !features || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(mask, 8) || misaligned(reward, 4) || misaligned(features, 16) ||
        misaligned(illegal_mask, 4))
        return QTTT_ERR_ALIGN;
    if (action_format == QTTT_ACT_PAIR && misaligned(action, 2)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    StepFeatArgs fa{};
    fa.st.state = state;
    fa.st.action = static_cast<const uint8_t*>(action);
    fa.st.coin = coin;
    fa.st.seed = seed;
    fa.st.game_base = game_base;
    fa.st.dword = domain_word(0u, epoch);
    fa.st.reward = reward;
    fa.st.done = done;
    fa.st.mask = mask;
    fa.st.status = status;
    fa.features = features;
    fa.illegal_mask = illegal_mask;
    if (action_format == QTTT_ACT_INDEX)
        return launch_step_features_fmt<QTTT_ACT_INDEX>(fa, flags, n, (cudaStream_t)stream);
    return launch_step_features_fmt<QTTT_ACT_PAIR>(fa, flags, n, (cudaStream_t)stream);
}

}  // extern "C"

template <int kFmt, int kMode>
static int launch_step_obs(const StepObsArgs& fa, int64_t n, cudaStream_t st) {
    const int64_t kSlice = 1ll << 31;
    const int act_bytes = kFmt == QTTT_ACT_PAIR ? 2 : 1;
    for (int64_t lo = 0; lo < n; lo += kSlice) {
        const int64_t m = n - lo < kSlice ? n - lo : kSlice;
        StepObsArgs b = fa;
        b.st.state = fa.st.state + lo;
        b.st.action = fa.st.action + act_bytes * lo;
        b.st.coin = fa.st.coin ? fa.st.coin + lo : nullptr;
        b.st.game_base = fa.st.game_base + (uint64_t)lo;
        b.st.reward = fa.st.reward ? fa.st.reward + lo : nullptr;
        b.st.done = fa.st.done ? fa.st.done + lo : nullptr;
        b.st.mask = fa.st.mask ? fa.st.mask + lo : nullptr;
        b.st.status = fa.st.status ? fa.st.status + lo : nullptr;
        b.st.n = (uint32_t)m;
        b.st.iters = iters_for(m, 4);
        b.obs.classical = fa.obs.classical + 9 * lo;
        b.obs.q1 = fa.obs.q1 + 10 * lo;
        b.obs.q2 = fa.obs.q2 + 8 * lo;
        b.obs.turn = fa.obs.turn + lo;
        k_step_obs<kFmt, kMode><<<chunk_grid(m, b.st.iters), kThreads, 0, st>>>(b);
        const int rc = check_launch();
        if (rc != QTTT_OK) return rc;
    }
    return QTTT_OK;
}

template <int kFmt>
static int launch_step_obs_fmt(const StepObsArgs& fa, uint32_t flags, int64_t n, cudaStream_t st) {
    if (flags & QTTT_STEP_FRESH) return launch_step_obs<kFmt, kStepFresh>(fa, n, st);
    if (flags & QTTT_STEP_AUTORESET) return launch_step_obs<kFmt, kStepAuto>(fa, n, st);
    if (flags & QTTT_STEP_AUTORESET_NEXT) return launch_step_obs<kFmt, kStepAutoNext>(fa, n, st);
    return launch_step_obs<kFmt, kStepPlain>(fa, n, st);
}

extern "C" {

int qttt_step_obs(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
                  uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags, float* reward,
                  uint8_t* done, uint64_t* mask, uint8_t* status, int8_t* classical, int8_t* q_p1,
                  int8_t* q_p2, uint8_t* turn, int64_t n, void* stream) {
    if (action_format != QTTT_ACT_INDEX && action_format != QTTT_ACT_PAIR) return QTTT_ERR_ARG;
    const uint32_t modes = flags & (QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT);
    if ((flags & ~(QTTT_STEP_FRESH | QTTT_STEP_AUTORESET | QTTT_STEP_AUTORESET_NEXT)) || (modes & (modes - 1)))
        return QTTT_ERR_ARG;
    if (n == 0) return QTTT_OK;
    if (!state || !action || !classical || !q_p1 || !q_p2 || !turn || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(mask, 8) || misaligned(reward, 4) || misaligned(q_p2, 8))
        return QTTT_ERR_ALIGN;
    if (action_format == QTTT_ACT_PAIR && misaligned(action, 2)) return QTTT_ERR_ALIGN;
    if (const int rc = device_ok()) return rc;
    StepObsArgs fa{};
    fa.st.state = state;
    fa.st.action = static_cast<const uint8_t*>(action);
    fa.st.coin = coin;
    fa.st.seed = seed;
    fa.st.game_base = game_base;
    fa.st.dword = domain_word(0u, epoch);
    fa.st.reward = reward;
    fa.st.done = done;
    fa.st.mask = mask;
    fa.st.status = status;
    fa.obs.classical = classical;
    fa.obs.q1 = q_p1;
    fa.obs.q2 = q_p2;
    fa.obs.turn = turn;
    fa.aligned16 = !(misaligned(classical, 16) || misaligned(q_p1, 16));
    if (action_format == QTTT_ACT_INDEX)
        return launch_step_obs_fmt<QTTT_ACT_INDEX>(fa, flags, n, (cudaStream_t)stream);
    return launch_step_obs_fmt<QTTT_ACT_PAIR>(fa, flags, n, (cudaStream_t)stream);
}

int qttt_pack(qttt_state* state, const int8_t* classical, const int8_t* moves,
              const uint8_t* n_moves, int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || !classical || !moves || !n_moves || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16)) return QTTT_ERR_ALIGN;
    if (n == 0) return QTTT_OK;
    k_pack<<<grid_for(k_pack, n), kThreads, 0, (cudaStream_t)stream>>>(state, classical, moves, n_moves, n);
    return check_launch();
}

int qttt_qeval_both(const qttt_state* state, const uint8_t* action, qttt_state* next0,
                    qttt_state* next1, uint64_t* board0, uint64_t* board1, int8_t* sq0,
                    int8_t* sq1, uint8_t* closes, float* result_prob, int64_t n, void* stream) {
    if (n == 0) return QTTT_OK;
    if (!state || !action || n < 0) return QTTT_ERR_ARG;
    if (misaligned(state, 16) || misaligned(next0, 16) || misaligned(next1, 16) ||
        misaligned(board0, 8) || misaligned(board1, 8) || misaligned(result_prob, 4))
        return QTTT_ERR_ALIGN;
    if (n == 0) return QTTT_OK;
    if (board0 && board1 && closes && !next0 && !next1 && !sq0 && !sq1 && !result_prob) {
        // 32-bit indices inside the kernel: batches beyond 2^31 boards go slice by slice
        const int64_t slice = (int64_t)1 << 30;
        for (int64_t at = 0; at < n; at += slice) {
            const int64_t m = n - at < slice ? n - at : slice;
            const int iters = iters_for(m, 8);
            k_qeval_boards<<<chunk_grid(m, iters), kThreads, 0, (cudaStream_t)stream>>>(
                state + at, action + at, board0 + at, board1 + at, closes + at, (uint32_t)m, iters);
        }
        return check_launch();
    }
    if (sq0 || sq1)
        k_qeval_both<true><<<chunk_grid(n, iters_for(n, 4)), kThreads, 0, (cudaStream_t)stream>>>(state, action, next0, next1, board0, board1, sq0, sq1, closes, result_prob, n);
    else
        k_qeval_both<false><<<chunk_grid(n, iters_for(n, 4)), kThreads, 0, (cudaStream_t)stream>>>(state, action, next0, next1, board0, board1, sq0, sq1, closes, result_prob, n);
    return check_launch();
}

int qttt_rollout(const qttt_state* roots, int64_t n_roots, int32_t n_rollouts, uint64_t seed,
                 int32_t* tallies, float* value, int64_t* steps_total, void* stream) {
    if (n_roots == 0 && n_rollouts > 0) return QTTT_OK;
    if (!roots || n_roots < 0 || n_rollouts <= 0 || n_roots > 0x7FFFFFFF) return QTTT_ERR_ARG;
    if (misaligned(roots, 16) || misaligned(tallies, 4) || misaligned(value, 4) || misaligned(steps_total, 8))
        return QTTT_ERR_ALIGN;
    if (n_roots == 0) return QTTT_OK;
    // one block per root up to 8 resident waves of blocks (then roots are taken grid-stride): the
    // block scheduler keeps every SM full until the last roots
    const int64_t cap = 8ll * resident_blocks(reinterpret_cast<const void*>(k_rollout), kThreads);
    k_rollout<<<(int)(n_roots < cap ? n_roots : cap), kThreads, 0, (cudaStream_t)stream>>>(
        roots, n_roots, n_rollouts, seed, tallies, value, reinterpret_cast<unsigned long long*>(steps_total));
    return check_launch();
}

int qttt_mcts_node_bytes(void) { return (int)sizeof(MctsNode); }

int qttt_mcts_init(void* pool, int64_t capacity, int32_t* meta, const qttt_state* roots,
                   int64_t n_roots, void* stream) {
    if (n_roots == 0) return QTTT_OK;
    if (!pool || !meta || !roots || capacity < 1 || n_roots < 0) return QTTT_ERR_ARG;
    if (misaligned(pool, 16) || misaligned(roots, 16) || misaligned(meta, 4)) return QTTT_ERR_ALIGN;
    k_mcts_init<<<grid_for(k_mcts_init, n_roots), kThreads, 0, (cudaStream_t)stream>>>(
        static_cast<MctsNode*>(pool), capacity, meta, roots, n_roots);
    return check_launch();
}

int qttt_mcts_run(void* pool, int64_t capacity, int32_t* meta, int32_t n_rollouts,
                  int32_t num_simulations, double c_puct, uint64_t seed, uint64_t root_base,
                  int64_t n_roots, void* stream) {
    if (n_roots == 0 || n_rollouts == 0) return QTTT_OK;
    if (!pool || !meta || capacity < 1 || n_roots < 0 || n_rollouts < 0 || num_simulations < 1 ||
        num_simulations > (int32_t)kMaxSims || n_roots > 0x7FFFFFFF || root_base + (uint64_t)n_roots > (1ull << 20))
        return QTTT_ERR_ARG;
    if (misaligned(pool, 16) || misaligned(meta, 4)) return QTTT_ERR_ALIGN;
    int threads = ((num_simulations + 31) / 32) * 32;
    threads = threads > kThreads ? kThreads : threads;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // one warp per root: with more roots than one-warp blocks can keep resident (11 per SM: the
    // tables), several roots share a block; fewer roots stay spread over all SMs, one per block
    if (threads == 32 && n_roots > (int64_t)sms * 11) {
        const int64_t blocks = (n_roots + kMctsWarps - 1) / kMctsWarps;
        k_mcts_run<true><<<(int)blocks, 32 * kMctsWarps, 0, (cudaStream_t)stream>>>(
            static_cast<MctsNode*>(pool), capacity, meta, n_roots, n_rollouts, num_simulations, c_puct, seed, root_base);
    } else if (threads == 32) {
        k_mcts_run<true><<<(int)n_roots, 32, 0, (cudaStream_t)stream>>>(
            static_cast<MctsNode*>(pool), capacity, meta, n_roots, n_rollouts, num_simulations, c_puct, seed, root_base);
    } else {
        k_mcts_run<false><<<(int)n_roots, threads, 0, (cudaStream_t)stream>>>(
            static_cast<MctsNode*>(pool), capacity, meta, n_roots, n_rollouts, num_simulations, c_puct, seed, root_base);
    }
    return check_launch();
}

int qttt_mcts_stats(const void* pool, int64_t capacity, const int32_t* meta, int32_t* n_visits,
                    double* q_values, int32_t* n_total, uint8_t* choose, int64_t n_roots, void* stream) {
    if (n_roots == 0) return QTTT_OK;
    if (!pool || !meta || capacity < 1 || n_roots < 0) return QTTT_ERR_ARG;
    if (misaligned(pool, 16) || misaligned(q_values, 8) || misaligned(n_visits, 4) || misaligned(n_total, 4))
        return QTTT_ERR_ALIGN;
    k_mcts_stats<<<grid_for(k_mcts_stats, n_roots), kThreads, 0, (cudaStream_t)stream>>>(
        static_cast<const MctsNode*>(pool), capacity, meta, n_visits, q_values, n_total, choose, n_roots);
    return check_launch();
}

int qttt_mcts_sync(void* pool, int64_t capacity, int32_t* meta, const uint8_t* action,
                   const qttt_state* now, int64_t n_roots, void* stream) {
    if (n_roots == 0) return QTTT_OK;
    if (!pool || !meta || !action || !now || capacity < 1 || n_roots < 0) return QTTT_ERR_ARG;
    if (misaligned(pool, 16) || misaligned(now, 16)) return QTTT_ERR_ALIGN;
    k_mcts_sync<<<grid_for(k_mcts_sync, n_roots), kThreads, 0, (cudaStream_t)stream>>>(
        static_cast<MctsNode*>(pool), capacity, meta, action, now, n_roots);
    return check_launch();
}

int qttt_sweep(int64_t game_lo, int64_t game_hi, uint64_t seed, int64_t* stats, void* stream) {
    if (game_hi < game_lo) return QTTT_ERR_ARG;
    if (game_hi == game_lo) return QTTT_OK;
    if (!stats) return QTTT_ERR_ARG;
    if (misaligned(stats, 8)) return QTTT_ERR_ALIGN;
    if (game_hi == game_lo) return QTTT_OK;
    k_sweep<<<chunk_grid(game_hi - game_lo, iters_for(game_hi - game_lo, kSweepIters)), kThreads, 0, (cudaStream_t)stream>>>(
        game_lo, game_hi, seed, reinterpret_cast<unsigned long long*>(stats));
    return check_launch();
}

}  // extern "C"
