// qttt_core.cuh -- packed game state and the per-game transition for batched quantum
// tic-tac-toe on sm_100a.  One thread owns one game; everything here is register-resident
// bitmask arithmetic (no loops over sets, no local memory).
//
// Reference semantics being reproduced (Oxel40/qtttgym, citations into /root/reference):
//   Board.make_move        qtttgym/board.py:9-25     legality, append, autofill
//   Board.update_qstructs  qtttgym/board.py:27-69    component lookup, cycle test, collapse
//   QEvalClassic.eval      qtttgym/qeval.py:5-51     which square each move collapses into
//   Board.check_win        qtttgym/board.py:71-115   line rounds
//   Env.step               qtttgym/env.py:34-53      reward (-0.0 / -1.0), terminated
//   GameState actions      mcts.py:19-27, 87-91      36-way legal mask
//
// How the collapse is computed here (NOT how the reference does it): before the closing move
// the component is a tree (a cycle would have collapsed earlier).  The measurement gives the
// closing move to one of its two squares `t` (coin), and every other move of the component
// then has exactly one free square left: its endpoint farther from `t`.  So the collapse map
// is "root the tree at t, every edge falls into its child endpoint".  The reference's leaf
// peeling + cycle walk (qeval.py:23-49) yields the same map; tests pin the two against each
// other through the oracle and fixtures recorded from the live reference.
//
// The same breadth-first absorption also answers "are a and b already connected?" (the
// cycle test, board.py:42), so one pass structure serves both.
//
// The header is __host__ __device__ clean so that tests can run the very same transition on
// the build container's CPU (tests/hostemu) before GPU time is spent.  The shipped library
// only ever calls it from kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QTTT_HD __host__ __device__ __forceinline__
#else
#define QTTT_HD inline
#endif

namespace qttt {

// ------------------------------------------------------------------------------------
// Packed state: 16 bytes per game, one uint4 load / store.
//
//   x : E0[0:9]  E1[9:18]  E2[18:27]  n_moves[27:31]
//   y : E3[0:9]  E4[9:18]  E5[18:27]  P3 squares 0..4 [27:32]
//   z : E6[0:9]  E7[9:18]  E8[18:27]  P3 squares 5..8 [27:31]
//   w : P0[0:9]  P1[9:18]  P2[18:27]  (5 spare bits, zero)
//
//   E_i : 9-bit square set of move i -- two bits for a spooky pair (a, b), one bit for the
//         autofill entry (s, s, i) (board.py:25), zero for an empty slot.  a < b is implicit.
//   P_k : bit-plane k of v[s] = board[s] + 1 (0 = not classical, 1..9 = owner index + 1).
//         Plane 0 is therefore "owned by X" (even move index), and the classical set is
//         P0|P1|P2|P3.
// ------------------------------------------------------------------------------------
struct State { uint32_t x, y, z, w; };

constexpr uint32_t M9 = 0x1FFu;

QTTT_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
QTTT_HD int ctz32(uint32_t v) {   // v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
QTTT_HD int flo32(uint32_t v) {   // index of highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
QTTT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Largest value among the lanes of the warp that are executing this call together (the value
// itself on the host).
// kFullWarp: the caller guarantees that all 32 lanes execute the call together (no __activemask
// query needed).
template <bool kFullWarp = false>
QTTT_HD uint32_t warp_max_u32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(kFullWarp ? 0xFFFFFFFFu : __activemask(), v);
#else
    return v;
#endif
}

// true if the predicate holds in any lane executing this call together (the value on the host)
template <bool kFullWarp = false>
QTTT_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(kFullWarp ? 0xFFFFFFFFu : __activemask(), p) != 0;
#else
    return p;
#endif
}

QTTT_HD State empty_state() { return State{0u, 0u, 0u, 0u}; }
QTTT_HD uint32_t n_moves(const State& s) { return (s.x >> 27) & 15u; }
QTTT_HD uint32_t plane3(const State& s) { return (s.y >> 27) | ((s.z >> 22) & 0x1E0u); }
QTTT_HD uint32_t plane0(const State& s) { return s.w & M9; }
QTTT_HD uint32_t plane1(const State& s) { return (s.w >> 9) & M9; }
QTTT_HD uint32_t plane2(const State& s) { return (s.w >> 18) & M9; }
QTTT_HD uint32_t classical(const State& s) {
    return ((s.w | (s.w >> 9) | (s.w >> 18)) & M9) | plane3(s);
}
template <int I>
QTTT_HD uint32_t edge(const State& s) {
    const uint32_t word = I < 3 ? s.x : (I < 6 ? s.y : s.z);
    return (word >> (9 * (I % 3))) & M9;
}
QTTT_HD uint32_t edge_dyn(const State& s, uint32_t i) {
    const uint32_t q = (i * 11u) >> 5;
    const uint32_t word = q == 0 ? s.x : (q == 1 ? s.y : s.z);
    return (word >> (9u * (i - 3u * q))) & M9;
}

// Lookup tables: a constant image in global memory (built at compile time), staged into
// shared memory by each block.  The step kernels only need the first kLutStepBytes.
struct LutImage {
    uint64_t legal[512];    // free-square set -> 36-bit legal mask            (mcts.py:19-27)
    uint16_t pair[256];     // action index -> E mask of (i, j); 0 for index >= 36 (mcts.py:339-343)
    uint8_t  line[512];     // square set -> 1 if it contains one of the 8 lines (board.py:84-110)
    uint32_t nrow[16][4];   // per len(moves): multipliers (see NRow)
    uint64_t spread[512];   // square set -> the same bits at a 4-bit stride
    uint8_t  rankpair[10][40];  // [f][k] -> r1 | r2 << 4: the k-th pair (r1 < r2) of f items, lexicographic
    uint16_t nthbit[512][9];    // [square set][r] -> one-hot of its r-th member (0 past the end)
};
// Row n of LutImage::nrow: everything the transition needs that depends only on n = len(moves),
// as multipliers so the work lands on the (otherwise idle) IMAD pipe:
//   mx, my, mz : E << (bit offset of move slot n) == E * m? in the word that holds slot n (else 0)
//   kp         : plane pattern of v = n + 1 for word w: (v&1) | (v&2)<<8 | (v&4)<<16
struct NRow { uint32_t mx, my, mz, kp; };
constexpr int kLutStepBytes = 512 * 8 + 256 * 2 + 512 + 256;   // legal + pair + line + nrow = 5376
constexpr int kLutQevalBytes = kLutStepBytes + 512 * 8;                // + spread = 9472
constexpr int kLutPolicyBytes = kLutQevalBytes + 400 + 512 * 9 * 2;    // + rankpair + nthbit = 19088 (whole image)
constexpr int kLutBytes = (int)sizeof(LutImage);        // 19088
static_assert(sizeof(LutImage) == 9472 + 400 + 9216, "LutImage layout");
static_assert(kLutPolicyBytes % 16 == 0 && kLutStepBytes % 16 == 0, "staged in 16-byte vectors");

constexpr LutImage make_lut_image() {
    LutImage t{};
    const uint16_t lines[8] = {0x007, 0x038, 0x1C0, 0x049, 0x092, 0x124, 0x054, 0x111};
    for (uint32_t m = 0; m < 512; ++m) {
        uint64_t lm = 0, sp = 0;
        int k = 0;
        for (int i = 0; i < 9; ++i)
            for (int j = i + 1; j < 9; ++j, ++k)
                if ((m >> i & 1u) && (m >> j & 1u)) lm |= 1ull << k;
        for (int s = 0; s < 9; ++s)
            if (m >> s & 1u) sp |= 1ull << (4 * s);
        uint8_t any = 0;
        for (int l = 0; l < 8; ++l)
            if ((m & lines[l]) == lines[l]) any = 1;
        t.legal[m] = lm;
        t.spread[m] = sp;
        t.line[m] = any;
    }
    int k = 0;
    for (int i = 0; i < 9; ++i)
        for (int j = i + 1; j < 9; ++j, ++k) t.pair[k] = (uint16_t)((1u << i) | (1u << j));
    for (uint32_t n = 0; n < 9; ++n) {
        const uint32_t m = 1u << (9u * (n % 3u)), v = n + 1u;
        t.nrow[n][0] = n < 3u ? m : 0u;
        t.nrow[n][1] = (n >= 3u && n < 6u) ? m : 0u;
        t.nrow[n][2] = n >= 6u ? m : 0u;
        t.nrow[n][3] = (v & 1u) | ((v & 2u) << 8) | ((v & 4u) << 16);
    }
    for (int f = 0; f < 10; ++f) {
        int kk = 0;
        for (int r1 = 0; r1 < f; ++r1)
            for (int r2 = r1 + 1; r2 < f; ++r2, ++kk) t.rankpair[f][kk] = (uint8_t)(r1 | (r2 << 4));
    }
    for (uint32_t m = 0; m < 512; ++m) {
        int r = 0;
        for (int sq = 0; sq < 9; ++sq)
            if (m >> sq & 1u) t.nthbit[m][r++] = (uint16_t)(1u << sq);
    }
    return t;
}

struct Luts {
    const uint64_t* legal;
    const uint16_t* pair;
    const uint8_t*  line;
    const NRow*     nrow;
    const uint8_t*  rankpair;
    const uint16_t* nthbit;
    const uint64_t* spread;
};
QTTT_HD Luts luts_from_image(const void* img) {
    const LutImage* t = reinterpret_cast<const LutImage*>(img);
    Luts l;
    l.legal = t->legal;
    l.pair = t->pair;
    l.line = t->line;
    l.nrow = reinterpret_cast<const NRow*>(t->nrow);
    l.rankpair = &t->rankpair[0][0];
    l.nthbit = &t->nthbit[0][0];
    l.spread = t->spread;
    return l;
}

// E mask for an (a, b) byte pair as passed to Env.step (any order; env.py:37-40).  0 = cannot
// be a move: same square (board.py:10-12) or off the board (IndexError path / negative).
QTTT_HD uint32_t pair_to_edge(uint32_t a, uint32_t b) {
    const bool ok = (a < 9u) & (b < 9u) & (a != b);
    return ok ? ((1u << a) | (1u << b)) : 0u;
}

struct StepResult {
    uint32_t illegal;     // 1: the reference would have raised -> no-op (env.py:36-43)
    uint32_t collapsed;   // 1: a measurement happened (the coin was consumed)
    uint32_t classical;   // classical squares after the step
    uint32_t n;           // len(moves) after the step
};

// ------------------------------------------------------------------------------------
// The absorption sweep.  absorb<I>() takes move slot I into the reached set R when it touches
// it, and books the square c it brought in (its child endpoint when the tree is rooted at
// the start square) straight into the bit-plane accumulators: move I owns c with index I,
// i.e. v = I + 1, whose planes 0..2 are the constant pattern kPlane[I] for word w (plane 3,
// only v = 8, goes to A3).  c is disjoint from everything accumulated so far, so add == or.
// ------------------------------------------------------------------------------------
template <int I> struct PlanePattern {
    static constexpr uint32_t v = I + 1;
    static constexpr uint32_t value = (v & 1u) | ((v & 2u) << 8) | ((v & 4u) << 16);
};

template <int I>
QTTT_HD void absorb(uint32_t E, uint32_t& R, uint32_t& W, uint32_t& A3) {
#if defined(__CUDA_ARCH__)
    // PTX pins the shape.  The integer ALU pipe (LOP3 / SHF / ISETP / SEL) issues one warp
    // instruction per TWO cycles, the FMA pipe (IMAD) one per cycle, so everything that can be a
    // multiply-add is one:
    //   h = E & R, p = (h != 0)      one LOP3 with a predicate output            (ALU pipe)
    //   c = E - h                    the endpoint not yet reached: IMAD h * -1 + E (FMA pipe)
    //   @p W += c * pattern(I)       books c as owned by move I in the plane word  (FMA pipe)
    //   @p R += c                                                                 (FMA pipe)
    // (Measured alternatives, DESIGN.md section 9: c as a second LOP3 -- the same speed, the kernel is
    // issue-bound, not ALU-pipe-bound; R and W in one 64-bit accumulator updated by a single
    // predicated IMAD.WIDE -- three instructions per slot but 9 % slower.)
    if (I < 7) {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 h, c;\n\t"
            "and.b32 h, %2, %1;\n\t"
            "setp.ne.u32 p, h, 0;\n\t"
            "mad.lo.u32 c, h, 0xffffffff, %2;\n\t"
            "@p mad.lo.u32 %0, c, %3, %0;\n\t"
            "@p mad.lo.u32 %1, c, 1, %1;\n\t}"
            : "+r"(W), "+r"(R) : "r"(E), "n"(PlanePattern<I>::value));
    } else {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 h, c;\n\t"
            "and.b32 h, %2, %1;\n\t"
            "setp.ne.u32 p, h, 0;\n\t"
            "mad.lo.u32 c, h, 0xffffffff, %2;\n\t"
            "@p mad.lo.u32 %0, c, 1, %0;\n\t"
            "@p mad.lo.u32 %1, c, 1, %1;\n\t}"
            : "+r"(A3), "+r"(R) : "r"(E));
    }
#else
    if (E & R) {
        const uint32_t c = E & ~R;
        if (I < 7) W += c * PlanePattern<I>::value; else A3 += c;
        R |= c;
    }
#endif
}

template <int I> QTTT_HD uint32_t slot(uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t word = I < 3 ? x : (I < 6 ? y : z);
    return (I % 3 == 0) ? (word & M9) : ((word >> (9 * (I % 3))) & M9);
}

constexpr int kFixedSweepMax = 5;   // up to this many slots the sweep runs a fixed schedule (below)
static_assert(kFixedSweepMax <= 5, "the fixed schedules of sweep / sweep2 list slots 0..4");

// Forward sweeps over move slots 0..N-1 until the reached set stops growing.
// `stop`: a set R cannot grow beyond, or ~0 when none is known.  With 8 moves on the board the
// live edges form ONE spanning tree of the free squares (every measured component of k squares
// consumed exactly k moves, so #free = #live edges + 9 - len(moves)): the sweep is complete the
// moment R holds every free square and the pass that would find nothing is skipped.
template <int N, bool kTargets>
QTTT_HD void sweep(uint32_t x, uint32_t y, uint32_t z, uint32_t& R, uint32_t& W, uint32_t& A3, uint32_t* T,
                   uint32_t stop = ~0u) {
    const uint32_t E0 = N > 0 ? slot<0>(x, y, z) : 0u, E1 = N > 1 ? slot<1>(x, y, z) : 0u;
    const uint32_t E2 = N > 2 ? slot<2>(x, y, z) : 0u, E3 = N > 3 ? slot<3>(x, y, z) : 0u;
    const uint32_t E4 = N > 4 ? slot<4>(x, y, z) : 0u, E5 = N > 5 ? slot<5>(x, y, z) : 0u;
    const uint32_t E6 = N > 6 ? slot<6>(x, y, z) : 0u, E7 = N > 7 ? slot<7>(x, y, z) : 0u;
    if (N <= kFixedSweepMax) {
        // Few slots: a fixed schedule with no convergence test.  A slot is absorbed in pass p when the
        // tree path from the start square to it turns back to a lower slot p - 1 times, so N passes
        // always suffice and the last one can only still absorb slot 0.  A warp of 32 games needs
        // (nearly) that many passes anyway, and the per-pass test is gone.
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
            if (N > 0) absorb<0>(E0, R, W, A3);
            if (N > 1) absorb<1>(E1, R, W, A3);
            if (N > 2) absorb<2>(E2, R, W, A3);
            if (N > 3) absorb<3>(E3, R, W, A3);
            if (N > 4) absorb<4>(E4, R, W, A3);
        }
        if (N > 0) absorb<0>(E0, R, W, A3);
    } else {
        // More slots: passes until nothing changes.  32 games together practically never finish in
        // fewer than 4 passes with 6 or 7 slots (2 with 8, where the spanning-tree stop applies), so
        // the first passes run without the test.
#pragma unroll
        for (int p = 0; p < (N == 8 ? 1 : 3); ++p) {
            absorb<0>(E0, R, W, A3);
            absorb<1>(E1, R, W, A3);
            absorb<2>(E2, R, W, A3);
            absorb<3>(E3, R, W, A3);
            absorb<4>(E4, R, W, A3);
            absorb<5>(E5, R, W, A3);
            if (N > 6) absorb<6>(E6, R, W, A3);
            if (N > 7) absorb<7>(E7, R, W, A3);
        }
        uint32_t before;
        do {
            before = R;
            if (N > 0) absorb<0>(E0, R, W, A3);
            if (N > 1) absorb<1>(E1, R, W, A3);
            if (N > 2) absorb<2>(E2, R, W, A3);
            if (N > 3) absorb<3>(E3, R, W, A3);
            if (N > 4) absorb<4>(E4, R, W, A3);
            if (N > 5) absorb<5>(E5, R, W, A3);
            if (N > 6) absorb<6>(E6, R, W, A3);
            if (N > 7) absorb<7>(E7, R, W, A3);
        } while (N > 1 && R != before && (N < 8 || R != stop));
    }
    if (kTargets) {
        // Which square did each absorbed edge bring in?  Only the qeval kernel asks: replay
        // the rooting from the start square (T[8]) with plain code.
        uint32_t Rr = T[8], Es[8] = {E0, E1, E2, E3, E4, E5, E6, E7}, prev;
        do {
            prev = Rr;
            for (int i = 0; i < N; ++i)
                if (Es[i] & Rr) { T[i] |= Es[i] & ~Rr; Rr |= Es[i]; }
        } while (Rr != prev);
    }
}

// The same sweep from TWO start squares at once (the two measurement outcomes of one closing
// move, qeval.py:35): one pass over the move slots advances both rootings, so enumerating both
// outcomes costs one loop instead of two transitions.  The two absorptions of a slot are
// independent instruction chains (ILP 2).
template <int N>
QTTT_HD void sweep2(uint32_t x, uint32_t y, uint32_t z, uint32_t& Ra, uint32_t& Wa, uint32_t& A3a,
                    uint32_t& Rb, uint32_t& Wb, uint32_t& A3b, uint32_t stop = ~0u) {
    const uint32_t E0 = N > 0 ? slot<0>(x, y, z) : 0u, E1 = N > 1 ? slot<1>(x, y, z) : 0u;
    const uint32_t E2 = N > 2 ? slot<2>(x, y, z) : 0u, E3 = N > 3 ? slot<3>(x, y, z) : 0u;
    const uint32_t E4 = N > 4 ? slot<4>(x, y, z) : 0u, E5 = N > 5 ? slot<5>(x, y, z) : 0u;
    const uint32_t E6 = N > 6 ? slot<6>(x, y, z) : 0u, E7 = N > 7 ? slot<7>(x, y, z) : 0u;
    if (N <= kFixedSweepMax) {      // the fixed schedule of sweep<N>
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
            if (N > 0) { absorb<0>(E0, Ra, Wa, A3a); absorb<0>(E0, Rb, Wb, A3b); }
            if (N > 1) { absorb<1>(E1, Ra, Wa, A3a); absorb<1>(E1, Rb, Wb, A3b); }
            if (N > 2) { absorb<2>(E2, Ra, Wa, A3a); absorb<2>(E2, Rb, Wb, A3b); }
            if (N > 3) { absorb<3>(E3, Ra, Wa, A3a); absorb<3>(E3, Rb, Wb, A3b); }
            if (N > 4) { absorb<4>(E4, Ra, Wa, A3a); absorb<4>(E4, Rb, Wb, A3b); }
        }
        if (N > 0) { absorb<0>(E0, Ra, Wa, A3a); absorb<0>(E0, Rb, Wb, A3b); }
    } else {
        // (running the first passes without the test, as sweep<N> does, measured 5-12 % slower here)
        uint32_t before;
        do {
            before = Ra + (Rb << 9);
            if (N > 0) { absorb<0>(E0, Ra, Wa, A3a); absorb<0>(E0, Rb, Wb, A3b); }
            if (N > 1) { absorb<1>(E1, Ra, Wa, A3a); absorb<1>(E1, Rb, Wb, A3b); }
            if (N > 2) { absorb<2>(E2, Ra, Wa, A3a); absorb<2>(E2, Rb, Wb, A3b); }
            if (N > 3) { absorb<3>(E3, Ra, Wa, A3a); absorb<3>(E3, Rb, Wb, A3b); }
            if (N > 4) { absorb<4>(E4, Ra, Wa, A3a); absorb<4>(E4, Rb, Wb, A3b); }
            if (N > 5) { absorb<5>(E5, Ra, Wa, A3a); absorb<5>(E5, Rb, Wb, A3b); }
            if (N > 6) { absorb<6>(E6, Ra, Wa, A3a); absorb<6>(E6, Rb, Wb, A3b); }
            if (N > 7) { absorb<7>(E7, Ra, Wa, A3a); absorb<7>(E7, Rb, Wb, A3b); }
        } while ((Ra + (Rb << 9)) != before && (N < 8 || (Ra & Rb) != stop));
    }
}

// Board.make_move for one game.  `enew`: E mask of the requested pair (0 = malformed);
// `coin`: 0 -> the closing move falls into its smaller square (qeval.py:35).
//
// kTargets: also report, per move index, the square set it collapsed into in this
// measurement (tgt[0..8], zero when not part of it) -- eval()'s return value by move index.
//
// Instruction budget notes: the integer ALU pipe (LOP3/SHF/ISETP/SEL) is the binding resource
// of the step kernel, so (a) the sweep touches only the n move slots that exist -- n is
// uniform across a warp when a batch is stepped in lock-step, so the switch does not diverge
// -- and (b) bit placement is written as multiply-add by table constants (IMAD pipe).
//
// kKnownC: the caller already holds the classical set of `s` (playouts carry it from the
// previous step's result) and passes it as `known_c`.
template <bool kTargets = false, bool kKnownC = false, bool kFullWarp = false>
QTTT_HD StepResult step_core(State& s, uint32_t enew, uint32_t coin, const Luts& L, uint32_t* tgt = nullptr,
                             uint32_t known_c = 0u) {
    const uint32_t x = s.x, y = s.y, z = s.z, w = s.w;
    const uint32_t n = (x >> 27) & 15u;
    const NRow row = L.nrow[n];
    const uint32_t C = kKnownC ? known_c : classical(s);
    const bool legal = (enew != 0u) & ((enew & C) == 0u) & (n < 9u);   // board.py:10-15
    enew = legal ? enew : 0u;

    // t = square the closing move would take; o = its other square.
    const uint32_t lo = enew & (0u - enew);
    const uint32_t t = coin ? (enew ^ lo) : lo;
    const uint32_t o = enew ^ t;

    // Live edges are exactly the moves whose squares are not classical; collapsed moves have
    // both squares classical and can never touch R (which starts on a free square), so the
    // raw E fields can be used unfiltered.
    uint32_t R = t;
    uint32_t W = t * row.kp;          // planes 0..2 of the closing move (index n) on square t
    uint32_t A3 = 0u;                 // plane 3 (v = 8) of squares taken by move 7, from sweep<8>
    uint32_t T[9] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, t};
    // An illegal request (a swallowed no-op) has nothing to sweep.  The case is chosen ONCE PER
    // WARP -- the largest len(moves) among its lanes -- so the switch never diverges: empty
    // slots hold E = 0 and can never touch R, which makes sweep<N> exact for every n <= N.  A
    // batch stepped in lock-step pays exactly its own n; a desynchronised batch (envs at
    // different plies in one warp) pays one sweep<max n> instead of one case per distinct n.
    switch (warp_max_u32<kFullWarp>(legal ? n : 0u)) {
        case 1: sweep<1, kTargets>(x, y, z, R, W, A3, T); break;
        case 2: sweep<2, kTargets>(x, y, z, R, W, A3, T); break;
        case 3: sweep<3, kTargets>(x, y, z, R, W, A3, T); break;
        case 4: sweep<4, kTargets>(x, y, z, R, W, A3, T); break;
        case 5: sweep<5, kTargets>(x, y, z, R, W, A3, T); break;
        case 6: sweep<6, kTargets>(x, y, z, R, W, A3, T); break;
        case 7: sweep<7, kTargets>(x, y, z, R, W, A3, T); break;
        case 8: sweep<8, kTargets>(x, y, z, R, W, A3, T, n == 8u ? (~C & M9) : ~0u); break;
        default: break;
    }

    // a, b already connected -> cycle (board.py:42); colf is the 0/1 multiplier form.
    const uint32_t colf = (R & o) != 0u ? 1u : 0u;

    // board[square] = move index for every move of the component (board.py:53-54): the
    // accumulated plane words are committed only when the move closed a cycle.
    uint32_t wn = w + W * colf;
    uint32_t Cn = C | (R * colf);

    // moves.append((a, b, n))  (board.py:19): slot n is empty, so add == or.
    uint32_t xn = x + enew * row.mx;
    uint32_t yn = y + enew * row.my;
    uint32_t zn = z + enew * row.mz;
    uint32_t inc = legal ? 1u : 0u;

    // Plane 3 (move indices 7 and 8) and the autofill exist only once 7 moves are on the board:
    // a branch on the WARP's largest len(moves), so earlier plies do not pay for them.
    if (warp_any<kFullWarp>(n >= 7u)) {
        A3 += (n >= 7u) ? t : 0u;       // the closing move itself: v = n + 1 in {8, 9}
        A3 *= colf;
        // Autofill (board.py:21-25): one free square left.  Only reachable right after the
        // collapse triggered by move 7, so the entry is always (s, s, 8): v = 9 -> planes 0, 3.
        const uint32_t fr = ~Cn & M9;
        const bool fill = (colf != 0u) & (popc32(fr) == 1);
        const uint32_t fs = fill ? fr : 0u;
        zn += fs << 18;
        wn += fs;
        A3 += fs;
        Cn |= fs;
        inc += fill ? 1u : 0u;
        yn += A3 << 27;                 // squares 0..4 of plane 3 (higher bits fall off the word)
        zn += (A3 >> 5) << 27;          // squares 5..8
    }
    xn += inc << 27;

    if (kTargets) {
        const uint32_t cm = 0u - colf;
#pragma unroll
        for (int i = 0; i < 8; ++i) tgt[i] = T[i] & cm;
        tgt[8] = 0u;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            if ((uint32_t)i == n) tgt[i] = t & cm;
    }

    s.x = xn; s.y = yn; s.z = zn; s.w = wn;
    StepResult out;
    out.illegal = legal ? 0u : 1u;
    out.collapsed = colf;
    out.classical = Cn;
    out.n = n + inc;
    return out;
}

// One outcome's commit, shared by step_both: the accumulated plane words become part of the
// board when the move closed a cycle (board.py:53-54), then the autofill (board.py:21-25).
// (xa, ya, za): the move words with the new move already appended.
QTTT_HD uint32_t commit_outcome(State& out, uint32_t xa, uint32_t ya, uint32_t za, uint32_t w, uint32_t C,
                                uint32_t R, uint32_t W, uint32_t A3, uint32_t colf, uint32_t inc, uint32_t& n_out) {
    uint32_t wn = w + W * colf;
    A3 *= colf;
    uint32_t Cn = C | (R * colf);
    const uint32_t fr = ~Cn & M9;
    const bool fill = (colf != 0u) & (popc32(fr) == 1);
    const uint32_t fs = fill ? fr : 0u;
    wn += fs;
    A3 += fs;
    Cn |= fs;
    inc += fill ? 1u : 0u;
    out.x = xa + (inc << 27);
    out.y = ya + (A3 << 27);
    out.z = za + (fs << 18) + ((A3 >> 5) << 27);
    out.w = wn;
    n_out = inc;
    return Cn;
}

// Board.make_move for a game whose len(moves) is the compile-time constant N (playouts that walk
// the plies of a game in order, all lanes of a warp at the same ply): no table row, no warp vote,
// no switch -- the slot offsets and plane patterns of move N are immediates and sweep<N> is called
// directly.  `enew` must be a legal pair of this state or 0 (no-op); C = classical(s).
template <int N>
QTTT_HD StepResult step_fixed(State& s, uint32_t enew, uint32_t coin, uint32_t C) {
    static_assert(N >= 0 && N <= 8, "a move needs a free slot");
    constexpr uint32_t v = N + 1;
    constexpr uint32_t kp = (v & 1u) | ((v & 2u) << 8) | ((v & 4u) << 16);
    const uint32_t x = s.x, y = s.y, z = s.z, w = s.w;
    const uint32_t lo = enew & (0u - enew);
    const uint32_t t = coin ? (enew ^ lo) : lo;
    const uint32_t o = enew ^ t;
    uint32_t R = t, W = t * kp, A3 = 0u;
    sweep<N, false>(x, y, z, R, W, A3, nullptr, N == 8 ? (~C & M9) : ~0u);
    const uint32_t colf = (R & o) != 0u ? 1u : 0u;                         // board.py:42
    uint32_t wn = w + W * colf;
    uint32_t Cn = C | (R * colf);
    constexpr uint32_t sh = 9u * (N % 3);
    uint32_t xn = x + (N < 3 ? (enew << sh) : 0u);                         // board.py:19
    uint32_t yn = y + ((N >= 3 && N < 6) ? (enew << sh) : 0u);
    uint32_t zn = z + (N >= 6 ? (enew << sh) : 0u);
    uint32_t inc = enew ? 1u : 0u;
    if (N >= 7) {                                                          // plane 3 and the autofill
        A3 += t;
        A3 *= colf;
        const uint32_t fr = ~Cn & M9;
        const bool fill = (colf != 0u) & (popc32(fr) == 1);                // board.py:21-25
        const uint32_t fs = fill ? fr : 0u;
        zn += fs << 18;
        wn += fs;
        A3 += fs;
        Cn |= fs;
        inc += fill ? 1u : 0u;
        yn += A3 << 27;
        zn += (A3 >> 5) << 27;
    }
    xn += inc << 27;
    s.x = xn; s.y = yn; s.z = zn; s.w = wn;
    StepResult out;
    out.illegal = enew ? 0u : 1u;
    out.collapsed = colf;
    out.classical = Cn;
    out.n = (uint32_t)N + inc;
    return out;
}

// Both measurement outcomes of one move in ONE sweep (board.py:42-56 with qeval.py:35 taking
// either value; what MCTS._step enumerates by rejection sampling, mcts.py:233-267).  The two
// outcomes are the rootings of the same tree at the closing move's two squares; sweep2 grows
// both in one pass over the move slots.  When no cycle closes the two successors are equal.
struct BothResult {
    uint32_t illegal, collapsed;
    uint32_t classical0, classical1;   // classical squares of the two successors
    uint32_t n0, n1;                   // len(moves) of the two successors
};
// The sweep both callers share: the two rootings of the closing move's component.
struct BothSweep {
    uint32_t n, C, enew, legal, colf;
    uint32_t Ra, Wa, A3a, Rb, Wb, A3b;
    NRow row;
};
template <bool kFullWarp = false>
QTTT_HD BothSweep sweep_both(const State& s, uint32_t enew, const Luts& L) {
    const uint32_t x = s.x, y = s.y, z = s.z;
    BothSweep r;
    r.n = (x >> 27) & 15u;
    r.row = L.nrow[r.n];
    r.C = classical(s);
    const bool legal = (enew != 0u) & ((enew & r.C) == 0u) & (r.n < 9u);   // board.py:10-15
    enew = legal ? enew : 0u;
    const uint32_t a = enew & (0u - enew), b = enew ^ a;                // coin 0 -> a, coin 1 -> b
    uint32_t Ra = a, Rb = b;
    uint32_t Wa = a * r.row.kp, Wb = b * r.row.kp;
    uint32_t A3a = (r.n >= 7u) ? a : 0u, A3b = (r.n >= 7u) ? b : 0u;
    switch (warp_max_u32<kFullWarp>(legal ? r.n : 0u)) {
        case 1: sweep2<1>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b); break;
        case 2: sweep2<2>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b); break;
        case 3: sweep2<3>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b); break;
        case 4: sweep2<4>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b); break;
        case 5: sweep2<5>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b); break;
        case 6: sweep2<6>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b); break;
        case 7: sweep2<7>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b); break;
        case 8: sweep2<8>(x, y, z, Ra, Wa, A3a, Rb, Wb, A3b, r.n == 8u ? (~r.C & M9) : ~0u); break;
        default: break;
    }
    r.colf = (Ra & b) != 0u ? 1u : 0u;                                  // board.py:42
    r.enew = enew;
    r.legal = legal ? 1u : 0u;
    r.Ra = Ra; r.Wa = Wa; r.A3a = A3a; r.Rb = Rb; r.Wb = Wb; r.A3b = A3b;
    return r;
}

QTTT_HD BothResult step_both(const State& s, uint32_t enew, const Luts& L, State& s0, State& s1) {
    const BothSweep k = sweep_both<false>(s, enew, L);
    const uint32_t xa = s.x + k.enew * k.row.mx, ya = s.y + k.enew * k.row.my, za = s.z + k.enew * k.row.mz;   // board.py:19
    BothResult r;
    uint32_t i0, i1;
    r.classical0 = commit_outcome(s0, xa, ya, za, s.w, k.C, k.Ra, k.Wa, k.A3a, k.colf, k.legal, i0);
    r.classical1 = commit_outcome(s1, xa, ya, za, s.w, k.C, k.Rb, k.Wb, k.A3b, k.colf, k.legal, i1);
    r.n0 = k.n + i0;
    r.n1 = k.n + i1;
    r.illegal = k.legal ^ 1u;
    r.collapsed = k.colf;
    return r;
}

// One outcome's board as 9 nibbles (square s at bits 4s..4s+3; 0 = free, else move index + 1)
// from its plane words: the sum over planes of spread(plane) << plane.  Nibbles never carry into
// each other, so the two 32-bit halves are combined separately with multiply-adds (FMA pipe).
struct alignas(8) Halves { uint32_t lo, hi; };
template <bool kPlane3 = true>
QTTT_HD uint64_t nibbles_of_planes(uint32_t p012, uint32_t p3, const Luts& L) {
    const Halves* sp = reinterpret_cast<const Halves*>(L.spread);
    const Halves s0 = sp[p012 & M9], s1 = sp[(p012 >> 9) & M9], s2 = sp[(p012 >> 18) & M9];
    uint32_t lo = s0.lo + 2u * s1.lo + 4u * s2.lo;
    uint32_t hi = s0.hi + 2u * s1.hi + 4u * s2.hi;
    if (kPlane3) {
        const Halves s3 = sp[p3 & M9];
        lo += 8u * s3.lo;
        hi += 8u * s3.hi;
    }
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

// Both outcome BOARDS of one move and nothing else (config 3's shape): the successor states are
// never assembled -- the plane words go straight from the sweep's accumulators to the nibble form.
// Marks 8 and 9 (plane 3) and the autofill exist only from the eighth move on, so that part runs
// only in warps holding such a game.
struct BoardsBoth { uint64_t board0, board1; uint32_t collapsed; };
template <bool kFullWarp = false>
QTTT_HD BoardsBoth boards_both(const State& s, uint32_t enew, const Luts& L) {
    const BothSweep k = sweep_both<kFullWarp>(s, enew, L);
    BoardsBoth r;
    r.collapsed = k.colf;
    uint32_t cf = k.colf;
#if defined(__CUDA_ARCH__)
    asm("" : "+r"(cf));      // keep `x + y * cf` a multiply-add (FMA pipe) instead of a select and an add (ALU pipe)
#endif
    if (warp_any<kFullWarp>(k.n >= 7u)) {
        const uint32_t p3 = plane3(s);
        // the autofill (board.py:21-25) can only be the ninth mark: planes 0 and 3
        const uint32_t fr0 = ~(k.C | k.Ra) & M9, fr1 = ~(k.C | k.Rb) & M9;
        const uint32_t fs0 = ((k.colf != 0u) & (popc32(fr0) == 1)) ? fr0 : 0u;
        const uint32_t fs1 = ((k.colf != 0u) & (popc32(fr1) == 1)) ? fr1 : 0u;
        r.board0 = nibbles_of_planes<true>(s.w + k.Wa * cf + fs0, p3 + k.A3a * cf + fs0, L);
        r.board1 = nibbles_of_planes<true>(s.w + k.Wb * cf + fs1, p3 + k.A3b * cf + fs1, L);
    } else {
        r.board0 = nibbles_of_planes<false>(s.w + k.Wa * cf, 0u, L);
        r.board1 = nibbles_of_planes<false>(s.w + k.Wb * cf, 0u, L);
    }
    return r;
}

// Env.step's scalar outputs from the post-step state (env.py:48-51).
//   reward bits: -(1**p) * float(win)  ->  0xBF800000 (-1.0) if any line exists else
//   0x80000000 (-0.0) -- quirk Q1;  terminated: win or len(moves) > 8.
QTTT_HD uint32_t any_line(const State& s, uint32_t C, const Luts& L) {
    const uint32_t X = s.w & M9;        // plane 0: owned by an even move index (X)
    const uint32_t O = C & ~X;
    return (uint32_t)L.line[X] | (uint32_t)L.line[O];
}
QTTT_HD uint32_t reward_bits(uint32_t win) { return win ? 0xBF800000u : 0x80000000u; }

// board.py:71-115 in plane form.  Returns (pX, pO) as in check_win: -1 or the round.
QTTT_HD void win_rounds(const State& s, const Luts& L, int& px, int& po) {
    const uint32_t P0 = plane0(s), P1 = plane1(s), P2 = plane2(s), P3 = plane3(s);
    const uint32_t C = P0 | P1 | P2 | P3;
    const uint32_t x8 = P0;                       // X squares: v in {1,3,5,7,9}
    const uint32_t x6 = P0 & ~P3;                 // v <= 7
    const uint32_t x4 = x6 & ~(P1 & P2);          // v <= 5
    const uint32_t o7 = C & ~P0;                  // O squares: v in {2,4,6,8}
    const uint32_t o5 = o7 & ~P3;                 // v <= 6
    px = L.line[x4] ? 4 : (L.line[x6] ? 6 : (L.line[x8] ? 8 : -1));
    po = L.line[o5] ? 5 : (L.line[o7] ? 7 : -1);
}
// mcts.py:52-65 / strat_eval.py:21-32: 1 = X, 2 = O, 0 = none (draw or unfinished).
QTTT_HD uint32_t winner_of(int px, int po) {
    if (px > 0 && po > 0) return px < po ? 1u : 2u;
    if (px > 0) return 1u;
    if (po > 0) return 2u;
    return 0u;
}

// ------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter (game_lo, game_hi, ply, domain), key = seed.
// x0 picks the action, bit 0 of x1 is the collapse coin.
// ------------------------------------------------------------------------------------
QTTT_HD void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                           uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// k-th (0-based) set bit of a 36-bit mask, k < popcount.  Branch-free binary search.
QTTT_HD uint32_t nth_set_bit36(uint64_t m, uint32_t k) {
    uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    uint32_t c = (uint32_t)popc32(lo);
    bool up = k >= c;
    uint32_t wd = up ? hi : lo;
    k -= up ? c : 0u;
    uint32_t base = up ? 32u : 0u;
#pragma unroll
    for (int width = 16; width >= 1; width >>= 1) {
        const uint32_t part = wd & ((1u << width) - 1u);
        c = (uint32_t)popc32(part);
        up = k >= c;
        wd = up ? (wd >> width) : part;
        k -= up ? c : 0u;
        base += up ? (uint32_t)width : 0u;
    }
    return base;
}

// The same draw by table: the legal actions are the unordered pairs of free squares, and the
// action-index order (mcts.py:339-349) restricted to them is the lexicographic order of rank
// pairs (r1 < r2) over the free squares.  So the floor(x0 * m / 2^32)-th legal action is the
// pair of the r1-th and r2-th free squares, with (r1, r2) the k-th pair of f = |free| items.
// Returns the E mask of the chosen pair (0 when fewer than two squares are free).
QTTT_HD uint32_t policy_edge(uint32_t free_set, uint32_t x0, const Luts& L) {
    const uint32_t f = (uint32_t)popc32(free_set);
    const uint32_t m = (f * (f - 1u)) >> 1;
    const uint32_t k = mulhi32(x0, m);
    const uint32_t rp = L.rankpair[f * 40u + k];
    const uint16_t* row = L.nthbit + free_set * 9u;
    const uint32_t e = (uint32_t)row[rp & 15u] | (uint32_t)row[rp >> 4];
    return m ? e : 0u;
}

// Word 3 of the Philox counter: the draw domain (0 step API / sweep, 1 rollouts, 2 MCTS
// selection, 3 MCTS playouts) in bits 0..7 and the EPOCH in bits 8..31.  The epoch is what
// separates the episodes an env slot plays one after the other: (seed, game, ply) alone would
// hand every episode the same coin at the same ply.  Callers bump it on every reset (and on
// every step of an auto-resetting batch); epoch 0 is the plain (seed, game, ply) stream.
QTTT_HD uint32_t domain_word(uint32_t domain, uint64_t epoch) { return domain | ((uint32_t)epoch << 8); }

// The random draw of one ply: (action word, coin bit) = f(seed, game, ply, domain).  One
// Philox4x32-10 block, counter (game_lo, game_hi, ply >> 1, domain), serves TWO consecutive
// plies: the even ply takes (x0, x1 & 1), the odd ply (x2, x3 & 1).
// (`domain` is counter word 3 as given: a plain domain number, or domain_word(domain, epoch).)
QTTT_HD void ply_draw(uint64_t seed, uint64_t game, uint32_t ply, uint32_t domain,
                      uint32_t& action_word, uint32_t& coin) {
    uint32_t c0 = (uint32_t)game, c1 = (uint32_t)(game >> 32), c2 = ply >> 1, c3 = domain;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    action_word = (ply & 1u) ? c2 : c0;
    coin = ((ply & 1u) ? c3 : c1) & 1u;
}

// The same draw for code that walks the plies of one game in order (playouts): the block
// computed at an even ply is kept for the odd ply that follows.
struct DrawCache { uint32_t word, coin, ply; };
QTTT_HD DrawCache empty_draw_cache() { return DrawCache{0u, 0u, 0xFFFFFFFFu}; }
QTTT_HD void ply_draw_cached(uint64_t seed, uint64_t game, uint32_t ply, uint32_t domain, DrawCache& cache,
                             uint32_t& action_word, uint32_t& coin) {
    if (cache.ply == ply) {
        action_word = cache.word;
        coin = cache.coin;
        return;
    }
    uint32_t c0 = (uint32_t)game, c1 = (uint32_t)(game >> 32), c2 = ply >> 1, c3 = domain;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (ply & 1u) {
        action_word = c2;
        coin = c3 & 1u;
    } else {
        action_word = c0;
        coin = c1 & 1u;
        cache.word = c2;
        cache.coin = c3 & 1u;
        cache.ply = ply + 1u;
    }
}

// uniform-random legal action + coin for (seed, game, ply, domain)
QTTT_HD void policy_draw(uint64_t seed, uint64_t game, uint32_t ply, uint32_t domain,
                         uint64_t legal_mask, uint32_t& action, uint32_t& coin) {
    uint32_t word;
    ply_draw(seed, game, ply, domain, word, coin);
    const uint32_t m = (uint32_t)popc32((uint32_t)legal_mask) + (uint32_t)popc32((uint32_t)(legal_mask >> 32));
    action = m ? nth_set_bit36(legal_mask, mulhi32(word, m)) : 255u;
}


// ====================================================================================
// Per-game bodies of the kernels (shared with the host emulation used by CPU tests).
// All pointer arguments are the full output arrays; `i` is the game's row.
// ====================================================================================

QTTT_HD float bits_to_float(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    union { uint32_t u; float f; } c; c.u = b; return c.f;
#endif
}

// Scalar outputs of Env.step for the post-step state (env.py:46-53).
QTTT_HD void emit_step_outputs(const State& s, const StepResult& r, uint32_t status, const Luts& L,
                               float* reward, uint8_t* done, uint64_t* mask, uint8_t* status_out,
                               int64_t i) {
    const uint32_t win = any_line(s, r.classical, L);
    if (reward) reinterpret_cast<uint32_t*>(reward)[i] = reward_bits(win);   // env.py:49 (bit pattern, see k_step)
    if (done) done[i] = (uint8_t)((win != 0u) | (r.n > 8u));                  // env.py:51
    if (mask) mask[i] = L.legal[~r.classical & M9];                           // mcts.py:87-91
    if (status_out) status_out[i] = (uint8_t)status;
}

// ------------------------------------------------------------------------------------
// One game's Env.step with everything around the transition that does not touch memory
// (shared by the step kernels and the host emulation the CPU tests run).
//   kStepPlain    : Env.step (env.py:34-53)
//   kStepFresh    : Env.reset then Env.step (env.py:55-57, 34-53): the incoming state is ignored
//   kStepAuto     : if the game is over on entry (a line exists or 9 entries in moves) it is
//                   reset first, then the action is applied -- a batch never idles
//   kStepAutoNext : as kStepAuto but the action of a just-reset game is ignored (the
//                   "next-step" autoreset convention of vector envs): the step returns the
//                   fresh game
// ------------------------------------------------------------------------------------
enum : int { kStepPlain = 0, kStepFresh = 1, kStepAuto = 2, kStepAutoNext = 3 };
constexpr uint32_t kStFinished = 2u, kStReset = 4u;

struct StepOut {
    uint32_t win;          // a line exists after the step (reward = win ? -1.0f : -0.0f)
    uint32_t done;         // env.py:51
    uint32_t classical;    // classical squares after the step (legal mask = legal[~classical])
    uint32_t status;       // QTTT_ST_* (| kStReset when the game was auto-reset on entry)
    uint32_t write_state;  // the state changed (an illegal no-op leaves it as it is)
    uint32_t action, coin; // the random policy's choice (kRandom), 255 / 0 when nothing was played
};

template <bool kRandom, int kMode, bool kFullWarp = false>
QTTT_HD StepOut step_game(State& s, uint32_t enew, bool have_coin, uint32_t coin, uint64_t seed, uint64_t game,
                          uint32_t dword, const Luts& L) {
    StepOut o;
    uint32_t was_reset = 0u, st_extra = 0u;
    if (kMode == kStepFresh) s = empty_state();
    if (kMode == kStepAuto || kMode == kStepAutoNext) {
        const uint32_t C0 = classical(s);
        if ((any_line(s, C0, L) != 0u) | (n_moves(s) >= 9u)) {               // mcts.py:52-65
            s = empty_state();
            was_reset = 1u;
        }
    }
    o.action = 255u;
    o.coin = 0u;
    if (kRandom) {
        const uint32_t C = classical(s);
        const uint32_t nm = n_moves(s);
        uint32_t act;
        policy_draw(seed, game, nm, dword, L.legal[~C & M9], act, coin);
        if (kMode == kStepPlain) {
            // terminated games are left untouched (the random policy has nothing to play)
            if ((any_line(s, C, L) != 0u) | (nm >= 9u)) { act = 255u; coin = 0u; st_extra = kStFinished; }
        }
        if (kMode == kStepAutoNext && was_reset) { act = 255u; coin = 0u; }
        enew = L.pair[act];
        o.action = act;
        o.coin = coin;
    } else {
        if (!have_coin) {
            uint32_t word;
            ply_draw(seed, game, n_moves(s), dword, word, coin);
        }
        if (kMode == kStepAutoNext && was_reset) enew = 0u;
    }
    const StepResult r = step_core<false, false, kFullWarp>(s, enew, coin, L);
    o.win = any_line(s, r.classical, L);
    o.done = (o.win != 0u) | (r.n > 8u);
    o.classical = r.classical;
    o.status = st_extra ? st_extra : ((kMode == kStepAutoNext && was_reset) ? 0u : r.illegal);
    o.status |= was_reset ? kStReset : 0u;
    o.write_state = (kMode == kStepFresh) | was_reset | (r.illegal ^ 1u);
    return o;
}

// mcts.py:52-65 terminal test on a state (winner exists or 9 entries in moves).
QTTT_HD uint32_t finished_winner(const State& s, const Luts& L, bool& terminal) {
    int px, po;
    win_rounds(s, L, px, po);
    const uint32_t w = winner_of(px, po);
    terminal = (w != 0u) | (n_moves(s) >= 9u);
    return w;
}

// One ply of MCTS._simulate (mcts.py:188-196): action ~ U(legal), coin ~ U{0,1}.
// Needs the policy tables (kLutPolicyBytes staged).
template <int N>
QTTT_HD StepResult playout_ply_fixed(State& s, uint32_t C, uint64_t seed, uint64_t game, uint32_t domain,
                                     const Luts& L, DrawCache& cache) {
    uint32_t word, coin;
    ply_draw_cached(seed, game, (uint32_t)N, domain, cache, word, coin);
    return step_fixed<N>(s, policy_edge(~C & M9, word, L), coin, C);
}
// The same for a run-time len(moves): one jump to the specialised ply (uniform across a warp whose
// lanes play from the same root; correct, if slower, when it is not).
QTTT_HD StepResult playout_ply(State& s, uint32_t C, uint64_t seed, uint64_t game, uint32_t domain,
                               const Luts& L, DrawCache& cache) {
    switch (n_moves(s)) {
        case 0: return playout_ply_fixed<0>(s, C, seed, game, domain, L, cache);
        case 1: return playout_ply_fixed<1>(s, C, seed, game, domain, L, cache);
        case 2: return playout_ply_fixed<2>(s, C, seed, game, domain, L, cache);
        case 3: return playout_ply_fixed<3>(s, C, seed, game, domain, L, cache);
        case 4: return playout_ply_fixed<4>(s, C, seed, game, domain, L, cache);
        case 5: return playout_ply_fixed<5>(s, C, seed, game, domain, L, cache);
        case 6: return playout_ply_fixed<6>(s, C, seed, game, domain, L, cache);
        case 7: return playout_ply_fixed<7>(s, C, seed, game, domain, L, cache);
        case 8: return playout_ply_fixed<8>(s, C, seed, game, domain, L, cache);
        default: break;
    }
    StepResult r;                       // 9 entries in moves: nothing can be played
    r.illegal = 1u; r.collapsed = 0u; r.classical = C; r.n = n_moves(s);
    return r;
}

QTTT_HD int board_value(uint32_t P0, uint32_t P1, uint32_t P2, uint32_t P3, int sq) {
    return (int)((P0 >> sq & 1u) | ((P1 >> sq & 1u) << 1) | ((P2 >> sq & 1u) << 2) |
                 ((P3 >> sq & 1u) << 3)) - 1;
}

// Env._observation and friends in tensor form (env.py:62-112, board.py:71-115, mcts.py:52-65,87-91).
QTTT_HD void observe_game(const State& s, const Luts& L, int8_t* classical_out, int8_t* moves,
                          uint8_t* nmoves, int8_t* q1, int8_t* q2, uint8_t* turn, int8_t* rounds,
                          float* reward_p1, uint8_t* winner, uint8_t* mask_bool, int64_t i) {
    const uint32_t P0 = plane0(s), P1 = plane1(s), P2 = plane2(s), P3 = plane3(s);
    const uint32_t C = P0 | P1 | P2 | P3;
    const uint32_t nm = n_moves(s);
    if (classical_out)
        for (int sq = 0; sq < 9; ++sq) classical_out[9 * i + sq] = (int8_t)board_value(P0, P1, P2, P3, sq);
    if (nmoves) nmoves[i] = (uint8_t)nm;
    if (turn) turn[i] = (uint8_t)(nm & 1u);                                   // env.py:83
    int c1 = 0, c2 = 0;
    if (q1) for (int k = 0; k < 10; ++k) q1[10 * i + k] = -1;
    if (q2) for (int k = 0; k < 8; ++k) q2[8 * i + k] = -1;
    for (uint32_t m = 0; m < 9; ++m) {
        const uint32_t E = edge_dyn(s, m);
        const bool present = (m < nm) & (E != 0u);
        const int a = present ? ctz32(E) : -1, b = present ? flo32(E) : -1;
        if (moves) { moves[18 * i + 2 * m] = (int8_t)a; moves[18 * i + 2 * m + 1] = (int8_t)b; }
        if (present && (E & C) == 0u) {                                       // env.py:73-78
            if (m & 1u) { if (q2 && c2 < 4) { q2[8 * i + 2 * c2] = (int8_t)a; q2[8 * i + 2 * c2 + 1] = (int8_t)b; } ++c2; }
            else        { if (q1 && c1 < 5) { q1[10 * i + 2 * c1] = (int8_t)a; q1[10 * i + 2 * c1 + 1] = (int8_t)b; } ++c1; }
        }
    }
    int px, po;
    win_rounds(s, L, px, po);
    if (rounds) { rounds[2 * i] = (int8_t)px; rounds[2 * i + 1] = (int8_t)po; }
    if (reward_p1) {                                                          // env.py:87-112
        const int a = px < 0 ? 10 : px, b = po < 0 ? 10 : po;
        reward_p1[i] = a < b ? 1.0f : (b < a ? -1.0f : 0.0f);
    }
    if (winner) winner[i] = (uint8_t)winner_of(px, po);                       // mcts.py:52-65
    if (mask_bool) {
        const uint64_t lm = L.legal[~C & M9];
        for (int k = 0; k < 36; ++k) mask_bool[36 * i + k] = (uint8_t)(lm >> k & 1ull);
    }
}

// (classical, moves, n_moves) -> packed state.
QTTT_HD State pack_game(const int8_t* classical_in, const int8_t* moves, const uint8_t* nmoves, int64_t i) {
    State s = empty_state();
    uint32_t nm = nmoves[i];
    nm = nm > 9u ? 9u : nm;
    uint32_t P3 = 0u;
    for (int sq = 0; sq < 9; ++sq) {
        const uint32_t v = (uint32_t)(classical_in[9 * i + sq] + 1) & 15u;
        s.w |= ((v & 1u) << sq) | (((v >> 1) & 1u) << (9 + sq)) | (((v >> 2) & 1u) << (18 + sq));
        P3 |= ((v >> 3) & 1u) << sq;
    }
    uint32_t wx = 0u, wy = 0u, wz = 0u;
    for (uint32_t m = 0; m < nm; ++m) {
        const uint32_t a = (uint32_t)moves[18 * i + 2 * m] & 15u, b = (uint32_t)moves[18 * i + 2 * m + 1] & 15u;
        const uint32_t E = (((1u << a) | (1u << b)) & M9) << (9u * (m % 3u));
        if (m < 3u) wx |= E; else if (m < 6u) wy |= E; else wz |= E;
    }
    s.x = wx | (nm << 27);
    s.y = wy | ((P3 & 0x1Fu) << 27);
    s.z = wz | ((P3 >> 5) << 27);
    return s;
}

QTTT_HD uint64_t board_nibbles(const State& s, const Luts& L) {
    return nibbles_of_planes<true>(s.w, plane3(s), L);
}

// Both measurement outcomes of one (position, action): board.py:42-56 + qeval.py:5-51 twice,
// i.e. what MCTS._step enumerates by rejection sampling (mcts.py:233-267).
template <bool kSquares = true>
QTTT_HD void qeval_game(const State& s, uint32_t action, const Luts& L, State* next0, State* next1,
                        uint64_t* board0, uint64_t* board1, int8_t* sq0, int8_t* sq1,
                        uint8_t* closes, float* result_prob, int64_t i) {
    const uint32_t enew = L.pair[action & 255u];
    float px_prob = 0.f, po_prob = 0.f;
    uint32_t col = 0u;
    if (!kSquares) {
        // one sweep serves both coins
        State t0, t1;
        const BothResult r = step_both(s, enew, L, t0, t1);
        col = r.collapsed;
        if (next0) next0[i] = t0;
        if (next1) next1[i] = t1;
        if (board0) board0[i] = board_nibbles(t0, L);
        if (board1) board1[i] = board_nibbles(t1, L);
        if (result_prob) {
            int px, po;
            win_rounds(t0, L, px, po);
            uint32_t wnr = winner_of(px, po);
            px_prob += wnr == 1u ? 0.5f : 0.f;
            po_prob += wnr == 2u ? 0.5f : 0.f;
            win_rounds(t1, L, px, po);
            wnr = winner_of(px, po);
            px_prob += wnr == 1u ? 0.5f : 0.f;
            po_prob += wnr == 2u ? 0.5f : 0.f;
        }
    } else {
        // eval()'s per-move return value is wanted: run the transition per coin with targets
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            State t = s;
            uint32_t tgt[9];
            const StepResult r = step_core<true>(t, enew, (uint32_t)c, L, tgt);
            col = r.collapsed;
            State* nx = c ? next1 : next0;
            uint64_t* bd = c ? board1 : board0;
            int8_t* sq = c ? sq1 : sq0;
            if (nx) nx[i] = t;
            if (bd) bd[i] = board_nibbles(t, L);
            if (sq) {
#pragma unroll
                for (int m = 0; m < 9; ++m) sq[9 * i + m] = (int8_t)(tgt[m] ? ctz32(tgt[m]) : -1);
            }
            if (result_prob) {
                int px, po;
                win_rounds(t, L, px, po);
                const uint32_t wnr = winner_of(px, po);
                px_prob += wnr == 1u ? 0.5f : 0.f;
                po_prob += wnr == 2u ? 0.5f : 0.f;
            }
        }
    }
    if (closes) closes[i] = (uint8_t)col;
    if (result_prob) {
        result_prob[3 * i] = px_prob;
        result_prob[3 * i + 1] = po_prob;
        result_prob[3 * i + 2] = 1.0f - px_prob - po_prob;
    }
}

// Squares that carry at least one live (uncollapsed) spooky mark = union of Board.qstructs.
QTTT_HD uint32_t live_squares(const State& s) {
    const uint32_t C = classical(s);
    uint32_t live = 0u;
#pragma unroll
    for (uint32_t m = 0; m < 9u; ++m) {
        const uint32_t E = edge_dyn(s, m);
        live |= (E & C) ? 0u : E;
    }
    return live;
}

// One element of GameState.to_vector() (mcts.py:67-85): an (18, 10) matrix; rows 0..8 one-hot
// of board[row] (column 9 for "not classical"), rows 9..17: 1/sqrt(9) at [square, t] for every
// move t that names the square (collapsed ones and the autofill entry included), and 1.0 in
// column 9 for squares outside every entangled component.
QTTT_HD float feature_element(const State& s, uint32_t live, uint32_t row, uint32_t col) {
    if (row < 9u) {
        const int b = board_value(plane0(s), plane1(s), plane2(s), plane3(s), (int)row);
        return ((b < 0 ? 9 : b) == (int)col) ? 1.0f : 0.0f;
    }
    const uint32_t sq = row - 9u;
    if (col == 9u) return (live >> sq & 1u) ? 0.0f : 1.0f;
    return (col < n_moves(s) && (edge_dyn(s, col) >> sq & 1u)) ? (1.0f / 3.0f) : 0.0f;
}

// Number of plies behind a position: the autofill entry (s, s, 8) is not a ply (mcts.py:243).
QTTT_HD uint32_t plies_of(const State& s) {
    const uint32_t nm = n_moves(s);
    const uint32_t e8 = (s.z >> 18) & M9;
    return nm - (uint32_t)((nm == 9u) & (popc32(e8) == 1));
}

// One uniform-random playout to a terminal state (mcts.py:185-208).  Returns the winner.
QTTT_HD uint32_t playout_game(State s, uint64_t seed, uint64_t game, uint32_t domain, const Luts& L,
                              uint32_t& steps, uint32_t& collapses) {
    uint32_t C = classical(s);
    bool terminal = (any_line(s, C, L) != 0u) | (n_moves(s) >= 9u);          // mcts.py:52-65
    DrawCache cache = empty_draw_cache();
    while (!terminal) {
        const StepResult r = playout_ply(s, C, seed, game, domain, L, cache);
        C = r.classical;
        ++steps;
        collapses += r.collapsed;
        terminal = (any_line(s, C, L) != 0u) | (r.n >= 9u);
    }
    bool t2;
    return finished_winner(s, L, t2);       // who has the earlier line: only needed at the end
}

}  // namespace qttt
