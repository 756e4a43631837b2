// TEST INFRASTRUCTURE ONLY -- runs the per-game functions of qtttgym_b200/csrc/qttt_core.cuh
// on the CPU with the same signatures as the C ABI (include/qttt_b200.h) minus streams, so
// that the transition logic can be checked against the oracle in the build container (which
// has no GPU) before GPU time is spent.  Never shipped, never loaded by qtttgym_b200: the
// product path has no CPU fallback.
#include <stdint.h>
#include <string.h>

#include "../../qtttgym_b200/csrc/qttt_core.cuh"
#include "../../qtttgym_b200/csrc/qttt_mcts.cuh"

using namespace qttt;

static const LutImage g_img = make_lut_image();
static Luts luts() { return luts_from_image(&g_img); }

// The per-game body is qttt_core.cuh's step_game -- the very function the step kernels call.
template <bool kRandom, int kMode>
static void emu_step_mode(State* state, const void* action, int fmt, const uint8_t* coin, uint64_t seed,
                          uint64_t game_base, uint64_t epoch, uint8_t* action_out, uint8_t* coin_out,
                          float* reward, uint8_t* done, uint64_t* mask, uint8_t* status, int64_t n) {
    const Luts L = luts();
    const uint8_t* act = static_cast<const uint8_t*>(action);
    const uint32_t dword = domain_word(0u, epoch);
    for (int64_t i = 0; i < n; ++i) {
        State s = state[i];
        uint32_t enew = 0u;
        if (!kRandom) enew = fmt == 0 ? (uint32_t)L.pair[act[i]] : pair_to_edge(act[2 * i], act[2 * i + 1]);
        const StepOut o = step_game<kRandom, kMode>(s, enew, coin != nullptr, coin ? (coin[i] & 1u) : 0u, seed,
                                                    game_base + (uint64_t)i, dword, L);
        if (o.write_state) state[i] = s;
        if (reward) reinterpret_cast<uint32_t*>(reward)[i] = reward_bits(o.win);
        if (done) done[i] = (uint8_t)o.done;
        if (mask) mask[i] = L.legal[~o.classical & M9];
        if (status) status[i] = (uint8_t)o.status;
        if (kRandom && action_out) action_out[i] = (uint8_t)o.action;
        if (kRandom && coin_out) coin_out[i] = (uint8_t)o.coin;
    }
}

template <bool kRandom>
static int emu_step_flags(uint32_t flags, State* state, const void* action, int fmt, const uint8_t* coin,
                          uint64_t seed, uint64_t game_base, uint64_t epoch, uint8_t* action_out,
                          uint8_t* coin_out, float* reward, uint8_t* done, uint64_t* mask, uint8_t* status,
                          int64_t n) {
    if (flags == 0u) emu_step_mode<kRandom, kStepPlain>(state, action, fmt, coin, seed, game_base, epoch, action_out, coin_out, reward, done, mask, status, n);
    else if (flags == 1u) emu_step_mode<kRandom, kStepFresh>(state, action, fmt, coin, seed, game_base, epoch, action_out, coin_out, reward, done, mask, status, n);
    else if (flags == 2u) emu_step_mode<kRandom, kStepAuto>(state, action, fmt, coin, seed, game_base, epoch, action_out, coin_out, reward, done, mask, status, n);
    else if (flags == 4u) emu_step_mode<kRandom, kStepAutoNext>(state, action, fmt, coin, seed, game_base, epoch, action_out, coin_out, reward, done, mask, status, n);
    else return -1;
    return 0;
}

extern "C" {

int emu_reset(State* state, uint64_t* mask, int64_t n) {
    for (int64_t i = 0; i < n; ++i) { state[i] = empty_state(); if (mask) mask[i] = 0xFFFFFFFFFull; }
    return 0;
}

int emu_step(State* state, const void* action, int fmt, const uint8_t* coin, uint64_t seed,
             uint64_t game_base, float* reward, uint8_t* done, uint64_t* mask, uint8_t* status,
             int64_t n) {
    return emu_step_flags<false>(0u, state, action, fmt, coin, seed, game_base, 0, nullptr, nullptr, reward, done, mask, status, n);
}

int emu_step_ex(State* state, const void* action, int fmt, const uint8_t* coin, uint64_t seed,
                uint64_t game_base, uint64_t epoch, uint32_t flags, float* reward, uint8_t* done,
                uint64_t* mask, uint8_t* status, int64_t n) {
    return emu_step_flags<false>(flags, state, action, fmt, coin, seed, game_base, epoch, nullptr, nullptr, reward, done, mask, status, n);
}

int emu_step_packed(State* state, const uint8_t* action_coin, uint16_t* result, int64_t n) {
    const Luts L = luts();
    for (int64_t i = 0; i < n; ++i) {
        State s = state[i];
        const uint32_t ac = action_coin[i];
        const StepResult r = step_core(s, (uint32_t)L.pair[ac & 63u], ac >> 7, L);
        if (!r.illegal) state[i] = s;
        const uint32_t win = any_line(s, r.classical, L) != 0u;
        const uint32_t term = win | (uint32_t)(r.n > 8u);
        result[i] = (uint16_t)((~r.classical & M9) | (term << 9) | (win << 10) | (r.illegal << 11));
    }
    return 0;
}

int emu_step_random(State* state, uint64_t seed, uint64_t game_base, uint8_t* action_out,
                    uint8_t* coin_out, float* reward, uint8_t* done, uint64_t* mask,
                    uint8_t* status, int64_t n) {
    return emu_step_flags<true>(0u, state, nullptr, 0, nullptr, seed, game_base, 0, action_out, coin_out, reward, done, mask, status, n);
}

int emu_step_random_ex(State* state, uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags,
                       uint8_t* action_out, uint8_t* coin_out, float* reward, uint8_t* done,
                       uint64_t* mask, uint8_t* status, int64_t n) {
    return emu_step_flags<true>(flags, state, nullptr, 0, nullptr, seed, game_base, epoch, action_out, coin_out, reward, done, mask, status, n);
}

int emu_observe(const State* state, int8_t* classical_out, int8_t* moves, uint8_t* nmoves,
                int8_t* q1, int8_t* q2, uint8_t* turn, int8_t* rounds, float* reward_p1,
                uint8_t* winner, uint8_t* mask_bool, int64_t n) {
    const Luts L = luts();
    for (int64_t i = 0; i < n; ++i)
        observe_game(state[i], L, classical_out, moves, nmoves, q1, q2, turn, rounds, reward_p1,
                     winner, mask_bool, i);
    return 0;
}

int emu_features(const State* state, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t live = live_squares(state[i]);
        for (uint32_t e = 0; e < 180; ++e) out[180 * i + e] = feature_element(state[i], live, e / 10u, e % 10u);
    }
    return 0;
}

int emu_get_mask(const State* state, uint8_t* illegal_mask, int64_t n) {
    const Luts L = luts();
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t legal = L.legal[~classical(state[i]) & M9];
        for (int k = 0; k < 36; ++k) illegal_mask[36 * i + k] = (uint8_t)(((legal >> k) & 1ull) ^ 1ull);
    }
    return 0;
}

int emu_pack(State* state, const int8_t* classical_in, const int8_t* moves, const uint8_t* nmoves,
             int64_t n) {
    for (int64_t i = 0; i < n; ++i) state[i] = pack_game(classical_in, moves, nmoves, i);
    return 0;
}

int emu_qeval_both(const State* state, const uint8_t* action, State* next0, State* next1,
                   uint64_t* board0, uint64_t* board1, int8_t* sq0, int8_t* sq1, uint8_t* closes,
                   float* result_prob, int64_t n) {
    const Luts L = luts();
    // the same dispatch as qttt_qeval_both: boards only / per-move squares wanted or not
    if (board0 && board1 && closes && !next0 && !next1 && !sq0 && !sq1 && !result_prob) {
        for (int64_t i = 0; i < n; ++i) {
            const BoardsBoth r = boards_both(state[i], (uint32_t)L.pair[action[i]], L);
            board0[i] = r.board0;
            board1[i] = r.board1;
            closes[i] = (uint8_t)r.collapsed;
        }
        return 0;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (sq0 || sq1)
            qeval_game<true>(state[i], action[i], L, next0, next1, board0, board1, sq0, sq1, closes,
                             result_prob, i);
        else
            qeval_game<false>(state[i], action[i], L, next0, next1, board0, board1, sq0, sq1, closes,
                              result_prob, i);
    }
    return 0;
}

int emu_rollout(const State* roots, int64_t n_roots, int32_t n_rollouts, uint64_t seed,
                int32_t* tallies, float* value, int64_t* steps_total) {
    const Luts L = luts();
    for (int64_t r = 0; r < n_roots; ++r) {
        int t[3] = {0, 0, 0};
        uint32_t steps = 0, cols = 0;
        for (int32_t j = 0; j < n_rollouts; ++j) {
            const uint32_t w = playout_game(roots[r], seed, (uint64_t)r * (uint64_t)n_rollouts + (uint64_t)j, 1u, L, steps, cols);
            t[w == 1u ? 0 : (w == 2u ? 1 : 2)]++;
        }
        if (tallies) { tallies[3 * r] = t[0]; tallies[3 * r + 1] = t[1]; tallies[3 * r + 2] = t[2]; }
        if (value) {
            const float v = (float)(t[0] - t[1]) / (float)n_rollouts;
            value[r] = (plies_of(roots[r]) & 1u) ? -v : v;
        }
        if (steps_total) *steps_total += steps;
    }
    return 0;
}

int emu_sweep(int64_t lo, int64_t hi, uint64_t seed, int64_t* stats) {
    const Luts L = luts();
    for (int64_t g = lo; g < hi; ++g) {
        uint32_t steps = 0, cols = 0;
        const uint32_t w = playout_game(empty_state(), seed, (uint64_t)g, 0u, L, steps, cols);
        stats[w == 1u ? 0 : (w == 2u ? 1 : 2)]++;
        stats[3] += steps; stats[4] += cols; stats[5]++; stats[6 + steps]++;
    }
    return 0;
}

int emu_mcts_node_bytes(void) { return (int)sizeof(MctsNode); }

int emu_mcts_init(MctsNode* pool, int64_t capacity, int32_t* meta, const State* roots, int64_t n_roots) {
    const Luts L = luts();
    for (int64_t r = 0; r < n_roots; ++r) mcts_init_root(pool + r * capacity, meta + r * kMetaStride, roots[r], L);
    return 0;
}

int emu_mcts_run(MctsNode* pool, int64_t capacity, int32_t* meta, int32_t n_rollouts, int32_t num_sims,
                 double c_puct, uint64_t seed, uint64_t root_base, int64_t n_roots) {
    const Luts L = luts();
    for (int64_t root = 0; root < n_roots; ++root) {
        MctsNode* tree = pool + root * capacity;
        int32_t* m = meta + root * kMetaStride;
        const uint64_t root_id = (root_base + (uint64_t)root) << 32;
        const int first = m[kMetaRollouts];
        for (int it = 0; it < n_rollouts; ++it) {
            const uint64_t base = root_id + (uint64_t)(uint32_t)(first + it);
            int path_node[12], path_act[12], depth;
            const int leaf = mcts_select(tree, m, capacity, seed, base, c_puct, L, path_node, path_act, depth);
            int r_tot = 0;
            for (int sim = 0; sim < num_sims; ++sim) r_tot += mcts_sim_reward(tree[leaf], seed, base, (uint32_t)sim, L);
            if (!tree[leaf].terminal) tree[leaf].has_p = 1;
            mcts_backprop(tree, path_node, path_act, depth, r_tot, num_sims);
        }
        m[kMetaRollouts] = first + n_rollouts;
    }
    return 0;
}

int emu_mcts_stats(const MctsNode* pool, int64_t capacity, const int32_t* meta, int32_t* n_out,
                   double* q_out, int32_t* ntot_out, uint8_t* choose_out, int64_t n_roots) {
    for (int64_t r = 0; r < n_roots; ++r) {
        const MctsNode& root = pool[r * capacity + meta[r * kMetaStride + kMetaRoot]];
        for (int a = 0; a < 36; ++a) {
            if (n_out) n_out[36 * r + a] = (int32_t)root.n[a];
            if (q_out) q_out[36 * r + a] = root.n[a] ? root.w[a] / (double)root.n[a] : 0.0;
        }
        if (ntot_out) ntot_out[r] = (int32_t)root.ntot;
        if (choose_out) choose_out[r] = (uint8_t)mcts_choose(root);
    }
    return 0;
}

int emu_mcts_sync(MctsNode* pool, int64_t capacity, int32_t* meta, const uint8_t* action,
                  const State* now, int64_t n_roots) {
    const Luts L = luts();
    for (int64_t r = 0; r < n_roots; ++r)
        mcts_sync(pool + r * capacity, meta + r * kMetaStride, capacity, (int)action[r], now[r], L);
    return 0;
}

}  // extern "C"
