// qttt_mcts.cuh -- the reference's MCTS search (mcts.py:132-337) over the packed transition,
// one tree per root in a caller-provided node pool in HBM.
//
//   MCTS.reset            mcts.py:139-164   mcts_init_root
//   _select/_uct_select   mcts.py:269-285   mcts_select (PUCT, first maximum wins, f64)
//   _expand_child/_step   mcts.py:210-267   mcts_expand (both collapse outcomes, coin 0 first)
//   _simulate/_reward     mcts.py:185-208   playout_game (qttt_core.cuh), domain 3
//   _rollout              mcts.py:166-173   value = sum(r if leaf.turn else -r) / num_simulations
//   _backpropogate        mcts.py:175-183   mcts_backprop (sign flip per level, f64, no FMA)
//   choose / sync         mcts.py:308-337   mcts_choose / mcts_sync
//
// Every random choice is drawn from Philox4x32-10 with the keys documented in
// oracle/mcts_oracle.py, which is pinned to the unmodified reference MCTS consuming the same
// stream.  All statistics are kept exactly as the reference keeps them (N int, W double,
// Q = W / N), with round-to-nearest double arithmetic in the reference's operation order, so
// the search is bit-identical.  The reference's transposition dict is keyed by
// hash(board + moves) with `moves` the full history, so it never merges distinct tree nodes:
// the search structure is a tree and is stored as one.
#pragma once
#include "qttt_core.cuh"

namespace qttt {

constexpr uint32_t kDomainSelect = 2u;
constexpr uint32_t kDomainSim = 3u;
constexpr uint64_t kMaxSims = 4096ull;

struct alignas(16) MctsNode {
    State    state;          // the position (Board.board / Board.moves)
    uint32_t ntot;           // Ntot
    uint8_t  has_p;          // P is not None (the node has been a simulated leaf)
    uint8_t  terminal;       // mcts.py:52-65
    uint8_t  winner;         // 0 none, 1 X, 2 O
    uint8_t  turn;           // GameState.turn (True = X to move at the root of an empty game)
    uint64_t legal;          // 36-bit action mask (node.actions)
    uint32_t n[36];          // N[a]
    int32_t  child[36][2];   // children[a] (node indices; -1 = not expanded / single child)
    double   w[36];          // W[a]
};
static_assert(sizeof(MctsNode) == 752, "MctsNode layout");

// per-root bookkeeping, int32[8]
//   [0] root node   [1] nodes ever taken from the pool's fresh end (high-water mark)
//   [2] rollouts done   [3] error bits   [4] head of the free list (-1 = empty)
//   [5] length of the free list   [6] most nodes alive at once
enum { kMetaRoot = 0, kMetaCount = 1, kMetaRollouts = 2, kMetaError = 3, kMetaFree = 4, kMetaFreeCount = 5,
       kMetaPeak = 6, kMetaStride = 8 };
enum { kMctsErrPoolFull = 1, kMctsErrNoSuchChild = 2 };

QTTT_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
QTTT_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
QTTT_HD double d_div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
QTTT_HD double d_sqrt(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return __builtin_sqrt(a);
#endif
}

QTTT_HD void mcts_init_node(MctsNode& nd, const State& s, bool turn, const Luts& L) {
    nd.state = s;
    nd.ntot = 0u;
    nd.has_p = 0;
    bool terminal;
    nd.winner = (uint8_t)finished_winner(s, L, terminal);
    nd.terminal = terminal ? 1 : 0;
    nd.turn = turn ? 1 : 0;
    nd.legal = L.legal[~classical(s) & M9];
    for (int a = 0; a < 36; ++a) { nd.n[a] = 0u; nd.child[a][0] = -1; nd.child[a][1] = -1; nd.w[a] = 0.0; }
}

// mcts.py:139-164: the root's turn is len(game.moves) % 2 == 0
QTTT_HD void mcts_init_root(MctsNode* tree, int32_t* meta, const State& s, const Luts& L) {
    mcts_init_node(tree[0], s, (n_moves(s) & 1u) == 0u, L);
    meta[kMetaRoot] = 0; meta[kMetaCount] = 1; meta[kMetaRollouts] = 0; meta[kMetaError] = 0;
    meta[kMetaFree] = -1; meta[kMetaFreeCount] = 0; meta[kMetaPeak] = 1;
}

// Node allocation: reuse a node reclaimed by mcts_sync (MCTS._prune, mcts.py:222-231) when there
// is one, else take the next fresh node of the pool.  Which slot a node lives in has no influence
// on the search (children are referenced by index), so reuse keeps the statistics bit-identical.
QTTT_HD int mcts_nodes_available(const int32_t* meta, int64_t capacity) {
    const int64_t fresh = capacity - (int64_t)meta[kMetaCount];
    return (int)(fresh > 0x3FFFFFFF ? 0x3FFFFFFF : fresh) + meta[kMetaFreeCount];
}
QTTT_HD int mcts_alloc(MctsNode* tree, int32_t* meta) {
    int id;
    if (meta[kMetaFree] >= 0) {
        id = meta[kMetaFree];
        meta[kMetaFree] = tree[id].child[0][0];        // the free list is threaded through child[0][0]
        meta[kMetaFreeCount] -= 1;
    } else {
        id = meta[kMetaCount];
        meta[kMetaCount] = id + 1;
    }
    const int live = meta[kMetaCount] - meta[kMetaFreeCount];
    if (live > meta[kMetaPeak]) meta[kMetaPeak] = live;
    return id;
}

// MCTS._prune (mcts.py:222-231): give the whole subtree under `c` back to the pool.  The nodes
// still to visit are chained through `ntot` (their statistics are dead), so no stack is needed.
QTTT_HD void mcts_free_subtree(MctsNode* tree, int32_t* meta, int c) {
    int head = c;
    tree[c].ntot = 0xFFFFFFFFu;
    while (head >= 0) {
        const int v = head;
        head = (int)tree[v].ntot;
        for (int a = 0; a < 36; ++a)
            for (int k = 0; k < 2; ++k) {
                const int ch = tree[v].child[a][k];
                if (ch >= 0) { tree[ch].ntot = (uint32_t)head; head = ch; }
            }
        tree[v].child[0][0] = meta[kMetaFree];
        meta[kMetaFree] = v;
        meta[kMetaFreeCount] += 1;
    }
}

// mcts.py:280-285: argmax_a Q[a] + c_puct * P[a] * sqrt(Ntot) / (1 + N[a]), first maximum wins
QTTT_HD int mcts_uct_select(const MctsNode& nd, double c_puct) {
    const uint32_t lo = (uint32_t)nd.legal, hi = (uint32_t)(nd.legal >> 32);
    const int m = popc32(lo) + popc32(hi);
    const double p = d_div(1.0, (double)m);
    const double ps = d_mul(p, d_sqrt((double)nd.ntot));
    int best = -1;
    double best_v = 0.0;
    for (int a = 0; a < 36; ++a) {
        if (!(nd.legal >> a & 1ull)) continue;
        const double u = d_div(ps, (double)(1u + nd.n[a]));
        const double q = nd.n[a] ? d_div(nd.w[a], (double)nd.n[a]) : 0.0;
        const double v = d_add(q, d_mul(c_puct, u));
        if (best < 0 || v > best_v) { best = a; best_v = v; }
    }
    return best;
}

// mcts.py:210-220 + 233-267: children[a] = [outcome of coin 0, outcome of coin 1] (one child when
// the move closes no cycle).  Returns false when the pool is exhausted.
QTTT_HD bool mcts_expand(MctsNode* tree, int32_t* meta, int64_t capacity, int node, int a, const Luts& L) {
    const uint32_t enew = L.pair[a];
    State s0 = tree[node].state, s1 = s0;
    const StepResult r0 = step_core(s0, enew, 0u, L);
    const int need = r0.collapsed ? 2 : 1;
    if (mcts_nodes_available(meta, capacity) < need) { meta[kMetaError] |= kMctsErrPoolFull; return false; }
    const bool turn = !tree[node].turn;
    const int c0 = mcts_alloc(tree, meta);
    mcts_init_node(tree[c0], s0, turn, L);
    tree[node].child[a][0] = c0;
    if (r0.collapsed) {
        step_core(s1, enew, 1u, L);
        const int c1 = mcts_alloc(tree, meta);
        mcts_init_node(tree[c1], s1, turn, L);
        tree[node].child[a][1] = c1;
    }
    return true;
}

// mcts.py:269-277.  path_node/path_act: the (node, action) pairs walked; returns the leaf.
QTTT_HD int mcts_select(MctsNode* tree, int32_t* meta, int64_t capacity, uint64_t seed, uint64_t base,
                        double c_puct, const Luts& L, int* path_node, int* path_act, int& depth) {
    int node = meta[kMetaRoot];
    depth = 0;
    while (tree[node].has_p && !tree[node].terminal) {
        const int a = mcts_uct_select(tree[node], c_puct);
        if (tree[node].child[a][0] < 0 && !mcts_expand(tree, meta, capacity, node, a, L)) break;
        path_node[depth] = node;
        path_act[depth] = a;
        uint32_t c0 = (uint32_t)base, c1 = (uint32_t)(base >> 32), c2 = (uint32_t)depth, c3 = kDomainSelect;
        philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
        const int second = tree[node].child[a][1];
        node = (second >= 0 && (c1 & 1u)) ? second : tree[node].child[a][0];
        ++depth;
    }
    return node;
}

// reward of one playout for the leaf's side to move: r if leaf.turn else -r  (mcts.py:171, 200-208)
QTTT_HD int mcts_sim_reward(const MctsNode& leaf, uint64_t seed, uint64_t base, uint32_t sim, const Luts& L) {
    uint32_t steps = 0, cols = 0;
    const uint32_t w = playout_game(leaf.state, seed, base * kMaxSims + sim, kDomainSim, L, steps, cols);
    const int r = w == 1u ? 1 : (w == 2u ? -1 : 0);
    return leaf.turn ? r : -r;
}

// mcts.py:175-183 with r = r_tot / num_simulations
QTTT_HD void mcts_backprop(MctsNode* tree, const int* path_node, const int* path_act, int depth,
                           int r_tot, int num_sims) {
    double r = d_div((double)r_tot, (double)num_sims);
    for (int d = depth - 1; d >= 0; --d) {
        MctsNode& nd = tree[path_node[d]];
        const int a = path_act[d];
        r = -r;
        nd.w[a] = d_add(nd.w[a], r);
        nd.n[a] += 1u;
        nd.ntot += 1u;
    }
}

// mcts.py:308-315: argmax over root actions of Q[a] (N[a] > 0) else -inf, first maximum wins
QTTT_HD int mcts_choose(const MctsNode& root) {
    int best = -1;
    double best_v = 0.0;
    bool best_inf = true;
    for (int a = 0; a < 36; ++a) {
        if (!(root.legal >> a & 1ull)) continue;
        const bool inf = root.n[a] == 0u;
        const double v = inf ? 0.0 : d_div(root.w[a], (double)root.n[a]);
        // compare (inf ? -infinity : v) > best
        const bool better = best < 0 || (!inf && (best_inf || v > best_v));
        if (better) { best = a; best_v = v; best_inf = inf; }
    }
    return best < 0 ? 255 : best;
}

// mcts.py:317-337: the root moves to the child of `action` whose position is `now`
QTTT_HD void mcts_sync(MctsNode* tree, int32_t* meta, int64_t capacity, int action, const State& now,
                       const Luts& L) {
    const int root = meta[kMetaRoot];
    if (action >= 36) return;            // "no move was played" (a finished game's filler): keep the root
    if (action < 0 || !(tree[root].legal >> action & 1ull)) { meta[kMetaError] |= kMctsErrNoSuchChild; return; }
    if (tree[root].child[action][0] < 0 && !mcts_expand(tree, meta, capacity, root, action, L)) return;
    int keep = -1;
    for (int k = 0; k < 2; ++k) {
        const int c = tree[root].child[action][k];
        if (c < 0) continue;
        const State& s = tree[c].state;
        if (s.x == now.x && s.y == now.y && s.z == now.z && s.w == now.w) { keep = c; break; }
    }
    if (keep < 0) { meta[kMetaError] |= kMctsErrNoSuchChild; return; }
    // mcts.py:330-337: every other child of the old root is pruned with its subtree, the old
    // root itself is dropped; their nodes go back to the pool
    for (int a = 0; a < 36; ++a)
        for (int k = 0; k < 2; ++k) {
            const int c = tree[root].child[a][k];
            if (c >= 0 && c != keep) mcts_free_subtree(tree, meta, c);
            tree[root].child[a][k] = -1;
        }
    mcts_free_subtree(tree, meta, root);
    meta[kMetaRoot] = keep;
}

}  // namespace qttt
