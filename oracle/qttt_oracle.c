/* TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of the qtttgym game-transition path.
 *
 * This file is the bulk *oracle* (checker) for the CUDA path: the same algorithm as
 * oracle/qttt_oracle.py (which see for the pinning story), in C so that the 1e6-game
 * equivalence suite and the CPU baseline finish in seconds.  It restates the reference
 * (Oxel40/qtttgym) function by function; citations are into /root/reference.  It is pinned
 * against the live reference through tests/test_oracle_vs_reference.py (build container) and
 * the recorded traces in tests/golden/ (everywhere).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the library built from this file.  The product never links or calls it.
 *
 * Build: python oracle/build.py   ->  oracle/_build/libqttt_oracle.so   (gcc -O2 -fopenmp)
 */
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* One game, reference-shaped (board.py:2-7).  40 bytes, no padding surprises. */
typedef struct {
    int8_t   board[9];    /* -1 = not classical, else owning move index        board.py:5  */
    int8_t   n_moves;     /* len(moves)                                                    */
    int8_t   mv[9][2];    /* moves[i] = (a, b, i), a < b; autofill (s, s, i)   board.py:19,25 */
    int8_t   n_comps;     /* len(qstructs)                                                 */
    int8_t   pad;
    uint16_t comp[5];     /* qstructs as 9-bit square sets, list order kept    board.py:6  */
} orc_game;

static const int8_t LINES[8][3] = {
    {0, 1, 2}, {3, 4, 5}, {6, 7, 8},   /* rows      board.py:85-90   */
    {0, 3, 6}, {1, 4, 7}, {2, 5, 8},   /* columns   board.py:93-98   */
    {2, 4, 6},                         /* anti-diag board.py:101-105 */
    {0, 4, 8},                         /* diag      board.py:106-110 */
};

static int8_t PAIR[36][2];
static int pair_ready = 0;
static void init_pairs(void) {          /* mcts.py:339-343 as a table */
    if (pair_ready) return;
    int k = 0;
    for (int i = 0; i < 9; ++i)
        for (int j = i + 1; j < 9; ++j) { PAIR[k][0] = (int8_t)i; PAIR[k][1] = (int8_t)j; ++k; }
    pair_ready = 1;
}

void orc_reset(orc_game *g) {           /* board.py:2-7 / env.py:55-57 */
    memset(g, 0, sizeof *g);
    for (int s = 0; s < 9; ++s) g->board[s] = -1;
    for (int i = 0; i < 9; ++i) g->mv[i][0] = g->mv[i][1] = -1;
}

static inline void drop_incident(int *list, int *deg, int p) {   /* set.remove, qeval.py:30,40,48 */
    int d = *deg;
    for (int i = 0; i < d; ++i)
        if (list[i] == p) { list[i] = list[d - 1]; break; }
    *deg = d - 1;
}

/* qeval.py:5-51.  mem[] = positions (move indices) of the component's moves in idx order,
 * the last one closed the cycle.  where[i] = square mem[i] collapses into. */
static void measure(const orc_game *g, const int *mem, int k, int coin, int *where) {
    int inc[9][9], deg[9];              /* incidence lists per square (qeval.py:12-19) */
    for (int s = 0; s < 9; ++s) deg[s] = 0;
    for (int p = 0; p < k; ++p) {
        where[p] = -1;
        int a = g->mv[mem[p]][0], b = g->mv[mem[p]][1];
        inc[a][deg[a]++] = p;
        inc[b][deg[b]++] = p;
    }
#define DROP(sq, p) drop_incident(inc[sq], &deg[sq], (p))
    /* qeval.py:23-31: peel pendant squares inwards */
    for (int start = 0; start < 9; ++start) {
        int sq = start;
        while (deg[sq] == 1) {
            int p = inc[sq][0];
            deg[sq] = 0;
            int a = g->mv[mem[p]][0], b = g->mv[mem[p]][1];
            int inner = (sq == a) ? b : a;
            where[p] = sq;
            DROP(inner, p);
            sq = inner;
        }
    }
    /* qeval.py:35-49: coin for the closing move, then walk the cycle hi -> lo */
    int last = k - 1;
    int lo = g->mv[mem[last]][0], hi = g->mv[mem[last]][1];
    where[last] = coin ? hi : lo;
    int sq = hi;
    int occupied = (where[last] == sq);
    DROP(sq, last);
    while (sq != lo) {
        int p = inc[sq][0];
        deg[sq] = 0;
        int a = g->mv[mem[p]][0], b = g->mv[mem[p]][1];
        int other = (b == sq) ? a : b;
        where[p] = occupied ? other : sq;
        DROP(other, p);
        occupied = (where[p] == other);
        sq = other;
    }
#undef DROP
}

/* board.py:9-25 + 27-69.  returns 1 = illegal (no-op, env.py:36-43), 0 = ok;
 * *collapsed set when a measurement happened (the coin was consumed). */
int orc_place(orc_game *g, int a, int b, int coin, int *collapsed) {
    if (collapsed) *collapsed = 0;
    if (a < 0 || a > 8 || b < 0 || b > 8) return 1;     /* IndexError path / out of domain */
    if (a == b) return 1;                               /* board.py:10-12 */
    if (g->board[a] != -1 || g->board[b] != -1) return 1;   /* board.py:14-15 */
    if (a > b) { int t = a; a = b; b = t; }             /* board.py:16-18 */
    int idx = g->n_moves;
    g->mv[idx][0] = (int8_t)a; g->mv[idx][1] = (int8_t)b;   /* board.py:19 */
    g->n_moves = (int8_t)(idx + 1);

    int ia = -1, ib = -2;                               /* board.py:28-40 */
    for (int i = 0; i < g->n_comps; ++i) if (g->comp[i] >> a & 1) { ia = i; break; }
    for (int i = 0; i < g->n_comps; ++i) if (g->comp[i] >> b & 1) { ib = i; break; }
    if (ia == ib) {                                     /* board.py:42-56 */
        int mem[9], where[9], k = 0;
        for (int i = 0; i < g->n_moves; ++i)
            if (g->comp[ia] >> g->mv[i][0] & 1) mem[k++] = i;
        measure(g, mem, k, coin & 1, where);
        for (int p = 0; p < k; ++p) g->board[where[p]] = (int8_t)mem[p];
        for (int i = ib; i + 1 < g->n_comps; ++i) g->comp[i] = g->comp[i + 1];
        g->n_comps--;
        if (collapsed) *collapsed = 1;
    } else if (ia >= 0 && ib >= 0) {                    /* board.py:58-61 */
        g->comp[ia] |= g->comp[ib];
        for (int i = ib; i + 1 < g->n_comps; ++i) g->comp[i] = g->comp[i + 1];
        g->n_comps--;
    } else {                                            /* board.py:62-69 */
        int i = ia > ib ? ia : ib;
        if (i < 0) { i = g->n_comps++; g->comp[i] = 0; }
        g->comp[i] |= (uint16_t)((1u << a) | (1u << b));
    }
    int free_cnt = 0, free_sq = -1;                     /* board.py:21-25 */
    for (int s = 0; s < 9; ++s) if (g->board[s] == -1) { ++free_cnt; free_sq = s; }
    if (free_cnt == 1) {
        int n = g->n_moves;
        g->board[free_sq] = (int8_t)n;
        g->mv[n][0] = g->mv[n][1] = (int8_t)free_sq;
        g->n_moves = (int8_t)(n + 1);
    }
    return 0;
}

/* board.py:71-115 */
void orc_win_rounds(const orc_game *g, int *px, int *po) {
    int x = 10, o = 10;
    for (int l = 0; l < 8; ++l) {
        int s = 0, mx = -1;
        for (int j = 0; j < 3; ++j) {
            int v = g->board[LINES[l][j]];
            if (v >= 0) s += (v & 1) * 2 - 1;
            if (v > mx) mx = v;
        }
        if (s == -3 && mx < x) x = mx;
        else if (s == 3 && mx < o) o = mx;
    }
    *px = x < 10 ? x : -1;
    *po = o < 10 ? o : -1;
}

uint64_t orc_legal_mask(const orc_game *g) {            /* mcts.py:19-27, 87-91 */
    init_pairs();
    uint64_t m = 0;
    for (int k = 0; k < 36; ++k)
        if (g->board[PAIR[k][0]] == -1 && g->board[PAIR[k][1]] == -1) m |= 1ull << k;
    return m;
}

int orc_winner(const orc_game *g) {                     /* mcts.py:52-65: 1 X, 2 O, 0 none */
    int px, po;
    orc_win_rounds(g, &px, &po);
    if (px > 0 && po > 0) return px < po ? 1 : 2;
    if (px > 0) return 1;
    if (po > 0) return 2;
    return 0;
}

float orc_reward_p1(const orc_game *g) {                /* env.py:87-112 */
    int px, po;
    orc_win_rounds(g, &px, &po);
    if (px < 0) px = 10;
    if (po < 0) po = 10;
    if (px < po) return 1.0f;
    if (po < px) return -1.0f;
    return 0.0f;
}

/* ------------------------------------------------------------------ batched env.step
 * env.py:34-53 for n independent games.  action: (a, b) int8 pairs.  Outputs may be NULL. */
void orc_step_batch(orc_game *games, int64_t n, const int8_t *action, const uint8_t *coin,
                    float *reward, uint8_t *done, uint64_t *mask, uint8_t *status,
                    uint8_t *collapsed_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        orc_game *g = &games[i];
        int col = 0;
        int st = orc_place(g, action[2 * i], action[2 * i + 1], coin ? coin[i] & 1 : 0, &col);
        int px, po;
        orc_win_rounds(g, &px, &po);
        int win = px > 0 || po > 0;
        if (reward) reward[i] = -1.0f * (float)win;     /* env.py:49 -> -0.0f / -1.0f */
        if (done) done[i] = (uint8_t)(win || g->n_moves > 8);   /* env.py:51 */
        if (mask) mask[i] = orc_legal_mask(g);
        if (status) status[i] = (uint8_t)st;
        if (collapsed_out) collapsed_out[i] = (uint8_t)col;
    }
}

void orc_reset_batch(orc_game *games, int64_t n) {
    for (int64_t i = 0; i < n; ++i) orc_reset(&games[i]);
}

/* env.py:68-85 in tensor form.  classical i8[n,9]; moves i8[n,9,2] (pad -1); n_moves u8[n];
 * q1 i8[n,5,2], q2 i8[n,4,2] (pad -1); turn u8[n]; plus a10/a13 derived values. */
void orc_observe_batch(const orc_game *games, int64_t n, int8_t *classical, int8_t *moves,
                       uint8_t *n_moves, int8_t *q1, int8_t *q2, uint8_t *turn,
                       int8_t *rounds, float *reward_p1, uint8_t *winner) {
    for (int64_t i = 0; i < n; ++i) {
        const orc_game *g = &games[i];
        if (classical) memcpy(classical + 9 * i, g->board, 9);
        if (moves) memcpy(moves + 18 * i, g->mv, 18);
        if (n_moves) n_moves[i] = (uint8_t)g->n_moves;
        if (turn) turn[i] = (uint8_t)(g->n_moves & 1);
        if (q1 || q2) {
            int used = 0, c1 = 0, c2 = 0;
            for (int s = 0; s < 9; ++s) if (g->board[s] >= 0) used |= 1 << g->board[s];
            if (q1) memset(q1 + 10 * i, -1, 10);
            if (q2) memset(q2 + 8 * i, -1, 8);
            for (int m = 0; m < g->n_moves; ++m) {
                if (used >> m & 1) continue;            /* env.py:74 */
                if (m & 1) { if (q2 && c2 < 4) { q2[8 * i + 2 * c2] = g->mv[m][0]; q2[8 * i + 2 * c2 + 1] = g->mv[m][1]; } ++c2; }
                else       { if (q1 && c1 < 5) { q1[10 * i + 2 * c1] = g->mv[m][0]; q1[10 * i + 2 * c1 + 1] = g->mv[m][1]; } ++c1; }
            }
        }
        if (rounds) { int px, po; orc_win_rounds(g, &px, &po); rounds[2 * i] = (int8_t)px; rounds[2 * i + 1] = (int8_t)po; }
        if (reward_p1) reward_p1[i] = orc_reward_p1(g);
        if (winner) winner[i] = (uint8_t)orc_winner(g);
    }
}

/* Build games from reference-shaped arrays (roots reached elsewhere); components are
 * re-derived from (board, moves) -- cf. SURVEY R1: the reference's MCTS.reset forgets them. */
void orc_from_arrays(orc_game *games, int64_t n, const int8_t *classical, const int8_t *moves,
                     const uint8_t *n_moves) {
    for (int64_t i = 0; i < n; ++i) {
        orc_game *g = &games[i];
        orc_reset(g);
        memcpy(g->board, classical + 9 * i, 9);
        g->n_moves = (int8_t)n_moves[i];
        for (int m = 0; m < g->n_moves; ++m) { g->mv[m][0] = moves[18 * i + 2 * m]; g->mv[m][1] = moves[18 * i + 2 * m + 1]; }
        for (int m = 0; m < g->n_moves; ++m) {
            int a = g->mv[m][0], b = g->mv[m][1];
            if (g->board[a] != -1) continue;            /* collapsed / autofill move */
            int ia = -1, ib = -1;
            for (int c = 0; c < g->n_comps; ++c) { if (g->comp[c] >> a & 1) ia = c; if (g->comp[c] >> b & 1) ib = c; }
            if (ia >= 0 && ib >= 0 && ia != ib) {
                g->comp[ia] |= g->comp[ib];
                for (int c = ib; c + 1 < g->n_comps; ++c) g->comp[c] = g->comp[c + 1];
                g->n_comps--;
            } else {
                int c = ia > ib ? ia : ib;
                if (c < 0) { c = g->n_comps++; g->comp[c] = 0; }
                g->comp[c] |= (uint16_t)((1u << a) | (1u << b));
            }
        }
    }
}

/* ------------------------------------------------------------------ Philox4x32-10
 * (Salmon et al. SC'11).  Not in the reference: defines the random-policy stream of the
 * CUDA path so that its rollouts can be replayed here. */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof c);
}

static inline int nth_set(uint64_t m, int k) {
    for (int p = 0; p < 64; ++p) if (m >> p & 1) { if (k == 0) return p; --k; }
    return -1;
}

/* The random draw of one ply: one Philox block, counter (game_lo, game_hi, ply >> 1, domain),
 * key = seed, serves two consecutive plies: even ply -> (x0, x1 & 1), odd ply -> (x2, x3 & 1). */
static inline void ply_draw(uint64_t seed, uint64_t gid, uint32_t ply, uint32_t domain,
                            uint32_t *word, uint32_t *coin) {
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), ply >> 1, domain};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    *word = (ply & 1u) ? c[2] : c[0];
    *coin = ((ply & 1u) ? c[3] : c[1]) & 1u;
}

/* mcts.py:185-208 (_simulate) with the Philox policy: action = floor(word*m/2^32)-th legal
 * action and the coin from ply_draw(seed, game, len(moves), domain).
 * Terminal per mcts.py:52-65.  Returns winner; *steps / *cols incremented. */
static int playout(orc_game *g, uint64_t seed, uint64_t gid, uint32_t domain,
                   int64_t *steps, int64_t *cols) {
    init_pairs();
    for (;;) {
        int w = orc_winner(g);
        if (w != 0 || g->n_moves == 9) return w;
        uint32_t word, coin;
        ply_draw(seed, gid, (uint32_t)g->n_moves, domain, &word, &coin);
        uint64_t mask = orc_legal_mask(g);
        int m = __builtin_popcountll(mask);
        int act = nth_set(mask, (int)(((uint64_t)word * (uint64_t)m) >> 32));
        int col = 0;
        orc_place(g, PAIR[act][0], PAIR[act][1], (int)coin, &col);
        ++*steps; *cols += col;
    }
}

/* config 5: games [lo, hi) from the empty board; stats = {X wins, O wins, draws, steps,
 * collapses, games} (strat_eval.py:66-94 tally convention).  hist9 (optional): games by
 * number of env-steps 0..9. */
void orc_selfplay(int64_t lo, int64_t hi, uint64_t seed, int64_t stats[6], int64_t *hist10) {
    int64_t xw = 0, ow = 0, dr = 0, st = 0, co = 0;
    int64_t h[10] = {0};
    init_pairs();
#pragma omp parallel for schedule(static) reduction(+ : xw, ow, dr, st, co) reduction(+ : h[:10])
    for (int64_t gid = lo; gid < hi; ++gid) {
        orc_game g;
        orc_reset(&g);
        int64_t s = 0, c = 0;
        int w = playout(&g, seed, (uint64_t)gid, 0u, &s, &c);
        xw += (w == 1); ow += (w == 2); dr += (w == 0);
        st += s; co += c;
        h[s]++;
    }
    stats[0] = xw; stats[1] = ow; stats[2] = dr; stats[3] = st; stats[4] = co; stats[5] = hi - lo;
    if (hist10) memcpy(hist10, h, sizeof h);
}

/* config 4: for each root, n_rollouts playouts with game id root*n_rollouts + r, domain 1.
 * tallies i32[n_roots,3] = (X wins, O wins, draws); steps_out (optional) total env-steps. */
void orc_rollout(const orc_game *roots, int64_t n_roots, int32_t n_rollouts, uint64_t seed,
                 int32_t *tallies, int64_t *steps_out) {
    int64_t st = 0;
    init_pairs();
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : st)
    for (int64_t r = 0; r < n_roots; ++r) {
        int32_t t[3] = {0, 0, 0};
        for (int32_t j = 0; j < n_rollouts; ++j) {
            orc_game g = roots[r];
            int64_t s = 0, c = 0;
            int w = playout(&g, seed, (uint64_t)r * (uint64_t)n_rollouts + (uint64_t)j, 1u, &s, &c);
            t[w == 1 ? 0 : (w == 2 ? 1 : 2)]++;
            st += s;
        }
        memcpy(tallies + 3 * r, t, sizeof t);
    }
    if (steps_out) *steps_out = st;
}

/* one playout with its (action, coin) trace, for replay tests */
int orc_playout_trace(orc_game *g, uint64_t seed, uint64_t gid, uint32_t domain,
                      uint8_t *acts, uint8_t *coins, int *n_out) {
    init_pairs();
    int n = 0;
    for (;;) {
        int w = orc_winner(g);
        if (w != 0 || g->n_moves == 9) { *n_out = n; return w; }
        uint32_t word, coin;
        ply_draw(seed, gid, (uint32_t)g->n_moves, domain, &word, &coin);
        uint64_t mask = orc_legal_mask(g);
        int m = __builtin_popcountll(mask);
        int act = nth_set(mask, (int)(((uint64_t)word * (uint64_t)m) >> 32));
        acts[n] = (uint8_t)act; coins[n] = (uint8_t)coin; ++n;
        orc_place(g, PAIR[act][0], PAIR[act][1], (int)coin, 0);
    }
}

/* config 3 (a5 / a14): for each game and cycle-closing action, both measurement outcomes.
 * out0/out1: the post-move boards for coin 0 / 1, packed 4 bits per square (value+1), as
 * u64; closes[i] = 1 when the action really closes a cycle (else both = plain placement).
 * sq0/sq1 (optional) i8[n,9]: square each move index collapses into (-1 = not in this
 * measurement) -- the literal return value of QEvalClassic.eval scattered by move index. */
void orc_qeval_both(const orc_game *games, int64_t n, const uint8_t *action, uint64_t *out0,
                    uint64_t *out1, uint8_t *closes, int8_t *sq0, int8_t *sq1) {
    init_pairs();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        for (int coin = 0; coin < 2; ++coin) {
            orc_game g = games[i];
            int col = 0;
            int8_t before[9];
            memcpy(before, g.board, 9);
            int n_before = g.n_moves;
            int st = action[i] < 36 ? orc_place(&g, PAIR[action[i]][0], PAIR[action[i]][1], coin, &col) : 1;
            uint64_t packed = 0;
            for (int s = 0; s < 9; ++s) packed |= (uint64_t)(g.board[s] + 1) << (4 * s);
            (coin ? out1 : out0)[i] = packed;
            if (closes) closes[i] = (uint8_t)(col && !st);
            int8_t *sq = coin ? sq1 : sq0;
            if (sq) {
                memset(sq + 9 * i, -1, 9);
                if (col) for (int s = 0; s < 9; ++s)
                    if (g.board[s] != before[s] && g.board[s] <= n_before) sq[9 * i + g.board[s]] = (int8_t)s;
            }
        }
    }
}

int orc_sizeof_game(void) { return (int)sizeof(orc_game); }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
