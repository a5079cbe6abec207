/*
 * xq_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never on the product path).
 *
 * Plain-C restatement of the Xiangqi rules arithmetic that the reference
 * (Symb0x76/ChineseChessZero) obtains from the un-vendored, un-pinned third-party
 * package `cchess` (windshadow233/python-chinese-chess; README.md:21 of the reference;
 * no version pin exists in requirements.txt:1-5).  The package source is NOT available
 * in the authoring container, so this file restates the *published rules of Xiangqi*
 * plus the python-chess-lineage conventions recalled in SURVEY.md App. A.
 *
 * PARITY STATUS: the legal-move SET, check and stalemate logic are pinned by the widely
 * published start-position perft values 44 / 1,920 / 79,666 / 3,290,240 / 133,312,995
 * (tests/test_oracle_rules.py).  The generation ORDER, the half-move-clock convention,
 * is_insufficient_material and the outcome() ordering are recollections: "parity
 * unpinned" for those, and each is a named policy constant below.
 *
 * Reference call sites this file stands in for (all into /root/reference):
 *   board.legal_moves      net.py:154-157      -> xq_legal_moves
 *   board.push / .copy     mcts.py:111,151; game.py:201 -> xq_push / xq_game_push
 *   board.piece_at         tools.py:92 (decode_board tools.py:74-106) -> xq_decode_board
 *   board.is_game_over / outcome / is_tie   mcts.py:116-126; tools.py:109-123;
 *                          game.py:208-216     -> xq_flags / xq_game_flags
 *   policy_value_fn input  net.py:160-177      -> xq_encode_search_planes
 *   move-id table          tools.py:172-272    -> xq_build_action_table
 *
 * Layout conventions (identical to the device board record, DESIGN.md "Data layout"):
 *   square = file + 9*rank, rank 0 = Red's back rank (tools.py:91);
 *   piece code = type | 8*(colour==BLACK); types PAWN=1 CANNON=2 ROOK=3 KNIGHT=4
 *   BISHOP=5 ADVISOR=6 KING=7 (channel = type-1, tools.py:96-100); 0 = empty.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XQ_PAWN 1
#define XQ_CANNON 2
#define XQ_ROOK 3
#define XQ_KNIGHT 4
#define XQ_BISHOP 5
#define XQ_ADVISOR 6
#define XQ_KING 7
#define XQ_BLACKBIT 8

#define XQ_FLAG_CHECK 1
#define XQ_FLAG_NOMOVES 2
#define XQ_FLAG_INSUFFICIENT 4
#define XQ_FLAG_FOURFOLD 8
#define XQ_FLAG_SIXTY 16

/* policy knobs (recalled behaviour, kept switchable) */
#define XQ_SIXTY_CLOCK 120 /* is_sixty_moves: halfmove_clock >= 120 and a legal move exists */
#define XQ_FOURFOLD_PRIOR 3 /* is_repetition(4): 3 earlier occurrences + now */

typedef struct {
    uint8_t sq[90];
    uint8_t turn;  /* 1 = RED to move, 0 = BLACK (cchess.RED == True) */
    uint8_t clock; /* half-move clock: plies since the last capture */
    uint8_t rep;   /* earlier occurrences of this (placement, turn) in the reversible window */
    uint8_t pad[3];
} xq_board; /* 96 bytes */

static inline int is_red(uint8_t c) { return c != 0 && !(c & XQ_BLACKBIT); }
static inline int is_black(uint8_t c) { return (c & XQ_BLACKBIT) != 0; }
static inline int colour_of(uint8_t c) { return is_red(c) ? 1 : 0; } /* only for c != 0 */
static inline int type_of(uint8_t c) { return c & 7; }
static inline int own(uint8_t c, int red_side) { return c != 0 && (is_red(c) == (red_side != 0)); }

void xq_start(xq_board *b) {
    static const uint8_t back[9] = {XQ_ROOK, XQ_KNIGHT, XQ_BISHOP, XQ_ADVISOR, XQ_KING,
                                    XQ_ADVISOR, XQ_BISHOP, XQ_KNIGHT, XQ_ROOK};
    memset(b, 0, sizeof(*b));
    for (int f = 0; f < 9; ++f) {
        b->sq[f] = back[f];
        b->sq[81 + f] = back[f] | XQ_BLACKBIT;
    }
    b->sq[2 * 9 + 1] = b->sq[2 * 9 + 7] = XQ_CANNON;
    b->sq[7 * 9 + 1] = b->sq[7 * 9 + 7] = XQ_CANNON | XQ_BLACKBIT;
    for (int f = 0; f < 9; f += 2) {
        b->sq[3 * 9 + f] = XQ_PAWN;
        b->sq[6 * 9 + f] = XQ_PAWN | XQ_BLACKBIT;
    }
    b->turn = 1;
}

/* ---- pseudo-legal destinations of the piece standing on `from` ---------------- */
static void piece_targets(const uint8_t *sq, int from, uint8_t *dest /*[90] 0/1*/) {
    const uint8_t pc = sq[from];
    const int red = is_red(pc);
    const int r = from / 9, f = from % 9;
    static const int dr4[4] = {1, -1, 0, 0}, df4[4] = {0, 0, 1, -1};
    memset(dest, 0, 90);
    switch (type_of(pc)) {
    case XQ_ROOK:
        for (int d = 0; d < 4; ++d) {
            int rr = r + dr4[d], ff = f + df4[d];
            while (rr >= 0 && rr < 10 && ff >= 0 && ff < 9) {
                uint8_t t = sq[rr * 9 + ff];
                if (t == 0) dest[rr * 9 + ff] = 1;
                else { if (!own(t, red)) dest[rr * 9 + ff] = 1; break; }
                rr += dr4[d]; ff += df4[d];
            }
        }
        break;
    case XQ_CANNON:
        for (int d = 0; d < 4; ++d) {
            int rr = r + dr4[d], ff = f + df4[d];
            int screen = 0;
            while (rr >= 0 && rr < 10 && ff >= 0 && ff < 9) {
                uint8_t t = sq[rr * 9 + ff];
                if (!screen) {
                    if (t == 0) dest[rr * 9 + ff] = 1; else screen = 1;
                } else if (t != 0) {
                    if (!own(t, red)) dest[rr * 9 + ff] = 1;
                    break;
                }
                rr += dr4[d]; ff += df4[d];
            }
        }
        break;
    case XQ_KNIGHT: {
        static const int kr[8] = {2, 2, -2, -2, 1, 1, -1, -1};
        static const int kf[8] = {1, -1, 1, -1, 2, -2, 2, -2};
        for (int i = 0; i < 8; ++i) {
            int rr = r + kr[i], ff = f + kf[i];
            if (rr < 0 || rr > 9 || ff < 0 || ff > 8) continue;
            int lr = r + (kr[i] == 2 ? 1 : kr[i] == -2 ? -1 : 0);
            int lf = f + (kf[i] == 2 ? 1 : kf[i] == -2 ? -1 : 0);
            if (sq[lr * 9 + lf] != 0) continue; /* hobbled horse */
            if (!own(sq[rr * 9 + ff], red)) dest[rr * 9 + ff] = 1;
        }
        break;
    }
    case XQ_BISHOP:
        for (int a = -2; a <= 2; a += 4)
            for (int c = -2; c <= 2; c += 4) {
                int rr = r + a, ff = f + c;
                if (rr < 0 || rr > 9 || ff < 0 || ff > 8) continue;
                if (red ? rr > 4 : rr < 5) continue;            /* river */
                if (sq[(r + a / 2) * 9 + f + c / 2] != 0) continue; /* blocked eye */
                if (!own(sq[rr * 9 + ff], red)) dest[rr * 9 + ff] = 1;
            }
        break;
    case XQ_ADVISOR:
        for (int a = -1; a <= 1; a += 2)
            for (int c = -1; c <= 1; c += 2) {
                int rr = r + a, ff = f + c;
                if (ff < 3 || ff > 5) continue;
                if (red ? (rr < 0 || rr > 2) : (rr < 7 || rr > 9)) continue;
                if (!own(sq[rr * 9 + ff], red)) dest[rr * 9 + ff] = 1;
            }
        break;
    case XQ_KING:
        for (int d = 0; d < 4; ++d) {
            int rr = r + dr4[d], ff = f + df4[d];
            if (ff < 3 || ff > 5) continue;
            if (red ? (rr < 0 || rr > 2) : (rr < 7 || rr > 9)) continue;
            if (!own(sq[rr * 9 + ff], red)) dest[rr * 9 + ff] = 1;
        }
        break;
    case XQ_PAWN: {
        int fr = r + (red ? 1 : -1);
        if (fr >= 0 && fr <= 9 && !own(sq[fr * 9 + f], red)) dest[fr * 9 + f] = 1;
        if (red ? r >= 5 : r <= 4) { /* crossed the river: may step sideways */
            if (f > 0 && !own(sq[r * 9 + f - 1], red)) dest[r * 9 + f - 1] = 1;
            if (f < 8 && !own(sq[r * 9 + f + 1], red)) dest[r * 9 + f + 1] = 1;
        }
        break;
    }
    default: break;
    }
}

/* Is `target` attacked by colour `by_red` (rook, cannon, knight, pawn, and the enemy king
 * looking down an open file -- the flying-general rule)? */
static int attacked(const uint8_t *sq, int target, int by_red) {
    const int r = target / 9, f = target % 9;
    static const int dr4[4] = {1, -1, 0, 0}, df4[4] = {0, 0, 1, -1};
    for (int d = 0; d < 4; ++d) {
        int rr = r + dr4[d], ff = f + df4[d], seen = 0, dist = 1;
        while (rr >= 0 && rr < 10 && ff >= 0 && ff < 9) {
            uint8_t t = sq[rr * 9 + ff];
            if (t != 0) {
                int theirs = own(t, by_red), ty = type_of(t);
                if (!seen) {
                    if (theirs) {
                        if (ty == XQ_ROOK) return 1;
                        if (ty == XQ_KING && df4[d] == 0) return 1;
                        if (ty == XQ_PAWN && dist == 1) {
                            /* pawn on (rr,ff) attacks forward, and sideways once across the river */
                            int fwd = by_red ? 1 : -1;
                            if (df4[d] == 0) { if (rr + fwd == r) return 1; }
                            else if (by_red ? rr >= 5 : rr <= 4) return 1;
                        }
                    }
                    seen = 1;
                } else {
                    if (theirs && ty == XQ_CANNON) return 1;
                    break;
                }
            }
            rr += dr4[d]; ff += df4[d]; ++dist;
        }
    }
    static const int kr[8] = {2, 2, -2, -2, 1, 1, -1, -1};
    static const int kf[8] = {1, -1, 1, -1, 2, -2, 2, -2};
    for (int i = 0; i < 8; ++i) {
        int nr = r + kr[i], nf = f + kf[i];
        if (nr < 0 || nr > 9 || nf < 0 || nf > 8) continue;
        uint8_t t = sq[nr * 9 + nf];
        if (type_of(t) != XQ_KNIGHT || !own(t, by_red)) continue;
        /* the knight travels (-kr,-kf); its leg is next to the knight in the long direction */
        int lr = nr - (kr[i] == 2 ? 1 : kr[i] == -2 ? -1 : 0);
        int lf = nf - (kf[i] == 2 ? 1 : kf[i] == -2 ? -1 : 0);
        if (sq[lr * 9 + lf] == 0) return 1;
    }
    return 0;
}

static int king_square(const uint8_t *sq, int red) {
    const uint8_t k = XQ_KING | (red ? 0 : XQ_BLACKBIT);
    const int lo = red ? 0 : 63, hi = red ? 27 : 90;
    for (int s = lo; s < hi; ++s) if (sq[s] == k) return s;
    for (int s = 0; s < 90; ++s) if (sq[s] == k) return s;
    return -1;
}

int xq_in_check(const xq_board *b) {
    int k = king_square(b->sq, b->turn);
    return k >= 0 && attacked(b->sq, k, !b->turn);
}

/* Generation-order policy, the mirror of ccz_order_policy (include/ccz_b200.h): legal moves are
 * ordered by (class_rank[piece type], from-square key, capture key, to-square key).  The default is
 * the recalled cchess order: non-pawn pieces by from-square DESCENDING, destinations
 * DESCENDING, then the pawns likewise.  scripts/pin_cchess.py derives the policy from a real cchess. */
typedef struct {
    uint8_t class_rank[8]; /* index = piece type 1..7 */
    uint8_t from_descending, to_descending, capture_mode, check_king_first;
} xq_order_policy;
static const xq_order_policy XQ_DEFAULT_POLICY = {{0, 1, 0, 0, 0, 0, 0, 0}, 1, 1, 0, 0};
static xq_order_policy g_policy = {{0, 1, 0, 0, 0, 0, 0, 0}, 1, 1, 0, 0};

int xq_set_order_policy(const uint8_t *p /* 12 bytes or NULL = default */) {
    xq_order_policy q = XQ_DEFAULT_POLICY;
    if (p) memcpy(&q, p, sizeof(q));
    if (q.capture_mode > 2 || q.from_descending > 1 || q.to_descending > 1 || q.check_king_first > 1) return -1;
    q.class_rank[0] = 0;
    g_policy = q;
    return 0;
}
void xq_get_order_policy(uint8_t *out /* 12 bytes */) { memcpy(out, &g_policy, sizeof(g_policy)); }

static uint32_t order_key(const uint8_t *sq, int from, int to, int in_check) {
    uint32_t cls = g_policy.class_rank[type_of(sq[from])];
    if (g_policy.check_king_first && in_check) cls = type_of(sq[from]) == XQ_KING ? 0u : cls + 1u; /* evasions: king first */
    const uint32_t fk = g_policy.from_descending ? 89 - from : from, tk = g_policy.to_descending ? 89 - to : to;
    const uint32_t ck = g_policy.capture_mode == 0 ? 0u : (uint32_t)((sq[to] != 0) != (g_policy.capture_mode == 2));
    return cls << 16 | fk << 9 | ck << 8 | tk;
}

static int xq_legal_moves_native(const xq_board *b, uint16_t *moves);

/* board.legal_moves (net.py:155) in the policy's order.  moves[i] = from<<8 | to.  Returns the count. */
int xq_legal_moves(const xq_board *b, uint16_t *moves) {
    const int n = xq_legal_moves_native(b, moves);
    if (memcmp(&g_policy, &XQ_DEFAULT_POLICY, sizeof(g_policy)) != 0) {
        uint32_t keys[128];
        const int in_check = xq_in_check(b);
        for (int i = 0; i < n; ++i) keys[i] = order_key(b->sq, moves[i] >> 8, moves[i] & 255, in_check);
        for (int i = 1; i < n; ++i) { /* insertion sort: n <= 119, keys distinct */
            const uint32_t k = keys[i];
            const uint16_t m = moves[i];
            int j = i - 1;
            for (; j >= 0 && keys[j] > k; --j) { keys[j + 1] = keys[j]; moves[j + 1] = moves[j]; }
            keys[j + 1] = k;
            moves[j + 1] = m;
        }
    }
    return n;
}

/* the default order generated directly: non-pawn pieces by from-square DESCENDING, destinations
 * DESCENDING, then pawns likewise */
static int xq_legal_moves_native(const xq_board *b, uint16_t *moves) {
    uint8_t dest[90], tmp[90];
    int n = 0;
    const int red = b->turn;
    for (int pass = 0; pass < 2; ++pass) {
        for (int from = 89; from >= 0; --from) {
            uint8_t pc = b->sq[from];
            if (!own(pc, red)) continue;
            if ((type_of(pc) == XQ_PAWN) != (pass == 1)) continue;
            piece_targets(b->sq, from, dest);
            for (int to = 89; to >= 0; --to) {
                if (!dest[to]) continue;
                memcpy(tmp, b->sq, 90);
                tmp[to] = pc; tmp[from] = 0;
                int k = type_of(pc) == XQ_KING ? to : king_square(tmp, red);
                if (k >= 0 && attacked(tmp, k, !red)) continue;
                moves[n++] = (uint16_t)(from << 8 | to);
            }
        }
    }
    return n;
}

/* board.push (mcts.py:111; game.py:201) on a bare position: capture resets the clock.
 * `rep` cannot be maintained without history -- see xq_game_push. */
void xq_push(xq_board *b, int from, int to) {
    uint8_t cap = b->sq[to];
    b->sq[to] = b->sq[from];
    b->sq[from] = 0;
    b->clock = cap ? 0 : (uint8_t)(b->clock < 255 ? b->clock + 1 : 255);
    b->turn ^= 1;
    b->rep = 0;
}

/* terminal predicates (tools.py:109-123; mcts.py:116-126; game.py:208-216) as a flag byte */
int xq_flags(const xq_board *b, int n_legal) {
    int fl = 0;
    if (xq_in_check(b)) fl |= XQ_FLAG_CHECK;
    if (n_legal == 0) fl |= XQ_FLAG_NOMOVES;
    int attackers = 0;
    for (int s = 0; s < 90; ++s) {
        int ty = type_of(b->sq[s]);
        if (b->sq[s] && (ty == XQ_PAWN || ty == XQ_CANNON || ty == XQ_ROOK || ty == XQ_KNIGHT)) attackers = 1;
    }
    if (!attackers) fl |= XQ_FLAG_INSUFFICIENT;
    if (b->rep >= XQ_FOURFOLD_PRIOR) fl |= XQ_FLAG_FOURFOLD;
    if (b->clock >= XQ_SIXTY_CLOCK && n_legal > 0) fl |= XQ_FLAG_SIXTY;
    return fl;
}

/* decode_board (tools.py:74-106): two int8 (7,10,9) one-hot arrays */
void xq_decode_board(const xq_board *b, int8_t *red /*[630]*/, int8_t *black /*[630]*/) {
    memset(red, 0, 630); memset(black, 0, 630);
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 9; ++j) {
            uint8_t pc = b->sq[j + i * 9];
            if (!pc) continue;
            int ch = type_of(pc) - 1;
            (is_red(pc) ? red : black)[ch * 90 + i * 9 + j] = 1;
        }
}

/* search-time net input (net.py:160-177): 7 zero states + current per side, turn plane.
 * out = 17*7*10*9 bf16 bit patterns (0x3F80 = 1.0). */
void xq_encode_search_planes(const xq_board *b, uint16_t *out /*[10710]*/) {
    int8_t red[630], black[630];
    xq_decode_board(b, red, black);
    memset(out, 0, 10710 * sizeof(uint16_t));
    for (int i = 0; i < 630; ++i) {
        if (red[i]) out[7 * 630 + i] = 0x3F80;
        if (black[i]) out[15 * 630 + i] = 0x3F80;
        if (b->turn) out[16 * 630 + i] = 0x3F80;
    }
}

/* ---- action table (tools.py:172-272) ----------------------------------------- */
static int16_t g_id_of[90 * 90];
static uint8_t g_from_of[2086], g_to_of[2086];
static int g_table_built = 0;

static int sq_from_uci(const char *s) { return (s[0] - 'a') + 9 * (s[1] - '0'); }

int xq_build_action_table(int16_t *id_of /*[8100] or NULL*/, uint8_t *from_of, uint8_t *to_of) {
    if (!g_table_built) {
        static const char *adv[16] = {"d0e1", "e1d0", "f0e1", "e1f0", "d2e1", "e1d2", "f2e1", "e1f2",
                                      "d9e8", "e8d9", "f9e8", "e8f9", "d7e8", "e8d7", "f7e8", "e8f7"};
        static const char *bis[32] = {
            "a2c0", "c0a2", "a2c4", "c4a2", "c0e2", "e2c0", "c4e2", "e2c4", "e2g0", "g0e2", "e2g4",
            "g4e2", "g0i2", "i2g0", "g4i2", "i2g4", "a7c5", "c5a7", "a7c9", "c9a7", "c5e7", "e7c5",
            "c9e7", "e7c9", "e7g5", "g5e7", "e7g9", "g9e7", "g5i7", "i7g5", "g9i7", "i7g9"};
        static const int ka[8] = {-2, -1, -2, 1, 2, -1, 2, 1}, kb[8] = {-1, -2, 1, -2, -1, 2, 1, 2};
        int idx = 0;
        for (int i = 0; i < 8100; ++i) g_id_of[i] = -1;
        for (int l1 = 0; l1 < 10; ++l1)
            for (int n1 = 0; n1 < 9; ++n1) {
                int dl[27], dn[27], m = 0;
                for (int t = 0; t < 10; ++t) { dl[m] = t; dn[m++] = n1; }
                for (int t = 0; t < 9; ++t) { dl[m] = l1; dn[m++] = t; }
                for (int k = 0; k < 8; ++k) { dl[m] = l1 + ka[k]; dn[m++] = n1 + kb[k]; }
                for (int k = 0; k < m; ++k) {
                    if (dl[k] == l1 && dn[k] == n1) continue;
                    if (dl[k] < 0 || dl[k] > 9 || dn[k] < 0 || dn[k] > 8) continue;
                    int from = n1 + 9 * l1, to = dn[k] + 9 * dl[k];
                    g_id_of[from * 90 + to] = (int16_t)idx; /* later duplicates overwrite, as the dict does */
                    g_from_of[idx] = (uint8_t)from; g_to_of[idx] = (uint8_t)to;
                    ++idx;
                }
            }
        for (int k = 0; k < 16; ++k, ++idx) {
            int from = sq_from_uci(adv[k]), to = sq_from_uci(adv[k] + 2);
            g_id_of[from * 90 + to] = (int16_t)idx; g_from_of[idx] = from; g_to_of[idx] = to;
        }
        for (int k = 0; k < 32; ++k, ++idx) {
            int from = sq_from_uci(bis[k]), to = sq_from_uci(bis[k] + 2);
            g_id_of[from * 90 + to] = (int16_t)idx; g_from_of[idx] = from; g_to_of[idx] = to;
        }
        g_table_built = idx;
    }
    if (id_of) memcpy(id_of, g_id_of, sizeof(g_id_of));
    if (from_of) memcpy(from_of, g_from_of, 2086);
    if (to_of) memcpy(to_of, g_to_of, 2086);
    return g_table_built;
}

/* ---- batch form used by parity tests and the cpu_baseline leg ------------------ */
void xq_batch_movegen_encode(const xq_board *boards, int n, int16_t *move_ids /*[n,128]*/,
                             int16_t *counts, uint8_t *flags, uint16_t *planes /*[n,10710] or NULL*/) {
    uint16_t mv[128];
    xq_build_action_table(NULL, NULL, NULL);
    for (int i = 0; i < n; ++i) {
        int c = xq_legal_moves(&boards[i], mv);
        for (int k = 0; k < 128; ++k)
            move_ids[(size_t)i * 128 + k] = k < c ? g_id_of[(mv[k] >> 8) * 90 + (mv[k] & 255)] : -1;
        counts[i] = (int16_t)c;
        flags[i] = (uint8_t)xq_flags(&boards[i], c);
        if (planes) xq_encode_search_planes(&boards[i], planes + (size_t)i * 10710);
    }
}

/* ---- perft and leaf collection (golden vectors: 44 / 1920 / 79666 / 3290240 / 133312995) */
uint64_t xq_perft(const xq_board *b, int depth) {
    uint16_t mv[128];
    int n = xq_legal_moves(b, mv);
    if (depth <= 1) return (uint64_t)n;
    uint64_t tot = 0;
    for (int i = 0; i < n; ++i) {
        xq_board c = *b;
        xq_push(&c, mv[i] >> 8, mv[i] & 255);
        tot += xq_perft(&c, depth - 1);
    }
    return tot;
}

static void collect_rec(const xq_board *b, int depth, xq_board *out, int64_t cap, int64_t *n) {
    if (depth == 0) { if (*n < cap) out[*n] = *b; ++*n; return; }
    uint16_t mv[128];
    int c = xq_legal_moves(b, mv);
    for (int i = 0; i < c; ++i) {
        xq_board t = *b;
        xq_push(&t, mv[i] >> 8, mv[i] & 255);
        collect_rec(&t, depth - 1, out, cap, n);
    }
}
/* all positions exactly `depth` plies from `root`, in generation order */
int64_t xq_collect_leaves(const xq_board *root, int depth, xq_board *out, int64_t cap) {
    int64_t n = 0;
    collect_rec(root, depth, out, cap, &n);
    return n;
}

/* ---- game with a move stack: exact repetition, as is_repetition(4) walks it ------ */
#define XQ_MAX_PLY 4096
typedef struct {
    xq_board cur;
    int ply;
    uint8_t from[XQ_MAX_PLY], to[XQ_MAX_PLY], captured[XQ_MAX_PLY], prev_clock[XQ_MAX_PLY];
} xq_game;

int xq_game_sizeof(void) { return (int)sizeof(xq_game); }
void xq_game_init(xq_game *g, const xq_board *b) {
    g->cur = *b; g->ply = 0;
    g->cur.rep = 0;
}

/* python-chess is_repetition: pop moves while they are reversible (non-capture), counting
 * positions whose (placement, turn) equals the current one; the position *before* an
 * irreversible move is never compared.  Exact comparison, no hashing. */
static int count_prior_occurrences(const xq_game *g) {
    uint8_t w[90];
    memcpy(w, g->cur.sq, 90);
    int turn = g->cur.turn, cnt = 0;
    for (int p = g->ply - 1; p >= 0; --p) {
        if (g->captured[p]) break;
        w[g->from[p]] = w[g->to[p]];
        w[g->to[p]] = 0;
        turn ^= 1;
        if (turn == g->cur.turn && memcmp(w, g->cur.sq, 90) == 0) ++cnt;
    }
    return cnt;
}

int xq_game_push(xq_game *g, int from, int to) {
    if (g->ply >= XQ_MAX_PLY) return -1;
    int p = g->ply++;
    g->from[p] = (uint8_t)from; g->to[p] = (uint8_t)to;
    g->captured[p] = g->cur.sq[to]; g->prev_clock[p] = g->cur.clock;
    xq_push(&g->cur, from, to);
    int c = count_prior_occurrences(g);
    g->cur.rep = (uint8_t)(c > 255 ? 255 : c);
    return 0;
}

int xq_game_pop(xq_game *g) {
    if (g->ply <= 0) return -1;
    int p = --g->ply;
    g->cur.sq[g->from[p]] = g->cur.sq[g->to[p]];
    g->cur.sq[g->to[p]] = g->captured[p];
    g->cur.clock = g->prev_clock[p];
    g->cur.turn ^= 1;
    int c = count_prior_occurrences(g);
    g->cur.rep = (uint8_t)(c > 255 ? 255 : c);
    return 0;
}

void xq_game_copy(xq_game *dst, const xq_game *src) {
    dst->cur = src->cur; dst->ply = src->ply;
    memcpy(dst->from, src->from, src->ply);
    memcpy(dst->to, src->to, src->ply);
    memcpy(dst->captured, src->captured, src->ply);
    memcpy(dst->prev_clock, src->prev_clock, src->ply);
}
const xq_board *xq_game_board(const xq_game *g) { return &g->cur; }
int xq_game_ply(const xq_game *g) { return g->ply; }
