// ccz_movegen.cuh -- K1: fused legal-move generation + terminal flags + bf16 plane encoding.
//
// Replaces, per position:
//   [move_action2move_id[Move.uci(m)] for m in board.legal_moves]      net.py:154-157
//   decode_board + 7 zero states + current + turn plane                tools.py:74-106, net.py:160-177
//   is_game_over / is_tie predicates                                   tools.py:109-123, mcts.py:116
//
// Work decomposition: one warp owns a QUAD of 4 consecutive positions.  4 x 10710 bf16 =
// 85,680 B is the smallest run of whole positions that is 16-byte aligned, so the plane region
// of a quad is streamed with full, aligned 16 B vector stores (512 B per warp store instruction)
// over compile-time ranges (zeros / turn fill), followed by <= 32 two-byte stores per position
// that drop the piece ones in.  Move generation is warp-cooperative and branch-light:
//   * lanes own WORK ITEMS in cchess generation order (one per step piece, four -- one per ray --
//     per rook/cannon); rays are resolved with bit scans on per-rank / per-file occupancy words,
//     step pieces walk a (piece, square) -> (target, blocking square) table in shared memory;
//   * legality (own king attacked / facing the enemy king after the move) is tested one
//     candidate per lane with the same occupancy words: the first two pieces on each of the four
//     king rays after the move, plus the knights that currently attack the king's neighbourhood.
// HBM traffic per position = 96 B in, 256+2+1 B of move data and 21,420 B of planes out.
#pragma once
#include "ccz_rules.cuh"

namespace ccz {

#ifdef CCZ_V_NOINLINE
#define CCZ_MG_INLINE __noinline__
#else
#define CCZ_MG_INLINE __forceinline__
#endif
#ifdef CCZ_V_NOUNROLL
#define CCZ_Q_UNROLL _Pragma("unroll 1")
#else
#define CCZ_Q_UNROLL
#endif
#ifndef CCZ_MG_MIN_BLOCKS
#define CCZ_MG_MIN_BLOCKS 4
#endif
constexpr int MG_WARPS = 8;
constexpr int WORDS_PER_POS = PLANE_ELEMS / 2; // 5355 32-bit words (two bf16) per position
// Step-piece table: knights [90][8] (colour independent), then {pawn, elephant, advisor, king} x
// {red, black} [8][90][4]; entries target DESCENDING: to | block_square << 7 (127 = none), 0xFFFF = end
constexpr int STEP_KNIGHT_ENTRIES = 90 * 8;
constexpr int STEP_TAB_ENTRIES = STEP_KNIGHT_ENTRIES + 8 * 90 * 4 + 8; // + 8 pad: the scan reads 8 slots
__device__ __align__(16) uint16_t d_step_tab[STEP_TAB_ENTRIES];
// quad-claim counters of the dynamic scheduler (one slot per in-flight launch, zeroed in-stream)
constexpr int MG_COUNTER_SLOTS = 16;
__device__ unsigned int d_mg_counter[MG_COUNTER_SLOTS];

struct __align__(16) MgWarpSmem {
    uint8_t boards[4][BOARD_BYTES]; // 384
    uint16_t mv[MAX_MOVES];         // pseudo-legal candidates, from<<8|to, generation order
    int16_t row[MAX_MOVES];         // legal action ids, generation order
    uint16_t items[32];             // work items: square | dir<<8 (dir 0..3 = slider ray, 4 = step piece)
    uint16_t occR[10];              // occupancy of each rank (9 bits)
    uint16_t occF[10];              // occupancy of each file (10 bits)
    int16_t cnt[4];
    uint8_t flg[4];
    uint8_t pad[4];
};

__device__ __forceinline__ bool enemy_piece(uint32_t c, bool red) { return c != 0u && ((c & 8u) != 0u) == red; }

// squares at distance 1.. along a line, as seen from bit `pos`: bit j = square at distance j+1
__device__ __forceinline__ uint32_t ray_bits(uint32_t line, int pos, bool positive) {
    if (positive) return line >> (pos + 1);
    return pos ? __brev(line << (32 - pos)) : 0u;
}

// knight jump i (0..7): square delta king->knight, and delta knight->its leg for the jump back
__device__ __forceinline__ void knight_geom(int i, int &dr, int &df, int &dleg) {
    dr = (i < 4) ? ((i & 2) ? -2 : 2) : ((i & 2) ? -1 : 1);
    df = (i < 4) ? ((i & 1) ? -1 : 1) : ((i & 1) ? -2 : 2);
    dleg = (i < 4) ? (dr > 0 ? -9 : 9) : (df > 0 ? -1 : 1);
}

// Is the king on `K` attacked by colour !red after the mover (side `red`) plays f->t with piece pc?
// f == t == -1 tests the position as it stands.  `km` = knight-threat mask of the CURRENT king
// square (bit i: an enemy knight sits on jump i); king moves re-derive it for the new square.
__device__ __forceinline__ bool attacked_after_move(const MgWarpSmem &w, const uint8_t *B, int f, int t, uint32_t pc,
                                                    int K, bool king_move, uint32_t km, bool red) {
    const int kr = K / 9, kf = K - 9 * kr;
    uint32_t rk = w.occR[kr], fl = w.occF[kf];
    if (f >= 0) {
        const int fr = f / 9, ff = f - 9 * fr, tr = t / 9, tf = t - 9 * tr;
        if (fr == kr) rk &= ~(1u << ff);
        if (ff == kf) fl &= ~(1u << fr);
        if (tr == kr) rk |= 1u << tf;
        if (tf == kf) fl |= 1u << tr;
    }
    bool att = false;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const bool vertical = d == 0 || d == 3, positive = d < 2;
        const int step = d == 0 ? 9 : d == 1 ? 1 : d == 2 ? -1 : -9;
        const uint32_t x = ray_bits(vertical ? fl : rk, vertical ? kr : kf, positive);
        const int d1 = __ffs(x);
        if (d1) {
            const int s1 = K + step * d1;
            const uint32_t c1 = s1 == t ? pc : (uint32_t)B[s1];
            if (enemy_piece(c1, red)) {
                const int ty = c1 & 7;
                bool a = ty == ROOK || (ty == KING && vertical);
                if (ty == PAWN && d1 == 1) {
                    // an enemy pawn attacks forward (towards our side) and sideways once across the river
                    if (vertical) a = a || (red ? d == 0 : d == 3);
                    else a = a || (red ? kr <= 4 : kr >= 5);
                }
                att = att || a;
            }
            const uint32_t x2 = x & (x - 1u);
            const int d2 = __ffs(x2);
            if (d2) {
                const int s2 = K + step * d2;
                const uint32_t c2 = s2 == t ? pc : (uint32_t)B[s2];
                att = att || (enemy_piece(c2, red) && (c2 & 7) == CANNON);
            }
        }
    }
    if (!king_move) {
        while (km) {
            const int i = __ffs(km) - 1;
            km &= km - 1u;
            int dr, df, dleg;
            knight_geom(i, dr, df, dleg);
            const int n = K + 9 * dr + df, leg = n + dleg;
            const bool leg_empty = (B[leg] == 0 || leg == f) && leg != t;
            att = att || (n != t && leg_empty);
        }
    } else {
        const uint32_t kn = KNIGHT | (red ? BLACKBIT : 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int dr, df, dleg;
            knight_geom(i, dr, df, dleg);
            const int nr = kr + dr, nf = kf + df;
            if (nr < 0 || nr > 9 || nf < 0 || nf > 8) continue;
            const int n = nr * 9 + nf, leg = n + dleg;
            if (B[n] != kn) continue;
            att = att || B[leg] == 0 || leg == f;
        }
    }
    return att;
}

// One position, one warp.  Writes w.row / returns count and flag byte (uniform across lanes).
__device__ CCZ_MG_INLINE void movegen_one(MgWarpSmem &w, const uint8_t *B, const int16_t *s_id_of,
                                            const uint16_t *s_step, int lane, int &n_legal_out, int &flags_out) {
    const bool red = B[OFF_TURN] != 0;
    const uint32_t c0 = B[lane], c1 = B[lane + 32], c2 = lane < 26 ? B[lane + 64] : 0u;
    const int t0 = c0 & 7, t1 = c1 & 7, t2 = c2 & 7;
    const bool o0 = own_piece(c0, red), o1 = own_piece(c1, red), o2 = own_piece(c2, red);
    const bool p0 = o0 && t0 == PAWN, p1 = o1 && t1 == PAWN, p2 = o2 && t2 == PAWN;
    const bool s0 = o0 && (t0 == ROOK || t0 == CANNON), s1 = o1 && (t1 == ROOK || t1 == CANNON),
               s2 = o2 && (t2 == ROOK || t2 == CANNON);
    const uint32_t FULL = 0xffffffffu;
    const uint32_t np0 = __ballot_sync(FULL, o0 && !p0), np1 = __ballot_sync(FULL, o1 && !p1),
                   np2 = __ballot_sync(FULL, o2 && !p2);
    const uint32_t sl0 = __ballot_sync(FULL, s0), sl1 = __ballot_sync(FULL, s1), sl2 = __ballot_sync(FULL, s2);
    const uint32_t pw0 = __ballot_sync(FULL, p0), pw1 = __ballot_sync(FULL, p1), pw2 = __ballot_sync(FULL, p2);
    // is_insufficient_material: no pawn / cannon / rook / knight of either colour
    const bool a0 = c0 && t0 <= KNIGHT, a1 = c1 && t1 <= KNIGHT, a2 = c2 && t2 <= KNIGHT;
    const bool insufficient = __ballot_sync(FULL, a0 || a1 || a2) == 0u;
    const uint32_t kc = KING | (red ? 0 : BLACKBIT);
    const uint32_t k0 = __ballot_sync(FULL, c0 == kc), k1 = __ballot_sync(FULL, c1 == kc),
                   k2 = __ballot_sync(FULL, c2 == kc);
    const int ksq = k0 ? __ffs(k0) - 1 : k1 ? 31 + __ffs(k1) : k2 ? 63 + __ffs(k2) : -1;

    // per-file (lanes 0..8) and per-rank (lanes 16..25) occupancy words
    {
        const bool isF = lane < 9, isR = lane >= 16 && lane < 26;
        if (isF || isR) {
            const int a = isF ? lane : 9 * (lane - 16), st = isF ? 9 : 1, nn = isF ? 10 : 9;
            uint32_t v = 0;
#pragma unroll
            for (int i = 0; i < 10; ++i)
                if (i < nn) v |= (uint32_t)(B[a + st * i] != 0) << i;
            if (isF) w.occF[lane] = (uint16_t)v;
            else w.occR[lane - 16] = (uint16_t)v;
        }
    }
    // work items in generation order: non-pawns by from-square descending, then pawns descending;
    // a rook / cannon contributes its four rays in descending-destination order (up, right, left, down)
    const int n_np = __popc(np0) + __popc(np1) + __popc(np2);
    const int n_sl = __popc(sl0) + __popc(sl1) + __popc(sl2);
    const int n_pw = __popc(pw0) + __popc(pw1) + __popc(pw2);
    const uint32_t above = lane == 31 ? 0u : (FULL << (lane + 1));
#define CCZ_PUT_ITEMS(OWN, PAWNF, SLIDERF, SQ, NPA, SLA, PWA)                                        \
    if (OWN) {                                                                                       \
        if (PAWNF) {                                                                                 \
            const int b_ = n_np + 3 * n_sl + (PWA);                                                  \
            if (b_ < 32) w.items[b_] = (uint16_t)((SQ) | (4 << 8));                                  \
        } else {                                                                                     \
            const int b_ = (NPA) + 3 * (SLA);                                                        \
            if (SLIDERF) {                                                                           \
                _Pragma("unroll") for (int d_ = 0; d_ < 4; ++d_)                                     \
                    if (b_ + d_ < 32) w.items[b_ + d_] = (uint16_t)((SQ) | (d_ << 8));               \
            } else if (b_ < 32) {                                                                    \
                w.items[b_] = (uint16_t)((SQ) | (4 << 8));                                           \
            }                                                                                        \
        }                                                                                            \
    }
    CCZ_PUT_ITEMS(o0, p0, s0, lane, __popc(np0 & above) + __popc(np1) + __popc(np2),
                  __popc(sl0 & above) + __popc(sl1) + __popc(sl2), __popc(pw0 & above) + __popc(pw1) + __popc(pw2))
    CCZ_PUT_ITEMS(o1, p1, s1, lane + 32, __popc(np1 & above) + __popc(np2), __popc(sl1 & above) + __popc(sl2),
                  __popc(pw1 & above) + __popc(pw2))
    CCZ_PUT_ITEMS(o2, p2, s2, lane + 64, __popc(np2 & above), __popc(sl2 & above), __popc(pw2 & above))
#undef CCZ_PUT_ITEMS
    // knights that attack the king's square right now (bit i = jump i), for the legality test
    uint32_t km = 0;
    {
        bool hit = false;
        if (lane < 8 && ksq >= 0) {
            int dr, df, dleg;
            knight_geom(lane, dr, df, dleg);
            const int nr = ksq / 9 + dr, nf = ksq % 9 + df;
            hit = nr >= 0 && nr <= 9 && nf >= 0 && nf <= 8 && B[nr * 9 + nf] == (uint32_t)(KNIGHT | (red ? BLACKBIT : 0));
        }
        km = __ballot_sync(FULL, hit);
    }
    __syncwarp();
    const int n_items = min(n_np + 3 * n_sl + n_pw, 32);

    // ---- pseudo-legal generation: one work item per lane --------------------------------------
    int from = 0, cnt = 0;
    uint32_t ok = 0;                 // step item: accepted table slots
    int quiet = 0, capd = 0, step = 0; // slider item: quiet squares, capture distance, square step
    bool positive = false;
    const uint16_t *ent = s_step;
    int dir = 4;
    if (lane < n_items) {
        const uint32_t it = w.items[lane];
        from = it & 255;
        dir = it >> 8;
        const uint32_t pc = B[from];
        const int ty = pc & 7;
        if (dir == 4) {
            const int kind4 = (ty == PAWN ? 0 : ty - 4) * 2 + (red ? 0 : 1);
            const int nsl = ty == KNIGHT ? 8 : 4;
            ent = ty == KNIGHT ? s_step + from * 8 : s_step + STEP_KNIGHT_ENTRIES + (kind4 * 90 + from) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t e = ent[i];
                const int to = e & 127, blk = (e >> 7) & 127;
                const bool good = i < nsl && e != 0xFFFFu && (blk == 127 || B[blk] == 0) && !own_piece(B[to], red);
                ok |= (uint32_t)good << i;
            }
            cnt = __popc(ok);
        } else {
            const int r = from / 9, f = from - 9 * r;
            const bool vertical = dir == 0 || dir == 3;
            positive = dir < 2;
            step = dir == 0 ? 9 : dir == 1 ? 1 : dir == 2 ? -1 : -9;
            const int lim = dir == 0 ? 9 - r : dir == 1 ? 8 - f : dir == 2 ? f : r;
            const uint32_t x = ray_bits(vertical ? w.occF[f] : w.occR[r], vertical ? r : f, positive);
            const int d1 = __ffs(x);
            quiet = d1 ? d1 - 1 : lim;
            int dc = d1;                       // rook captures the first piece on the ray
            if (ty == CANNON) dc = __ffs(x & (x - 1u)); // cannon the second (exactly one screen)
            if (dc && enemy_piece(B[from + step * dc], red)) capd = dc;
            cnt = quiet + (capd ? 1 : 0);
        }
    }
    const int incl = warp_incl_scan(cnt, lane);
    const int M = min(__shfl_sync(FULL, incl, 31), MAX_MOVES);
    int j = incl - cnt;
    const uint32_t fhi = (uint32_t)from << 8;
    if (dir == 4) {
        while (ok) {
            const int i = __ffs(ok) - 1;
            ok &= ok - 1u;
            if (j < MAX_MOVES) w.mv[j] = (uint16_t)(fhi | (ent[i] & 127u));
            ++j;
        }
    } else {
        // destinations descending: positive rays far -> near (capture first), negative rays near -> far
        for (int i = 0; i < cnt; ++i, ++j) {
            int d;
            if (positive) d = capd ? (i == 0 ? capd : quiet + 1 - i) : quiet - i;
            else d = i < quiet ? i + 1 : capd;
            if (j < MAX_MOVES) w.mv[j] = (uint16_t)(fhi | (uint32_t)(from + step * d));
        }
    }
    __syncwarp();

    // ---- legality: one candidate per lane per round; index M is the null move (in-check test) ----
    int n_legal = 0;
    bool in_check = false;
    const uint32_t below = (1u << lane) - 1u;
    for (int base = 0; base <= M; base += 32) {
        const int k = base + lane;
        bool legal = false, chk = false;
        int id = -1;
        if (k < M) {
            const uint32_t mv = w.mv[k];
            const int f = mv >> 8, t = mv & 255;
            const uint32_t pc = B[f];
            const bool king_move = (pc & 7) == KING;
            legal = ksq < 0 || !attacked_after_move(w, B, f, t, pc, king_move ? t : ksq, king_move, km, red);
            id = s_id_of[f * 90 + t];
        } else if (k == M) {
            chk = ksq >= 0 && attacked_after_move(w, B, -1, -1, 0u, ksq, false, km, red);
        }
        const uint32_t bal = __ballot_sync(FULL, legal);
        if (legal) w.row[n_legal + __popc(bal & below)] = (int16_t)id;
        n_legal += __popc(bal);
        in_check |= __ballot_sync(FULL, chk) != 0u;
    }
    __syncwarp();
    int fl = 0;
    if (in_check) fl |= 1;
    if (n_legal == 0) fl |= 2;
    if (insufficient) fl |= 4;
    if (B[OFF_REP] >= 3) fl |= 8;
    if (B[OFF_CLOCK] >= 120 && n_legal > 0) fl |= 16;
    n_legal_out = n_legal;
    flags_out = fl;
}

// value of 32-bit word `off` (two bf16 elements) of a position's (17,7,10,9) input; generic
// path, only used for the ragged tail (n % 4 != 0)
__device__ __forceinline__ uint32_t plane_word(const uint8_t *B, int off) {
    // word ranges: [0,2205) zeros | [2205,2520) red current (play 7) | [2520,4725) zeros |
    //              [4725,5040) black current (play 15) | [5040,5355) turn plane (play 16)
    if (off >= 5040) return B[OFF_TURN] ? 0x3F803F80u : 0u;
    int d;
    uint32_t colour;
    if (off >= 2205 && off < 2520) { d = off - 2205; colour = 0u; }
    else if (off >= 4725) { d = off - 4725; colour = 8u; }
    else return 0u;
    const int p = d / 45;
    const int s = 2 * (d - 45 * p);
    const uint32_t code = (uint32_t)(p + 1) | colour;
    return (B[s] == code ? 0x00003F80u : 0u) | (B[s + 1] == code ? 0x3F800000u : 0u);
}

__device__ __forceinline__ void fill_chunks(uint4 *o4, int begin, int end, uint4 v, int lane) {
    for (int c = begin + lane; c < end; c += 32) o4[c] = v;
}

// Planes of a full quad.  In 16-byte chunks the 4 x 5355 words split at compile-time boundaries
// into zero runs (plays 0..15, the ones are dropped in afterwards), turn-plane runs (play 16 =
// all ones when RED is to move, net.py:170-173) and six chunks that straddle two runs.
__device__ __forceinline__ void encode_quad(const MgWarpSmem &w, uint32_t *out, int lane) {
    uint4 *o4 = reinterpret_cast<uint4 *>(out);
    const uint32_t ONE2 = 0x3F803F80u;
    const uint32_t u0 = w.boards[0][OFF_TURN] ? ONE2 : 0u, u1 = w.boards[1][OFF_TURN] ? ONE2 : 0u,
                   u2 = w.boards[2][OFF_TURN] ? ONE2 : 0u, u3 = w.boards[3][OFF_TURN] ? ONE2 : 0u;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    fill_chunks(o4, 0, 1260, z, lane);
    fill_chunks(o4, 1339, 2598, z, lane);
    fill_chunks(o4, 2678, 3937, z, lane);
    fill_chunks(o4, 4017, 5276, z, lane);
    fill_chunks(o4, 1260, 1338, make_uint4(u0, u0, u0, u0), lane);
    fill_chunks(o4, 2599, 2677, make_uint4(u1, u1, u1, u1), lane);
    fill_chunks(o4, 3938, 4016, make_uint4(u2, u2, u2, u2), lane);
    fill_chunks(o4, 5277, 5355, make_uint4(u3, u3, u3, u3), lane);
    if (lane < 6) {
        const int c = lane == 0 ? 1338 : lane == 1 ? 2598 : lane == 2 ? 2677 : lane == 3 ? 3937 : lane == 4 ? 4016 : 5276;
        const uint4 v = lane == 0 ? make_uint4(u0, u0, u0, 0u)
                      : lane == 1 ? make_uint4(0u, 0u, 0u, u1)
                      : lane == 2 ? make_uint4(u1, u1, 0u, 0u)
                      : lane == 3 ? make_uint4(0u, 0u, u2, u2)
                      : lane == 4 ? make_uint4(u2, 0u, 0u, 0u)
                                  : make_uint4(0u, u3, u3, u3);
        o4[c] = v;
    }
    __syncwarp(); // order the zero fill before the ones below (same addresses, different lanes)
    uint16_t *o16 = reinterpret_cast<uint16_t *>(out);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint8_t *B = w.boards[q];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int s = lane + 32 * r;
            if (s < 90) {
                const uint32_t c = B[s];
                // play 7 (red) starts at plane 49, play 15 (black) at plane 105; channel = type - 1
                if (c) o16[q * PLANE_ELEMS + (((c & 8u) ? 105 : 49) + (int)(c & 7u) - 1) * 90 + s] = 0x3F80;
            }
        }
    }
}

// ---- generation-order policy (ccz_order_policy in include/ccz_b200.h) ---------------------------------
// `board.legal_moves` order decides every MCTS tie-break (mcts.py:59-61) and cchess is not available to
// pin it, so the order is a POLICY: moves are sorted by (class_rank[piece type], from-square key, capture
// key, to-square key).  The default policy is the order movegen_one generates natively (non-pawns by
// from-square descending, destinations descending, then pawns) and costs nothing; any other policy runs
// the SORTED instantiation, which re-orders the legal ids of a position with a warp rank sort.
__constant__ uint8_t d_order_policy[12] = {0, 1, 0, 0, 0, 0, 0, 0, /*from_desc*/ 1, /*to_desc*/ 1, /*capture_mode*/ 0,
                                           /*check_king_first*/ 0};

__device__ __forceinline__ uint32_t order_key(const uint8_t *B, int id, bool in_check) {
    const int f = d_from_of[id], t = d_to_of[id];
    uint32_t cls = d_order_policy[B[f] & 7];
    if (d_order_policy[11] && in_check) cls = (B[f] & 7) == KING ? 0u : cls + 1u; // evasions: king moves first
    const uint32_t fk = d_order_policy[8] ? 89 - f : f, tk = d_order_policy[9] ? 89 - t : t;
    const uint32_t mode = d_order_policy[10];
    const uint32_t ck = mode == 0 ? 0u : (uint32_t)((B[t] != 0) != (mode == 2));
    return cls << 16 | fk << 9 | ck << 8 | tk;
}

// re-order w.row[0 .. n) by the policy key (keys are distinct: (from, to) is); whole warp
__device__ __forceinline__ void sort_row_by_policy(MgWarpSmem &w, const uint8_t *B, uint32_t *keys /*smem [128]*/, int n,
                                                   bool in_check, int lane) {
    uint32_t my_key[4];
    int my_id[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        my_id[r] = i < n ? w.row[i] : -1;
        my_key[r] = i < n ? order_key(B, my_id[r], in_check) : 0xffffffffu;
        if (i < n) keys[i] = my_key[r];
    }
    __syncwarp();
    int rank[4] = {0, 0, 0, 0};
    for (int j = 0; j < n; ++j) {
        const uint32_t kj = keys[j];
#pragma unroll
        for (int r = 0; r < 4; ++r) rank[r] += kj < my_key[r];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (lane + 32 * r < n) w.row[rank[r]] = (int16_t)my_id[r];
    __syncwarp();
}

template <bool SORTED>
__global__ void __launch_bounds__(MG_WARPS * 32, CCZ_MG_MIN_BLOCKS)
movegen_encode_kernel(const uint8_t *__restrict__ boards, int n, int16_t *__restrict__ move_ids,
                      int16_t *__restrict__ counts, uint8_t *__restrict__ flags, uint32_t *__restrict__ planes,
                      unsigned int *__restrict__ claim) {
    __shared__ uint32_t s_keys[SORTED ? MG_WARPS * MAX_MOVES : 1];
    __shared__ __align__(16) int16_t s_id_of[8100];
    __shared__ __align__(16) uint16_t s_step[STEP_TAB_ENTRIES];
    __shared__ MgWarpSmem s_w[MG_WARPS];
    for (int i = threadIdx.x; i < 8100 / 2; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_id_of)[i] = reinterpret_cast<const uint32_t *>(d_id_of)[i];
    static_assert(STEP_TAB_ENTRIES % 8 == 0, "uint4 copy");
    for (int i = threadIdx.x; i < STEP_TAB_ENTRIES / 8; i += blockDim.x)
        reinterpret_cast<uint4 *>(s_step)[i] = reinterpret_cast<const uint4 *>(d_step_tab)[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    MgWarpSmem &w = s_w[warp];
    const int n_quads = (n + 3) >> 2;
    // persistent warps claim quads from a global counter: one resident wave, no tail imbalance
    while (true) {
        int quad = 0;
        if (lane == 0) quad = (int)atomicAdd(claim, 1u);
        quad = __shfl_sync(0xffffffffu, quad, 0);
        if (quad >= n_quads) break;
        const int base = quad * 4;
        const int nb = min(4, n - base);
        if (lane < nb * 6)
            reinterpret_cast<uint4 *>(&w.boards[0][0])[lane] =
                __ldg(reinterpret_cast<const uint4 *>(boards + (size_t)base * BOARD_BYTES) + lane);
        __syncwarp();
CCZ_Q_UNROLL
        for (int q = 0; q < nb; ++q) {
            int n_legal, fl;
            movegen_one(w, w.boards[q], s_id_of, s_step, lane, n_legal, fl);
            if (SORTED)
                sort_row_by_policy(w, w.boards[q], s_keys + (SORTED ? warp * MAX_MOVES : 0), n_legal, (fl & CCZ_FLAG_CHECK) != 0, lane);
#ifdef CCZ_DEBUG_MV
            if (planes != nullptr) {
                uint16_t *dump = reinterpret_cast<uint16_t *>(planes) + (size_t)(base + q) * PLANE_ELEMS;
                for (int i = lane; i < MAX_MOVES; i += 32) dump[i] = w.mv[i];
                if (lane < 32) dump[128 + lane] = w.items[lane];
                if (lane < 10) { dump[160 + lane] = w.occR[lane]; dump[170 + lane] = w.occF[lane]; }
            }
#endif
            // one coalesced 256-byte row: 4 ids per lane, -1 padded
            const int i0 = lane * 4;
            const uint32_t v0 = i0 + 0 < n_legal ? (uint16_t)w.row[i0 + 0] : 0xFFFFu;
            const uint32_t v1 = i0 + 1 < n_legal ? (uint16_t)w.row[i0 + 1] : 0xFFFFu;
            const uint32_t v2 = i0 + 2 < n_legal ? (uint16_t)w.row[i0 + 2] : 0xFFFFu;
            const uint32_t v3 = i0 + 3 < n_legal ? (uint16_t)w.row[i0 + 3] : 0xFFFFu;
            reinterpret_cast<uint2 *>(move_ids + (size_t)(base + q) * MAX_MOVES)[lane] =
                make_uint2(v0 | (v1 << 16), v2 | (v3 << 16));
            if (lane == 0) { w.cnt[q] = (int16_t)n_legal; w.flg[q] = (uint8_t)fl; }
            __syncwarp();
        }
        if (lane < nb) {
            counts[base + lane] = w.cnt[lane];
            flags[base + lane] = w.flg[lane];
        }
#ifndef CCZ_DEBUG_MV
        if (planes != nullptr) {
            uint32_t *out = planes + (size_t)base * WORDS_PER_POS;
            if (nb == 4) {
                encode_quad(w, out, lane);
            } else { // ragged tail: word-granular generic path
                const int total_words = nb * WORDS_PER_POS;
                for (int i = lane; i < total_words; i += 32) {
                    const int q = i / WORDS_PER_POS;
                    out[i] = plane_word(w.boards[q], i - q * WORDS_PER_POS);
                }
            }
        }
#endif
        __syncwarp();
    }
}

// thread per board: fill with the start position
__global__ void boards_start_kernel(uint8_t *boards, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 6) return;
    reinterpret_cast<uint4 *>(boards)[i] = reinterpret_cast<const uint4 *>(d_start_board)[i % 6];
}

} // namespace ccz
