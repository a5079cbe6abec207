// ccz_rules.cuh -- device-side Xiangqi rules on a 96-byte board record held in shared memory.
//
// Replaces the arithmetic the reference gets from cchess:
//   board.legal_moves  (net.py:154-157)   -> gen_targets + king_attacked_after
//   board.push         (mcts.py:111)      -> apply_move
// Board record: sq[90] (code = type | 8*black; PAWN=1 CANNON=2 ROOK=3 KNIGHT=4 BISHOP=5
// ADVISOR=6 KING=7; square = file + 9*rank, rank 0 = Red's back rank, tools.py:91),
// [90] turn (1 = RED), [91] half-move clock, [92] repetition count, [93..95] zero.
#pragma once
#include <stdint.h>

namespace ccz {

constexpr int PAWN = 1, CANNON = 2, ROOK = 3, KNIGHT = 4, BISHOP = 5, ADVISOR = 6, KING = 7;
constexpr int BLACKBIT = 8;
constexpr int BOARD_BYTES = 96;
constexpr int OFF_TURN = 90, OFF_CLOCK = 91, OFF_REP = 92;
constexpr int MAX_MOVES = 128;
constexpr int N_ACTIONS = 2086;
constexpr int PLANE_ELEMS = 10710;
constexpr int KEY_WINDOW = 128;

// constant tables, uploaded once per device by ccz_init()
__device__ __align__(16) int16_t d_id_of[8100];      // id_of[from*90+to], -1 = not an action (tools.py:232-269)
__device__ uint8_t d_from_of[N_ACTIONS]; // action id -> from square
__device__ uint8_t d_to_of[N_ACTIONS];   // action id -> to square
__device__ int16_t d_flip_of[N_ACTIONS]; // action id -> id of the file-mirrored move (tools.py:133-164)
__device__ uint64_t d_zkeys[16 * 90 + 1]; // position-key table [code][sq]; last = BLACK-to-move key
__device__ __align__(16) uint8_t d_start_board[BOARD_BYTES];

__device__ __forceinline__ bool own_piece(uint32_t c, bool red) { return c != 0u && ((c & 8u) == 0u) == red; }

struct Mask90 {
    uint32_t w0, w1, w2;
    __device__ __forceinline__ void set(int s) {
        const uint32_t b = 1u << (s & 31);
        const int k = s >> 5;
        w0 |= (k == 0) ? b : 0u;
        w1 |= (k == 1) ? b : 0u;
        w2 |= (k == 2) ? b : 0u;
    }
    __device__ __forceinline__ int count() const { return __popc(w0) + __popc(w1) + __popc(w2); }
};

// Pseudo-legal destinations of the piece `pc` standing on `from` (one thread, board in smem).
__device__ __forceinline__ Mask90 gen_targets(const uint8_t *B, int from, uint32_t pc, bool red) {
    Mask90 m{0u, 0u, 0u};
    const int r = from / 9, f = from - 9 * r;
    const int ty = pc & 7;
    if (ty == ROOK || ty == CANNON) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int step = d == 0 ? 9 : d == 1 ? -9 : d == 2 ? 1 : -1;
            const int lim = d == 0 ? 9 - r : d == 1 ? r : d == 2 ? 8 - f : f;
            int s = from;
            bool screen = false;
            for (int i = 0; i < lim; ++i) {
                s += step;
                const uint32_t c = B[s];
                if (!screen) {
                    if (c == 0u) { m.set(s); continue; }
                    if (ty == ROOK) { if (!own_piece(c, red)) m.set(s); break; }
                    screen = true;
                } else if (c != 0u) {
                    if (!own_piece(c, red)) m.set(s);
                    break;
                }
            }
        }
    } else if (ty == KNIGHT) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int dr = (i < 4) ? ((i & 2) ? -2 : 2) : ((i & 2) ? -1 : 1);
            const int df = (i < 4) ? ((i & 1) ? -1 : 1) : ((i & 1) ? -2 : 2);
            const int rr = r + dr, ff = f + df;
            if (rr < 0 || rr > 9 || ff < 0 || ff > 8) continue;
            const int leg = (i < 4) ? from + (dr > 0 ? 9 : -9) : from + (df > 0 ? 1 : -1);
            if (B[leg] != 0) continue;
            const int t = rr * 9 + ff;
            if (!own_piece(B[t], red)) m.set(t);
        }
    } else if (ty == BISHOP) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int dr = (i & 2) ? -2 : 2, df = (i & 1) ? -2 : 2;
            const int rr = r + dr, ff = f + df;
            if (rr < 0 || rr > 9 || ff < 0 || ff > 8) continue;
            if (red ? rr > 4 : rr < 5) continue;
            if (B[from + (dr / 2) * 9 + df / 2] != 0) continue;
            const int t = rr * 9 + ff;
            if (!own_piece(B[t], red)) m.set(t);
        }
    } else if (ty == ADVISOR) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = r + ((i & 2) ? -1 : 1), ff = f + ((i & 1) ? -1 : 1);
            if (ff < 3 || ff > 5) continue;
            if (red ? (rr < 0 || rr > 2) : (rr < 7 || rr > 9)) continue;
            const int t = rr * 9 + ff;
            if (!own_piece(B[t], red)) m.set(t);
        }
    } else if (ty == KING) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = r + (i == 0 ? 1 : i == 1 ? -1 : 0), ff = f + (i == 2 ? 1 : i == 3 ? -1 : 0);
            if (ff < 3 || ff > 5) continue;
            if (red ? (rr < 0 || rr > 2) : (rr < 7 || rr > 9)) continue;
            const int t = rr * 9 + ff;
            if (!own_piece(B[t], red)) m.set(t);
        }
    } else if (ty == PAWN) {
        const int fr = r + (red ? 1 : -1);
        if (fr >= 0 && fr <= 9 && !own_piece(B[fr * 9 + f], red)) m.set(fr * 9 + f);
        if (red ? r >= 5 : r <= 4) {
            if (f > 0 && !own_piece(B[from - 1], red)) m.set(from - 1);
            if (f < 8 && !own_piece(B[from + 1], red)) m.set(from + 1);
        }
    }
    return m;
}

// After moving `pc` from `from` to `to` (from == to == -1: no move), is square `k` attacked by
// colour `by_red`?  Covers rook, cannon-over-one-screen, hobbled knight, pawn and the enemy
// king down an open file (flying general).
__device__ __forceinline__ bool king_attacked_after(const uint8_t *B, int from, int to, uint32_t pc, int k,
                                                    bool by_red) {
    const int r = k / 9, f = k - 9 * r;
#define CCZ_GET(s) ((s) == from ? 0u : ((s) == to ? pc : (uint32_t)B[(s)]))
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int step = d == 0 ? 9 : d == 1 ? -9 : d == 2 ? 1 : -1;
        const int lim = d == 0 ? 9 - r : d == 1 ? r : d == 2 ? 8 - f : f;
        int s = k;
        bool seen = false;
        for (int i = 1; i <= lim; ++i) {
            s += step;
            const uint32_t c = CCZ_GET(s);
            if (c == 0u) continue;
            const bool theirs = own_piece(c, by_red);
            const int ty = c & 7;
            if (!seen) {
                if (theirs) {
                    if (ty == ROOK) return true;
                    if (ty == KING && d < 2) return true;
                    if (ty == PAWN && i == 1) {
                        if (d < 2) {
                            // pawn on s attacks its forward square: s + fwd == k  <=>  step == -fwd
                            if ((by_red ? 9 : -9) == -step) return true;
                        } else {
                            const int pr = s / 9;
                            if (by_red ? pr >= 5 : pr <= 4) return true;
                        }
                    }
                }
                seen = true;
            } else {
                if (theirs && ty == CANNON) return true;
                break;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int dr = (i < 4) ? ((i & 2) ? -2 : 2) : ((i & 2) ? -1 : 1);
        const int df = (i < 4) ? ((i & 1) ? -1 : 1) : ((i & 1) ? -2 : 2);
        const int nr = r + dr, nf = f + df;
        if (nr < 0 || nr > 9 || nf < 0 || nf > 8) continue;
        const int n = nr * 9 + nf;
        const uint32_t c = CCZ_GET(n);
        if ((c & 7) != KNIGHT || !own_piece(c, by_red)) continue;
        // the knight travels (-dr,-df); its leg sits next to it along the long axis
        const int leg = (i < 4) ? n - (dr > 0 ? 9 : -9) : n - (df > 0 ? 1 : -1);
        if (CCZ_GET(leg) == 0u) return true;
    }
#undef CCZ_GET
    return false;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ uint64_t zkey(uint32_t code, int sq) { return d_zkeys[code * 90 + sq]; }
__device__ __forceinline__ uint64_t zkey_turn() { return d_zkeys[16 * 90]; }

// Apply action `mv` to the smem board (single thread): squares, turn, clock.  Returns the
// updated position key; *captured tells whether the window of reversible positions restarts.
__device__ __forceinline__ uint64_t apply_move(uint8_t *B, int mv, uint64_t key, bool *captured) {
    const int f = d_from_of[mv], t = d_to_of[mv];
    const uint32_t pc = B[f], cap = B[t];
    B[t] = (uint8_t)pc;
    B[f] = 0;
    B[OFF_TURN] ^= 1;
    key ^= zkey(pc, f) ^ zkey(pc, t) ^ zkey_turn();
    if (cap) key ^= zkey(cap, t);
    *captured = cap != 0u;
    return key;
}

} // namespace ccz
