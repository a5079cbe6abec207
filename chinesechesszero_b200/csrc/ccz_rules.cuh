// ccz_rules.cuh -- shared device-side definitions: the 96-byte board record, the constant tables
// and board.push.
//
// Board record (replaces cchess.Board): sq[90] (code = type | 8*black;
// PAWN=1 CANNON=2 ROOK=3 KNIGHT=4 BISHOP=5 ADVISOR=6 KING=7; square = file + 9*rank, rank 0 = Red's
// back rank, tools.py:91), [90] turn (1 = RED), [91] half-move clock (plies since the last
// capture), [92] repetition count (earlier occurrences of this position in the reversible
// window), [93..95] zero.
//   board.push  (mcts.py:111, game.py:201)  -> apply_move
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
__device__ __align__(16) int16_t d_id_of[8100]; // id_of[from*90+to], -1 = not an action (tools.py:232-269)
__device__ uint8_t d_from_of[N_ACTIONS];        // action id -> from square
__device__ uint8_t d_to_of[N_ACTIONS];          // action id -> to square
__device__ int16_t d_flip_of[N_ACTIONS];        // action id -> id of the file-mirrored move (tools.py:133-164)
__device__ uint64_t d_zkeys[16 * 90 + 1];       // position-key table [code][sq]; last = BLACK-to-move key
__device__ __align__(16) uint8_t d_start_board[BOARD_BYTES];

__device__ __forceinline__ bool own_piece(uint32_t c, bool red) { return c != 0u && ((c & 8u) == 0u) == red; }

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

// Apply action `mv` to the smem board (single thread): squares and turn.  Returns the updated
// position key; *captured tells whether the window of reversible positions restarts (the caller
// maintains the half-move clock and the key window).
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
