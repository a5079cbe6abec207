// ccz_movegen.cuh -- K1: fused legal-move generation + terminal flags + bf16 plane encoding.
//
// Replaces, per position:
//   [move_action2move_id[Move.uci(m)] for m in board.legal_moves]      net.py:154-157
//   decode_board + 7 zero states + current + turn plane                tools.py:74-106, net.py:160-177
//   is_game_over / is_tie predicates                                   tools.py:109-123, mcts.py:116
//
// Work decomposition: one warp owns a QUAD of 4 consecutive positions.  4 x 10710 bf16 =
// 85,680 B is the smallest run of whole positions that is 16-byte aligned, so every plane store
// of the quad is a full, aligned 16 B vector and the warp streams 512 B per store instruction.
// Move generation is warp-cooperative: lanes own pieces, then candidate moves; all board reads
// hit a 96-byte shared-memory copy.  HBM traffic per position = 96 B in, 256+2+1 B of move
// data and 21,420 B of planes out (DESIGN.md "K1").
#pragma once
#include "ccz_rules.cuh"

namespace ccz {

constexpr int MG_WARPS = 8;
constexpr int WORDS_PER_POS = PLANE_ELEMS / 2; // 5355 32-bit words (two bf16) per position

struct __align__(16) MgWarpSmem {
    uint8_t boards[4][BOARD_BYTES]; // 384
    uint16_t mv[MAX_MOVES];         // pseudo-legal candidates, from<<8|to, generation order
    int16_t row[MAX_MOVES];         // legal action ids, generation order
    uint8_t plist[16];              // own pieces in generation order
    int16_t cnt[4];
    uint8_t flg[4];
    uint8_t pad[4];
};

// One position, one warp.  Writes w.row / returns count and flag byte (uniform across lanes).
__device__ __forceinline__ void movegen_one(MgWarpSmem &w, const uint8_t *B, const int16_t *s_id_of, int lane,
                                            int &n_legal_out, int &flags_out) {
    const bool red = B[OFF_TURN] != 0;
    const uint32_t c0 = B[lane], c1 = B[lane + 32], c2 = lane < 26 ? B[lane + 64] : 0u;
    const bool o0 = own_piece(c0, red), o1 = own_piece(c1, red), o2 = own_piece(c2, red);
    const bool p0 = o0 && (c0 & 7) == PAWN, p1 = o1 && (c1 & 7) == PAWN, p2 = o2 && (c2 & 7) == PAWN;
    const uint32_t np0 = __ballot_sync(0xffffffffu, o0 && !p0), np1 = __ballot_sync(0xffffffffu, o1 && !p1),
                   np2 = __ballot_sync(0xffffffffu, o2 && !p2);
    const uint32_t pw0 = __ballot_sync(0xffffffffu, p0), pw1 = __ballot_sync(0xffffffffu, p1),
                   pw2 = __ballot_sync(0xffffffffu, p2);
    // is_insufficient_material: no pawn / cannon / rook / knight of either colour
    const bool a0 = c0 && (c0 & 7) <= KNIGHT, a1 = c1 && (c1 & 7) <= KNIGHT, a2 = c2 && (c2 & 7) <= KNIGHT;
    const bool insufficient = __ballot_sync(0xffffffffu, a0 || a1 || a2) == 0u;
    const uint32_t kc = KING | (red ? 0 : BLACKBIT);
    const uint32_t k0 = __ballot_sync(0xffffffffu, c0 == kc), k1 = __ballot_sync(0xffffffffu, c1 == kc),
                   k2 = __ballot_sync(0xffffffffu, c2 == kc);
    const int ksq = k0 ? __ffs(k0) - 1 : k1 ? 31 + __ffs(k1) : k2 ? 63 + __ffs(k2) : -1;

    // generation order: non-pawns by from-square descending, then pawns descending
    const int n_np = __popc(np0) + __popc(np1) + __popc(np2);
    const int n_pw = __popc(pw0) + __popc(pw1) + __popc(pw2);
    const uint32_t above = lane == 31 ? 0u : (0xffffffffu << (lane + 1));
    if (o0) {
        const int rk = p0 ? n_np + __popc(pw0 & above) + __popc(pw1) + __popc(pw2)
                          : __popc(np0 & above) + __popc(np1) + __popc(np2);
        if (rk < 16) w.plist[rk] = (uint8_t)lane;
    }
    if (o1) {
        const int rk = p1 ? n_np + __popc(pw1 & above) + __popc(pw2) : __popc(np1 & above) + __popc(np2);
        if (rk < 16) w.plist[rk] = (uint8_t)(lane + 32);
    }
    if (o2) {
        const int rk = p2 ? n_np + __popc(pw2 & above) : __popc(np2 & above);
        if (rk < 16) w.plist[rk] = (uint8_t)(lane + 64);
    }
    __syncwarp();
    const int n_pieces = min(n_np + n_pw, 16);

    Mask90 m{0u, 0u, 0u};
    int from = 0;
    if (lane < n_pieces) {
        from = w.plist[lane];
        m = gen_targets(B, from, B[from], red);
    }
    const int cnt = m.count();
    const int incl = warp_incl_scan(cnt, lane);
    const int M = min(__shfl_sync(0xffffffffu, incl, 31), MAX_MOVES);
    int j = incl - cnt;
    const uint32_t fhi = (uint32_t)from << 8;
    while (m.w2) { const int b = 31 - __clz(m.w2); m.w2 ^= 1u << b; if (j < MAX_MOVES) w.mv[j] = (uint16_t)(fhi | (64 + b)); ++j; }
    while (m.w1) { const int b = 31 - __clz(m.w1); m.w1 ^= 1u << b; if (j < MAX_MOVES) w.mv[j] = (uint16_t)(fhi | (32 + b)); ++j; }
    while (m.w0) { const int b = 31 - __clz(m.w0); m.w0 ^= 1u << b; if (j < MAX_MOVES) w.mv[j] = (uint16_t)(fhi | b); ++j; }
    __syncwarp();

    // legality: one candidate per lane per round; candidate index M is the null move (in-check test)
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
            const int kk = (pc & 7) == KING ? t : ksq;
            legal = kk < 0 || !king_attacked_after(B, f, t, pc, kk, !red);
            id = s_id_of[f * 90 + t];
        } else if (k == M) {
            chk = ksq >= 0 && king_attacked_after(B, -1, -1, 0u, ksq, !red);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, legal);
        if (legal) w.row[n_legal + __popc(bal & below)] = (int16_t)id;
        n_legal += __popc(bal);
        in_check |= __ballot_sync(0xffffffffu, chk) != 0u;
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

// value of 32-bit word `off` (two bf16 elements) of a position's (17,7,10,9) input
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

__global__ void __launch_bounds__(MG_WARPS * 32)
movegen_encode_kernel(const uint8_t *__restrict__ boards, int n, int16_t *__restrict__ move_ids,
                      int16_t *__restrict__ counts, uint8_t *__restrict__ flags, uint32_t *__restrict__ planes) {
    __shared__ __align__(16) int16_t s_id_of[8100];
    __shared__ MgWarpSmem s_w[MG_WARPS];
    for (int i = threadIdx.x; i < 8100 / 2; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_id_of)[i] = reinterpret_cast<const uint32_t *>(d_id_of)[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    MgWarpSmem &w = s_w[warp];
    const int n_quads = (n + 3) >> 2;
    for (int quad = blockIdx.x * MG_WARPS + warp; quad < n_quads; quad += gridDim.x * MG_WARPS) {
        const int base = quad * 4;
        const int nb = min(4, n - base);
        if (lane < nb * 6)
            reinterpret_cast<uint4 *>(&w.boards[0][0])[lane] =
                __ldg(reinterpret_cast<const uint4 *>(boards + (size_t)base * BOARD_BYTES) + lane);
        __syncwarp();
        for (int q = 0; q < nb; ++q) {
            int n_legal, fl;
            movegen_one(w, w.boards[q], s_id_of, lane, n_legal, fl);
            // one coalesced 256-byte row: 4 ids per lane, -1 padded
            const int i0 = lane * 4;
            uint32_t lo, hi;
            {
                const uint32_t v0 = i0 + 0 < n_legal ? (uint16_t)w.row[i0 + 0] : 0xFFFFu;
                const uint32_t v1 = i0 + 1 < n_legal ? (uint16_t)w.row[i0 + 1] : 0xFFFFu;
                const uint32_t v2 = i0 + 2 < n_legal ? (uint16_t)w.row[i0 + 2] : 0xFFFFu;
                const uint32_t v3 = i0 + 3 < n_legal ? (uint16_t)w.row[i0 + 3] : 0xFFFFu;
                lo = v0 | (v1 << 16);
                hi = v2 | (v3 << 16);
            }
            reinterpret_cast<uint2 *>(move_ids + (size_t)(base + q) * MAX_MOVES)[lane] = make_uint2(lo, hi);
            if (lane == 0) { w.cnt[q] = (int16_t)n_legal; w.flg[q] = (uint8_t)fl; }
            __syncwarp();
        }
        if (lane < nb) {
            counts[base + lane] = w.cnt[lane];
            flags[base + lane] = w.flg[lane];
        }
        if (planes != nullptr) {
            const int total_words = nb * WORDS_PER_POS;
            uint32_t *out = planes + (size_t)base * WORDS_PER_POS;
            const int n_chunks = (total_words + 3) >> 2;
            for (int c = lane; c < n_chunks; c += 32) {
                const int w0 = c * 4;
                const int q = w0 / WORDS_PER_POS;
                const int off = w0 - q * WORDS_PER_POS;
                const bool full = w0 + 4 <= total_words;
                if (full && (off + 3 < 2205 || (off >= 2520 && off + 3 < 4725))) {
                    reinterpret_cast<uint4 *>(out)[c] = make_uint4(0u, 0u, 0u, 0u);
                    continue;
                }
                uint32_t v[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    int qq = q, oo = off + t;
                    if (oo >= WORDS_PER_POS) { oo -= WORDS_PER_POS; ++qq; }
                    v[t] = qq < nb ? plane_word(w.boards[qq], oo) : 0u;
                }
                if (full) {
                    reinterpret_cast<uint4 *>(out)[c] = make_uint4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (w0 + t < total_words) out[w0 + t] = v[t];
                }
            }
        }
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
