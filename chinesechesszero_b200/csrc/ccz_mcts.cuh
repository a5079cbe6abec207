// ccz_mcts.cuh -- K2..K7: flat-arena MCTS kernels, one warp per game, lockstep over all games.
//
// Replaces the reference's TreeNode object graph (mcts.py:7-178).  Exact-parity rules kept
// (SURVEY.md App. B): unvisited children score +inf and ties go to the first child in
// generation order (mcts.py:47-48,59-61); U = fp32(c_puct*P) * sqrt_fp64(N_parent) / (1+N) in
// fp64, Q is an fp32 incremental mean with every operation rounded separately (mcts.py:50-52,
// 69-71); priors are gathered, never renormalised (net.py:202-203); the leaf receives -v, its
// parent +v, ... (mcts.py:73-78,129); terminal leaves are 0.0 for draws, -1.0 for the
// mated / stalemated side to move (mcts.py:116-126).
#pragma once
#include "../../include/ccz_b200.h"
#include "ccz_rules.cuh"
#include <math_constants.h>

namespace ccz {

constexpr int MCTS_WARPS = 4;

struct __align__(16) SelWarpSmem {
    uint8_t board[BOARD_BYTES];
    uint64_t keys[KEY_WINDOW];
};

// number of earlier occurrences of `key` among keys[0 .. clock-1]
__device__ __forceinline__ int count_repeats(const uint64_t *keys, int clock, uint64_t key, int lane) {
    int c = 0;
#pragma unroll
    for (int t = 0; t < KEY_WINDOW / 32; ++t) {
        const int i = lane + 32 * t;
        c += __popc(__ballot_sync(0xffffffffu, i < clock && keys[i] == key));
    }
    return c;
}

// K3: Node.select / puct_value + path pushes (mcts.py:41-61,105-111)
__global__ void __launch_bounds__(MCTS_WARPS * 32)
mcts_select_kernel(ccz_arena a, float c_puct, uint8_t *leaf_boards, int32_t *leaf_nodes) {
    __shared__ SelWarpSmem s_w[MCTS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    SelWarpSmem &w = s_w[warp];
    const size_t nb = (size_t)g * a.node_cap;
    const int32_t *visits = a.d_visits + nb;
    const float *value = a.d_value + nb;
    const float *prior = a.d_prior + nb;

    if (lane < 6)
        reinterpret_cast<uint4 *>(w.board)[lane] =
            reinterpret_cast<const uint4 *>(a.d_root_boards + (size_t)g * BOARD_BYTES)[lane];
    __syncwarp();
    int clock = w.board[OFF_CLOCK];
    if (clock > KEY_WINDOW - 1) clock = KEY_WINDOW - 1;
    const uint64_t *rk = a.d_root_keys + (size_t)g * KEY_WINDOW;
#pragma unroll
    for (int t = 0; t < KEY_WINDOW / 32; ++t) {
        const int i = lane + 32 * t;
        if (i <= clock) w.keys[i] = rk[i];
    }
    __syncwarp();
    uint64_t key = w.keys[clock];

    int node = a.d_root[g];
    while (true) {
        const int nc = a.d_n_child[nb + node];
        if (nc <= 0) break;
        const int fc = a.d_first_child[nb + node];
        const double sq = sqrt((double)visits[node]); // np.sqrt(parent.visits): fp64, correctly rounded
        double best = -CUDART_INF;
        int best_i = 0x7fffffff;
        for (int i = lane; i < nc; i += 32) {
            const int n = visits[fc + i];
            double sc;
            if (n == 0) {
                sc = CUDART_INF;
            } else {
                const float cp = __fmul_rn(c_puct, prior[fc + i]); // np.float32(c_puct * prob)
                const double u = __ddiv_rn(__dmul_rn((double)cp, sq), (double)(1 + n));
                sc = __dadd_rn((double)value[fc + i], u);
            }
            if (sc > best) { best = sc; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
        }
        if (best_i == 0x7fffffff) best_i = 0; // all-NaN scores: max() keeps the first item
        node = fc + best_i;
        if (lane == 0) {
            const int mv = a.d_move[nb + node];
            bool captured;
            key = apply_move(w.board, mv, key, &captured);
            clock = captured ? 0 : min(clock + 1, KEY_WINDOW - 1);
            w.keys[clock] = key;
            w.board[OFF_CLOCK] = (uint8_t)clock;
        }
        __syncwarp();
        clock = w.board[OFF_CLOCK];
        key = w.keys[clock];
    }
    const int rep = count_repeats(w.keys, clock, key, lane);
    if (lane == 0) w.board[OFF_REP] = (uint8_t)min(rep, 255);
    __syncwarp();
    if (lane < 6)
        reinterpret_cast<uint4 *>(leaf_boards + (size_t)g * BOARD_BYTES)[lane] =
            reinterpret_cast<const uint4 *>(w.board)[lane];
    if (lane == 0) leaf_nodes[g] = node;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// K4+K5: Node.expand with softmax-gather priors, terminal values, update_recursive
__global__ void __launch_bounds__(MCTS_WARPS * 32)
mcts_expand_backup_kernel(ccz_arena a, const int32_t *leaf_nodes, const float *policy, int policy_kind,
                          const float *values, const int16_t *move_ids, const int16_t *counts,
                          const uint8_t *flags) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    const size_t nb = (size_t)g * a.node_cap;
    const int node = leaf_nodes[g];
    const int fl = flags[g];
    const int cnt = counts[g];
    float v;
    if ((fl & (CCZ_FLAG_TIE_MASK | CCZ_FLAG_NOMOVES)) == 0) {
        // mcts.py:117-119: not over, not a tie -> expand with (action, prob) in generation order
        const float *pol = policy + (size_t)g * N_ACTIONS;
        float mx = 0.f, inv = 1.f;
        if (policy_kind == CCZ_POLICY_LOGITS) {
            mx = -CUDART_INF_F;
            for (int i = lane; i < N_ACTIONS; i += 32) mx = fmaxf(mx, pol[i]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int i = lane; i < N_ACTIONS; i += 32) sum += expf(pol[i] - mx);
            inv = 1.f / warp_sum(sum);
        }
        const int first = a.d_n_nodes[g];
        if (first + cnt > a.node_cap) {
            if (lane == 0) a.d_status[g] |= CCZ_STATUS_NODE_OVERFLOW;
        } else {
            for (int i = lane; i < cnt; i += 32) {
                const int id = move_ids[(size_t)g * MAX_MOVES + i];
                // ids are always valid for reachable positions; a piece on an impossible square
                // (hand-made records) has no action id and gets prior 0 instead of an OOB read
                float p = (id >= 0 && id < N_ACTIONS) ? pol[id] : 0.f;
                if (policy_kind == CCZ_POLICY_LOGITS && id >= 0) p = expf(p - mx) * inv;
                const size_t c = nb + first + i;
                a.d_visits[c] = 0;
                a.d_value[c] = 0.f;
                a.d_prior[c] = p;
                a.d_move[c] = (int16_t)id;
                a.d_first_child[c] = -1;
                a.d_n_child[c] = 0;
                a.d_parent[c] = node;
            }
            if (lane == 0) {
                a.d_first_child[nb + node] = first;
                a.d_n_child[nb + node] = (int16_t)cnt;
                a.d_n_nodes[g] = first + cnt;
            }
        }
        v = values[g];
    } else if (fl & CCZ_FLAG_TIE_MASK) {
        v = 0.0f; // mcts.py:120-122
    } else {
        v = -1.0f; // mcts.py:123-126: winner is never the side to move
    }
    // mcts.py:129,63-78: leaf gets -v, parent +v, ...;  Q += 1.0*(x - Q)/N in fp32
    if (lane == 0) {
        float x = -v;
        int cur = node;
        while (cur >= 0) {
            const size_t c = nb + cur;
            const int n = a.d_visits[c] + 1;
            a.d_visits[c] = n;
            const float q = a.d_value[c];
            a.d_value[c] = __fadd_rn(q, __fdiv_rn(__fsub_rn(x, q), (float)n));
            x = -x;
            cur = a.d_parent[c];
        }
    }
}

// root children read-out (mcts.py:163-164)
__global__ void __launch_bounds__(MCTS_WARPS * 32)
mcts_root_visits_kernel(ccz_arena a, int16_t *acts, int32_t *visits, int16_t *counts) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    const size_t nb = (size_t)g * a.node_cap;
    const int root = a.d_root[g];
    const int nc = a.d_n_child[nb + root];
    const int fc = a.d_first_child[nb + root];
    for (int i = lane; i < MAX_MOVES; i += 32) {
        const bool ok = i < nc;
        acts[(size_t)g * MAX_MOVES + i] = ok ? a.d_move[nb + fc + i] : (int16_t)-1;
        visits[(size_t)g * MAX_MOVES + i] = ok ? a.d_visits[nb + fc + i] : 0;
    }
    if (lane == 0) counts[g] = (int16_t)nc;
}

__device__ __forceinline__ void fresh_root(const ccz_arena &d, size_t nb, int g) {
    d.d_visits[nb] = 0;
    d.d_value[nb] = 0.f;
    d.d_prior[nb] = 1.0f; // Node(None, 1.0), mcts.py:94,178
    d.d_move[nb] = -1;
    d.d_first_child[nb] = -1;
    d.d_n_child[nb] = 0;
    d.d_parent[nb] = -1;
    d.d_root[g] = 0;
    d.d_n_nodes[g] = 1;
}

// push `mv` on a board record + key window held by one warp (window in global memory)
__device__ __forceinline__ void push_with_keys(uint8_t *B /*smem*/, uint64_t *keys /*global*/, int mv, int lane) {
    int clock = min((int)B[OFF_CLOCK], KEY_WINDOW - 1);
    uint64_t key = 0;
    if (lane == 0) {
        bool captured;
        key = apply_move(B, mv, keys[clock], &captured);
        clock = captured ? 0 : min(clock + 1, KEY_WINDOW - 1);
        keys[clock] = key;
        B[OFF_CLOCK] = (uint8_t)clock;
    }
    __syncwarp();
    clock = B[OFF_CLOCK];
    key = __shfl_sync(0xffffffffu, key, 0);
    const int rep = count_repeats(keys, clock, key, lane);
    if (lane == 0) B[OFF_REP] = (uint8_t)min(rep, 255);
    __syncwarp();
}

__device__ __forceinline__ uint64_t board_key(const uint8_t *B, int lane) {
    uint64_t k = 0;
    for (int s = lane; s < 90; s += 32)
        if (B[s]) k ^= zkey(B[s], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k ^= __shfl_xor_sync(0xffffffffu, k, o);
    if (B[OFF_TURN] == 0) k ^= zkey_turn();
    return k;
}

// K7: MCTS.update_with_move (mcts.py:168-178): keep the chosen child's sub-tree (breadth-first
// compaction src -> dst, children stay contiguous and ordered), advance the root position.
__global__ void __launch_bounds__(MCTS_WARPS * 32)
mcts_advance_kernel(ccz_arena src, ccz_arena dst, const int16_t *chosen) {
    __shared__ __align__(16) uint8_t s_board[MCTS_WARPS][BOARD_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= src.n_games) return;
    const size_t nb = (size_t)g * src.node_cap;
    const int mv = chosen[g];
    uint8_t *B = s_board[warp];
    uint64_t *dkeys = dst.d_root_keys + (size_t)g * KEY_WINDOW;
    const uint64_t *skeys = src.d_root_keys + (size_t)g * KEY_WINDOW;
    if (lane == 0) dst.d_status[g] = src.d_status[g];

    if (mv == -1) { // new game
        if (lane < 6) {
            const uint4 v = reinterpret_cast<const uint4 *>(d_start_board)[lane];
            reinterpret_cast<uint4 *>(dst.d_root_boards + (size_t)g * BOARD_BYTES)[lane] = v;
            reinterpret_cast<uint4 *>(B)[lane] = v;
        }
        __syncwarp();
        const uint64_t k = board_key(B, lane);
        if (lane == 0) { dkeys[0] = k; fresh_root(dst, nb, g); }
        return;
    }
    // carry the position and its key window over
    if (lane < 6)
        reinterpret_cast<uint4 *>(B)[lane] =
            reinterpret_cast<const uint4 *>(src.d_root_boards + (size_t)g * BOARD_BYTES)[lane];
    for (int i = lane; i < KEY_WINDOW; i += 32) dkeys[i] = skeys[i];
    __syncwarp();
    if (mv >= 0) push_with_keys(B, dkeys, mv, lane);
    if (lane < 6)
        reinterpret_cast<uint4 *>(dst.d_root_boards + (size_t)g * BOARD_BYTES)[lane] =
            reinterpret_cast<const uint4 *>(B)[lane];

    // locate the chosen child
    const int root = src.d_root[g];
    const int rnc = src.d_n_child[nb + root];
    const int rfc = src.d_first_child[nb + root];
    int child = -1;
    if (mv >= 0) {
        for (int base = 0; base < rnc && child < 0; base += 32) {
            const int i = base + lane;
            const uint32_t hit = __ballot_sync(0xffffffffu, i < rnc && src.d_move[nb + rfc + i] == mv);
            if (hit) child = rfc + base + __ffs(hit) - 1;
        }
    }
    if (child < 0) { // mcts.py:177-178: unknown move (or reset) -> fresh root
        if (lane == 0) fresh_root(dst, nb, g);
        return;
    }
    if (lane == 0) {
        dst.d_visits[nb] = src.d_visits[nb + child];
        dst.d_value[nb] = src.d_value[nb + child];
        dst.d_prior[nb] = src.d_prior[nb + child];
        dst.d_move[nb] = src.d_move[nb + child];
        dst.d_first_child[nb] = src.d_first_child[nb + child]; // old index, remapped below
        dst.d_n_child[nb] = src.d_n_child[nb + child];
        dst.d_parent[nb] = -1;
        dst.d_root[g] = 0;
    }
    int head = 0, tail = 1;
    while (head < tail) {
        __syncwarp();
        const int i = head + lane;
        const bool live = i < tail;
        const int nc = live ? (int)dst.d_n_child[nb + i] : 0;
        const int ofc = live ? dst.d_first_child[nb + i] : -1;
        const int incl = warp_incl_scan(nc, lane);
        const int nfc = tail + incl - nc;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (live && nc > 0) dst.d_first_child[nb + i] = nfc;
        uint32_t todo = __ballot_sync(0xffffffffu, nc > 0);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ncl = __shfl_sync(0xffffffffu, nc, l);
            const int ofl = __shfl_sync(0xffffffffu, ofc, l);
            const int nfl = __shfl_sync(0xffffffffu, nfc, l);
            for (int j = lane; j < ncl; j += 32) {
                const size_t s = nb + ofl + j, d = nb + nfl + j;
                dst.d_visits[d] = src.d_visits[s];
                dst.d_value[d] = src.d_value[s];
                dst.d_prior[d] = src.d_prior[s];
                dst.d_move[d] = src.d_move[s];
                dst.d_first_child[d] = src.d_first_child[s];
                dst.d_n_child[d] = src.d_n_child[s];
                dst.d_parent[d] = head + l;
            }
        }
        head = min(head + 32, tail);
        tail += total;
    }
    if (lane == 0) dst.d_n_nodes[g] = tail;
}

// Node(None, 1.0) over the start position for every game (mcts.py:94; game.py:148)
__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_reset_kernel(ccz_arena a, const uint8_t *mask) {
    __shared__ __align__(16) uint8_t s_board[MCTS_WARPS][BOARD_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    if (mask != nullptr && mask[g] == 0) return;
    uint8_t *B = s_board[warp];
    if (lane < 6) {
        const uint4 v = reinterpret_cast<const uint4 *>(d_start_board)[lane];
        reinterpret_cast<uint4 *>(a.d_root_boards + (size_t)g * BOARD_BYTES)[lane] = v;
        reinterpret_cast<uint4 *>(B)[lane] = v;
    }
    __syncwarp();
    const uint64_t k = board_key(B, lane);
    if (lane == 0) {
        a.d_root_keys[(size_t)g * KEY_WINDOW] = k;
        a.d_status[g] = 0;
        fresh_root(a, (size_t)g * a.node_cap, g);
    }
}

// K2: board.push on a batch of records (game.py:201), optional key windows for exact repetition
__global__ void __launch_bounds__(MCTS_WARPS * 32)
board_push_kernel(uint8_t *boards, const int16_t *move_ids, int n, uint64_t *keys) {
    __shared__ __align__(16) uint8_t s_board[MCTS_WARPS][BOARD_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= n) return;
    const int mv = move_ids[g];
    if (mv < 0 || mv >= N_ACTIONS) return;
    uint8_t *B = s_board[warp];
    if (lane < 6)
        reinterpret_cast<uint4 *>(B)[lane] = reinterpret_cast<const uint4 *>(boards + (size_t)g * BOARD_BYTES)[lane];
    __syncwarp();
    if (keys != nullptr) {
        push_with_keys(B, keys + (size_t)g * KEY_WINDOW, mv, lane);
    } else if (lane == 0) {
        bool captured;
        apply_move(B, mv, 0ull, &captured);
        const int clock = B[OFF_CLOCK];
        B[OFF_CLOCK] = (uint8_t)(captured ? 0 : min(clock + 1, 255));
        B[OFF_REP] = 0;
    }
    __syncwarp();
    if (lane < 6)
        reinterpret_cast<uint4 *>(boards + (size_t)g * BOARD_BYTES)[lane] = reinterpret_cast<const uint4 *>(B)[lane];
}

// key window initialisation for history-less boards: keys[clock] = key(board)
__global__ void __launch_bounds__(MCTS_WARPS * 32) board_keys_init_kernel(const uint8_t *boards, int n, uint64_t *keys) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= n) return;
    const uint8_t *B = boards + (size_t)g * BOARD_BYTES;
    const uint64_t k = board_key(B, lane);
    if (lane == 0) keys[(size_t)g * KEY_WINDOW + min((int)B[OFF_CLOCK], KEY_WINDOW - 1)] = k;
}

} // namespace ccz
