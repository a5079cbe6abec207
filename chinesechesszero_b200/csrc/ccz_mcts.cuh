// ccz_mcts.cuh -- K2..K7: pooled-arena MCTS kernels, one warp per game, lockstep over all games.
//
// Replaces the reference's TreeNode object graph (mcts.py:7-178).  Exact-parity rules kept
// (SURVEY.md App. B): unvisited children score +inf and ties go to the first child in
// generation order (mcts.py:47-48,59-61); U = fp32(c_puct*P) * sqrt_fp64(N_parent) / (1+N) in
// fp64, Q is an fp32 incremental mean with every operation rounded separately (mcts.py:50-52,
// 69-71); priors are gathered, never renormalised (net.py:202-203); the leaf receives -v, its
// parent +v, ... (mcts.py:73-78,129); terminal leaves are 0.0 for draws, -1.0 for the
// mated / stalemated side to move (mcts.py:116-126).
//
// The reference's tree has no capacity (Node.expand allocates without bound, mcts.py:31-39, and
// update_with_move keeps arbitrarily large sub-trees, mcts.py:168-178), so the nodes of all games
// come from ONE pool of pages (ccz_arena in include/ccz_b200.h): a game bump-allocates child runs in
// its current page and pops another page from a global free ring when a run does not fit.  Kernels
// either pop (expand, advance-compact, migrate) or push (advance-release, reset, reserve) pages,
// never both, so the ring needs no ordering between a slot write and its read.
#pragma once
#include "../../include/ccz_b200.h"
#include "ccz_rules.cuh"
#include <math_constants.h>

namespace ccz {

constexpr int MCTS_WARPS = 4;
constexpr uint32_t FULL = 0xffffffffu;

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
        c += __popc(__ballot_sync(FULL, i < clock && keys[i] == key));
    }
    return c;
}

// ---- node records as vectors: ccz_node = int4 {N, Q bits, P bits, first_child}, ccz_link = int2 {parent, move | n_child << 16}
__device__ __forceinline__ int4 *node_vec(const ccz_arena &a) { return reinterpret_cast<int4 *>(a.d_nodes); }
__device__ __forceinline__ int2 *link_vec(const ccz_arena &a) { return reinterpret_cast<int2 *>(a.d_links); }
__device__ __forceinline__ int link_word(int move, int n_child) { return (move & 0xffff) | (n_child << 16); }
__device__ __forceinline__ int link_move(int w) { return (int)(short)(w & 0xffff); }
__device__ __forceinline__ int link_n_child(int w) { return w >> 16; }

// Node(None, 1.0) (mcts.py:94,178) at pool index idx
__device__ __forceinline__ void fresh_root(const ccz_arena &a, int idx) {
    node_vec(a)[idx] = make_int4(0, __float_as_int(0.f), __float_as_int(1.0f), -1);
    link_vec(a)[idx] = make_int2(-1, link_word(-1, 0));
}

// ---- page pool ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long *pool_ctl(const ccz_arena &a) {
    return reinterpret_cast<unsigned long long *>(a.d_pool_ctl);
}

// one thread: take a page off the free ring, -1 when the ring is empty.  Pop-only kernels: TAIL does not
// move while they run, so a ticket >= TAIL means "empty" for good and is handed back; a plain fetch-add
// keeps thousands of simultaneous pops (every game crosses its page boundaries at about the same playout,
// and every advance starts with one) at one L2 atomic each -- a compare-and-swap loop on this one address
// cost 14 ms per advance at 4096 games (profiles/r02_mcts_advance_cas_ncu.csv).
__device__ __forceinline__ int pool_pop(const ccz_arena &a) {
    unsigned long long *ctl = pool_ctl(a);
    const unsigned long long tail = *reinterpret_cast<volatile unsigned long long *>(ctl + CCZ_CTL_TAIL);
    const unsigned long long h = atomicAdd(ctl + CCZ_CTL_HEAD, 1ull);
    if (h >= tail) {
        atomicAdd(ctl + CCZ_CTL_HEAD, ~0ull); // -1: give the ticket back
        return -1;
    }
    atomicMin(reinterpret_cast<long long *>(ctl + CCZ_CTL_MIN_FREE), (long long)(tail - h - 1));
    return a.d_free_ring[h % (unsigned long long)a.n_pages];
}

// whole warp: give back the pages list[from .. n) except `keep` (push-only kernels)
__device__ __forceinline__ void pool_push_list(const ccz_arena &a, const int32_t *list, int from, int n, int keep, int lane) {
    unsigned long long *ctl = pool_ctl(a);
    for (int base = from; base < n; base += 32) {
        const int i = base + lane;
        const int page = i < n ? list[i] : -1;
        const bool has = page >= 0 && page != keep;
        const uint32_t m = __ballot_sync(FULL, has);
        if (m == 0u) continue;
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(ctl + CCZ_CTL_TAIL, (unsigned long long)__popc(m));
        b = __shfl_sync(FULL, b, 0);
        if (has) a.d_free_ring[(b + __popc(m & ((1u << lane) - 1u))) % (unsigned long long)a.n_pages] = page;
    }
}

// whole warp, uniform arguments: a run of cnt (<= 128) contiguous slots inside one page of the game
// whose allocation state is (page, off, np, list); -1 = pool exhausted or page list full
__device__ __forceinline__ int alloc_run(const ccz_arena &a, int cnt, int &page, int &off, int &np,
                                         volatile int32_t *list, volatile int32_t *fill, int lane) {
    if (off + cnt > (1 << a.page_shift)) {
        int p = -1;
        if (np < a.max_pages) {
            if (lane == 0) p = pool_pop(a);
            p = __shfl_sync(FULL, p, 0);
        }
        if (p < 0) return -1;
        if (lane == 0) {
            if (fill != nullptr) fill[np - 1] = off;
            list[np] = p;
        }
        ++np;
        page = p;
        off = 0;
    }
    const int idx = (page << a.page_shift) + off;
    off += cnt;
    return idx;
}

__device__ __forceinline__ int32_t *page_list(const ccz_arena &a, int half, int g) {
    return a.d_page_list + ((size_t)half * a.n_games + g) * a.max_pages;
}

// ---- TMA bulk copy + mbarrier (the STAGED variant of K3) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// K3: Node.select / puct_value + path pushes (mcts.py:41-61,105-111).  Per level the child run -- one
// contiguous array of 16-byte records (N, Q, P, first_child), <= 128 of them, inside one pool page --
// reaches the lanes either
//   STAGED = false: as one coalesced 16-byte load per lane straight into registers, or
//   STAGED = true : through shared memory, staged by ONE TMA bulk copy (cp.async.bulk + mbarrier) issued by
//                   lane 0, after which the lanes read their records from the staging buffer;
// the scores are reduced by warp shuffles (lowest index wins ties), the winner's link word (move, n_child)
// is the only other dependent load, and the move is replayed on a shared-memory board with its
// incremental key.  Both variants are bit-identical; DESIGN.md has the measured difference.
template <bool STAGED>
__global__ void __launch_bounds__(MCTS_WARPS * 32)
mcts_select_kernel(ccz_arena a, float c_puct, uint8_t *leaf_boards, int32_t *leaf_nodes) {
    __shared__ SelWarpSmem s_w[MCTS_WARPS];
    __shared__ __align__(16) int4 s_kids[STAGED ? MCTS_WARPS : 1][STAGED ? MAX_MOVES : 1];
    __shared__ __align__(8) uint64_t s_bar[MCTS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    const uint32_t bar = smem_u32(&s_bar[warp]);
    if (STAGED) {
        if (lane == 0) mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    if (g >= a.n_games) return;
    SelWarpSmem &w = s_w[warp];
    const int4 *nodes = node_vec(a);
    const int2 *links = link_vec(a);

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
    int4 cur = nodes[node];
    int nc = link_n_child(links[node].y);
    uint32_t phase = 0;
    while (nc > 0) {
        const int fc = cur.w;
        const int4 *kids = nodes + fc;
        if (STAGED) {
            if (lane == 0) {
                mbar_expect_tx(bar, (uint32_t)nc * 16u);
                bulk_g2s(smem_u32(&s_kids[warp][0]), kids, (uint32_t)nc * 16u, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            kids = &s_kids[warp][0];
        }
        const double sq = sqrt((double)cur.x); // np.sqrt(parent.visits): fp64, correctly rounded
        double best = -CUDART_INF;
        int best_i = 0x7fffffff;
        for (int i = lane; i < nc; i += 32) {
            const int4 c = kids[i];
            double sc;
            if (c.x == 0) {
                sc = CUDART_INF;
            } else {
                const float cp = __fmul_rn(c_puct, __int_as_float(c.z)); // np.float32(c_puct * prob)
                const double u = __ddiv_rn(__dmul_rn((double)cp, sq), (double)(1 + c.x));
                sc = __dadd_rn((double)__int_as_float(c.y), u);
            }
            if (sc > best) { best = sc; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, best_i, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
        }
        if (best_i == 0x7fffffff) best_i = 0; // all-NaN scores: max() keeps the first item
        node = fc + best_i;
        cur = kids[best_i]; // staged: shared memory; direct: an L1 hit, the line was just read by the owning lane
        const int lw = links[node].y;
        nc = link_n_child(lw);
        if (lane == 0) {
            bool captured;
            key = apply_move(w.board, link_move(lw), key, &captured);
            clock = captured ? 0 : min(clock + 1, KEY_WINDOW - 1);
            w.keys[clock] = key;
            w.board[OFF_CLOCK] = (uint8_t)clock;
        }
        __syncwarp(); // also: every lane is done with the staging buffer before the next bulk copy lands
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
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
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
    int4 *nodes = node_vec(a);
    int2 *links = link_vec(a);
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
        int page = a.d_alloc_page[g], off = a.d_alloc_off[g], np = a.d_n_pages[g];
        const int first = cnt > 0 ? alloc_run(a, cnt, page, off, np, page_list(a, a.d_list_sel[g], g), nullptr, lane) : -1;
        if (first < 0) {
            // last resort (ccz_mcts_reserve before the search makes it unreachable): the leaf stays a leaf
            if (lane == 0 && cnt > 0) {
                a.d_status[g] |= CCZ_STATUS_EXPAND_FAILED;
                atomicAdd(pool_ctl(a) + CCZ_CTL_EXPAND_FAILED, 1ull);
            }
        } else {
            for (int i = lane; i < cnt; i += 32) {
                const int id = move_ids[(size_t)g * MAX_MOVES + i];
                // ids are always valid for reachable positions; a piece on an impossible square
                // (hand-made records) has no action id and gets prior 0 instead of an OOB read
                float p = (id >= 0 && id < N_ACTIONS) ? pol[id] : 0.f;
                if (policy_kind == CCZ_POLICY_LOGITS && id >= 0) p = expf(p - mx) * inv;
                nodes[first + i] = make_int4(0, __float_as_int(0.f), __float_as_int(p), -1);
                links[first + i] = make_int2(node, link_word(id, 0));
            }
            if (lane == 0) {
                reinterpret_cast<int *>(nodes + node)[3] = first;
                links[node].y = link_word(link_move(links[node].y), cnt);
                a.d_n_nodes[g] += cnt;
                a.d_alloc_page[g] = page;
                a.d_alloc_off[g] = off;
                a.d_n_pages[g] = np;
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
            int2 *nq = reinterpret_cast<int2 *>(nodes + cur); // {N, Q}
            const int2 o = *nq;
            const int n = o.x + 1;
            const float q = __int_as_float(o.y);
            *nq = make_int2(n, __float_as_int(__fadd_rn(q, __fdiv_rn(__fsub_rn(x, q), (float)n))));
            x = -x;
            cur = links[cur].x;
        }
    }
}

// root children read-out (mcts.py:163-164)
__global__ void __launch_bounds__(MCTS_WARPS * 32)
mcts_root_visits_kernel(ccz_arena a, int16_t *acts, int32_t *visits, int16_t *counts) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    const int4 *nodes = node_vec(a);
    const int2 *links = link_vec(a);
    const int root = a.d_root[g];
    const int nc = link_n_child(links[root].y);
    const int fc = nodes[root].w;
    for (int i = lane; i < MAX_MOVES; i += 32) {
        const bool ok = i < nc;
        acts[(size_t)g * MAX_MOVES + i] = ok ? (int16_t)link_move(links[fc + i].y) : (int16_t)-1;
        visits[(size_t)g * MAX_MOVES + i] = ok ? nodes[fc + i].x : 0;
    }
    if (lane == 0) counts[g] = (int16_t)nc;
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
    key = __shfl_sync(FULL, key, 0);
    const int rep = count_repeats(keys, clock, key, lane);
    if (lane == 0) B[OFF_REP] = (uint8_t)min(rep, 255);
    __syncwarp();
}

__device__ __forceinline__ uint64_t board_key(const uint8_t *B, int lane) {
    uint64_t k = 0;
    for (int s = lane; s < 90; s += 32)
        if (B[s]) k ^= zkey(B[s], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k ^= __shfl_xor_sync(FULL, k, o);
    if (B[OFF_TURN] == 0) k ^= zkey_turn();
    return k;
}

// start position + its key into game g's root record (whole warp; B = 96 bytes of shared memory)
__device__ __forceinline__ void root_to_start(const ccz_arena &a, int g, uint8_t *B, int lane) {
    if (lane < 6) {
        const uint4 v = reinterpret_cast<const uint4 *>(d_start_board)[lane];
        reinterpret_cast<uint4 *>(a.d_root_boards + (size_t)g * BOARD_BYTES)[lane] = v;
        reinterpret_cast<uint4 *>(B)[lane] = v;
    }
    __syncwarp();
    const uint64_t k = board_key(B, lane);
    if (lane == 0) a.d_root_keys[(size_t)g * KEY_WINDOW] = k;
}

// lane 0: game g restarts with a single unvisited root at slot 0 of `page`
__device__ __forceinline__ void set_fresh_tree(const ccz_arena &a, int g, int page) {
    const int idx = page << a.page_shift;
    fresh_root(a, idx);
    a.d_root[g] = idx;
    a.d_alloc_page[g] = page;
    a.d_alloc_off[g] = 1;
    a.d_n_nodes[g] = 1;
}

struct Compacted {
    int np, page, off, n_nodes;
    bool ok;
};

// Breadth-first copy of the sub-tree under src node `src_root` into pages of dst (whole warp):
// the new nodes are their own queue; child runs stay contiguous, ordered and inside one page.
// `list[0] = first_page` is the first page (already owned); further pages are popped from dst's
// ring and appended to `list`; fill[i] receives the nodes used in page i once the copy moves on.
__device__ __forceinline__ Compacted compact_subtree(const ccz_arena &src, const ccz_arena &dst, int src_root,
                                                     int first_page, volatile int32_t *list, volatile int32_t *fill,
                                                     int lane) {
    const int shift = dst.page_shift;
    const int4 *sn = node_vec(src);
    const int2 *sl = link_vec(src);
    int4 *dn = node_vec(dst);
    int2 *dl = link_vec(dst);
    Compacted r{1, first_page, 1, 1, true};
    if (lane == 0) {
        list[0] = first_page;
        const int2 k = sl[src_root];
        dn[first_page << shift] = sn[src_root]; // first_child: still the OLD index, remapped when dequeued
        dl[first_page << shift] = make_int2(-1, k.y);
    }
    int qp = 0, qo = 0; // queue cursor: page number in `list`, slot in that page
    while (true) {
        __syncwarp();
        const bool last = qp == r.np - 1;
        const int page_end = last ? r.off : fill[qp];
        if (qo >= page_end) {
            if (last) break;
            ++qp;
            qo = 0;
            continue;
        }
        const int qbase = (list[qp] << shift) + qo;
        const bool live = qo + lane < page_end;
        int nc = 0, ofc = -1;
        if (live) {
            nc = link_n_child(__ldcg(dl + qbase + lane).y);
            ofc = __ldcg(dn + qbase + lane).w;
        }
        int my_fc = -1;
        uint32_t todo = __ballot_sync(FULL, nc > 0);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ncl = __shfl_sync(FULL, nc, l);
            const int ofl = __shfl_sync(FULL, ofc, l);
            const int nfl = alloc_run(dst, ncl, r.page, r.off, r.np, list, fill, lane);
            if (nfl < 0) {
                r.ok = false;
                return r;
            }
            r.n_nodes += ncl;
            if (lane == l) my_fc = nfl;
            for (int j = lane; j < ncl; j += 32) {
                const int4 c = sn[ofl + j];
                const int2 k = sl[ofl + j];
                dn[nfl + j] = c;
                dl[nfl + j] = make_int2(qbase + l, k.y);
            }
        }
        if (live && nc > 0) reinterpret_cast<int *>(dn + qbase + lane)[3] = my_fc;
        qo = min(qo + 32, page_end);
    }
    return r;
}

// K7a: MCTS.update_with_move (mcts.py:168-178), pop-only half: advance the root position and its key
// window, compact the chosen child's sub-tree into fresh pages (new page list in the idle half of
// d_page_list).  K7b below returns the old pages and flips the halves.
__global__ void __launch_bounds__(MCTS_WARPS * 32)
mcts_advance_compact_kernel(ccz_arena a, const int16_t *chosen) {
    __shared__ __align__(16) uint8_t s_board[MCTS_WARPS][BOARD_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    const int mv = chosen[g];
    uint8_t *B = s_board[warp];
    const int sel = a.d_list_sel[g];
    const int32_t *old_list = page_list(a, sel, g);
    int32_t *new_list = page_list(a, 1 - sel, g);

    int child = -1;
    if (mv == CCZ_ADVANCE_NEW_GAME) {
        root_to_start(a, g, B, lane);
        if (lane == 0) a.d_status[g] = 0;
    } else if (mv >= 0 && mv < N_ACTIONS) {
        if (lane < 6)
            reinterpret_cast<uint4 *>(B)[lane] =
                reinterpret_cast<const uint4 *>(a.d_root_boards + (size_t)g * BOARD_BYTES)[lane];
        __syncwarp();
        push_with_keys(B, a.d_root_keys + (size_t)g * KEY_WINDOW, mv, lane);
        if (lane < 6)
            reinterpret_cast<uint4 *>(a.d_root_boards + (size_t)g * BOARD_BYTES)[lane] =
                reinterpret_cast<const uint4 *>(B)[lane];
        // locate the chosen child
        const int root = a.d_root[g];
        const int rnc = link_n_child(link_vec(a)[root].y);
        const int rfc = node_vec(a)[root].w;
        for (int base = 0; base < rnc && child < 0; base += 32) {
            const int i = base + lane;
            const uint32_t hit = __ballot_sync(FULL, i < rnc && link_move(link_vec(a)[rfc + i].y) == mv);
            if (hit) child = rfc + base + __ffs(hit) - 1;
        }
    } else if (mv == CCZ_ADVANCE_KEEP) {
        child = a.d_root[g]; // idle slot: same position, same tree (compacted like any other)
    }
    if (child >= 0) {
        int p = -1;
        if (lane == 0) p = pool_pop(a);
        p = __shfl_sync(FULL, p, 0);
        if (p >= 0) {
            const Compacted r = compact_subtree(a, a, child, p, new_list, a.d_page_fill + (size_t)g * a.max_pages, lane);
            if (lane == 0) {
                if (r.ok) {
                    a.d_root[g] = p << a.page_shift;
                    a.d_alloc_page[g] = r.page;
                    a.d_alloc_off[g] = r.off;
                    a.d_n_nodes[g] = r.n_nodes;
                } else { // pool exhausted half-way: fresh root; the pages taken so far stay owned until the next advance
                    set_fresh_tree(a, g, p);
                    a.d_status[g] |= CCZ_STATUS_TREE_DROPPED;
                    atomicAdd(pool_ctl(a) + CCZ_CTL_TREES_DROPPED, 1ull);
                }
                a.d_n_pages_new[g] = r.np;
            }
            return;
        }
        if (lane == 0) { // not a single free page: fresh root in the game's own first page
            a.d_status[g] |= CCZ_STATUS_TREE_DROPPED;
            atomicAdd(pool_ctl(a) + CCZ_CTL_TREES_DROPPED, 1ull);
        }
    }
    // new game, dropped tree or unknown move (mcts.py:177-178): fresh root in the first page, which is kept
    if (lane == 0) {
        const int p0 = old_list[0];
        new_list[0] = p0;
        set_fresh_tree(a, g, p0);
        a.d_n_pages_new[g] = 1;
    }
}

// K7b, push-only half: the pages of the old tree go back to the ring, the page lists swap
__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_advance_release_kernel(ccz_arena a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    const int sel = a.d_list_sel[g];
    pool_push_list(a, page_list(a, sel, g), 0, a.d_n_pages[g], page_list(a, 1 - sel, g)[0], lane);
    if (lane == 0) {
        a.d_list_sel[g] = 1 - sel;
        a.d_n_pages[g] = a.d_n_pages_new[g];
    }
}

// lane-uniform: drop game g's tree in place (fresh root in its first page, other pages pushed)
__device__ __forceinline__ void drop_tree_in_place(const ccz_arena &a, int g, int lane) {
    const int32_t *list = page_list(a, a.d_list_sel[g], g);
    pool_push_list(a, list, 1, a.d_n_pages[g], -1, lane);
    if (lane == 0) {
        set_fresh_tree(a, g, list[0]);
        a.d_n_pages[g] = 1;
    }
}

// Node(None, 1.0) over the start position for the selected games (mcts.py:94; game.py:148); push-only
__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_reset_kernel(ccz_arena a, const uint8_t *mask) {
    __shared__ __align__(16) uint8_t s_board[MCTS_WARPS][BOARD_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    if (mask != nullptr && mask[g] == 0) return;
    root_to_start(a, g, s_board[warp], lane);
    drop_tree_in_place(a, g, lane);
    if (lane == 0) a.d_status[g] = 0;
}

// pre-search guarantee (see ccz_mcts_reserve); push-only
__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_reserve_kernel(ccz_arena a, int pages_per_game) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    const volatile unsigned long long *ctl = pool_ctl(a);
    const long long free_pages = (long long)(ctl[CCZ_CTL_TAIL] - ctl[CCZ_CTL_HEAD]);
    const int np = a.d_n_pages[g];
    const bool low = free_pages < (long long)a.n_games * pages_per_game;
    const int share = a.n_pages / a.n_games - pages_per_game;
    const bool drop = np + pages_per_game > a.max_pages || (low && np > max(share, 1));
    if (!drop) return;
    drop_tree_in_place(a, g, lane);
    if (lane == 0) {
        a.d_status[g] |= CCZ_STATUS_TREE_DROPPED;
        atomicAdd(pool_ctl(a) + CCZ_CTL_TREES_DROPPED, 1ull);
    }
}

// free ring + control words of a new pool: game g owns page g, pages n_games.. are free
__global__ void mcts_pool_init_kernel(ccz_arena a) {
    const int n_free = a.n_pages - a.n_games;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_free; i += gridDim.x * blockDim.x)
        a.d_free_ring[i] = a.n_games + i;
    if (blockIdx.x == 0 && threadIdx.x < CCZ_CTL_WORDS) {
        long long v = 0;
        if (threadIdx.x == CCZ_CTL_TAIL || threadIdx.x == CCZ_CTL_MIN_FREE) v = n_free;
        a.d_pool_ctl[threadIdx.x] = v;
    }
}
__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_games_init_kernel(ccz_arena a) {
    __shared__ __align__(16) uint8_t s_board[MCTS_WARPS][BOARD_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= a.n_games) return;
    root_to_start(a, g, s_board[warp], lane);
    if (lane == 0) {
        a.d_list_sel[g] = 0;
        page_list(a, 0, g)[0] = g;
        a.d_n_pages[g] = 1;
        a.d_n_pages_new[g] = 1;
        a.d_status[g] = 0;
        set_fresh_tree(a, g, g);
    }
}

// whole trees into a larger pool (ccz_mcts_migrate); pops from dst only
__global__ void __launch_bounds__(MCTS_WARPS * 32) mcts_migrate_kernel(ccz_arena src, ccz_arena dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * MCTS_WARPS + warp;
    if (g >= src.n_games) return;
    if (lane < 6)
        reinterpret_cast<uint4 *>(dst.d_root_boards + (size_t)g * BOARD_BYTES)[lane] =
            reinterpret_cast<const uint4 *>(src.d_root_boards + (size_t)g * BOARD_BYTES)[lane];
    for (int i = lane; i < KEY_WINDOW; i += 32)
        dst.d_root_keys[(size_t)g * KEY_WINDOW + i] = src.d_root_keys[(size_t)g * KEY_WINDOW + i];
    if (g == 0 && lane == 0) {
        dst.d_pool_ctl[CCZ_CTL_EXPAND_FAILED] = src.d_pool_ctl[CCZ_CTL_EXPAND_FAILED];
        dst.d_pool_ctl[CCZ_CTL_TREES_DROPPED] = src.d_pool_ctl[CCZ_CTL_TREES_DROPPED];
    }
    int32_t *list = page_list(dst, dst.d_list_sel[g], g);
    const int p0 = list[0];
    const Compacted r = compact_subtree(src, dst, src.d_root[g], p0, list, dst.d_page_fill + (size_t)g * dst.max_pages, lane);
    if (lane == 0) {
        dst.d_status[g] = src.d_status[g];
        if (r.ok) {
            dst.d_root[g] = p0 << dst.page_shift;
            dst.d_alloc_page[g] = r.page;
            dst.d_alloc_off[g] = r.off;
            dst.d_n_nodes[g] = r.n_nodes;
        } else {
            set_fresh_tree(dst, g, p0);
            dst.d_status[g] |= CCZ_STATUS_TREE_DROPPED;
            atomicAdd(pool_ctl(dst) + CCZ_CTL_TREES_DROPPED, 1ull);
        }
        dst.d_n_pages[g] = r.np;
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
