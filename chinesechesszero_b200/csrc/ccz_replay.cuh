// ccz_replay.cuh -- K8: replay densification (CollectPipeline.preprocess / flip_data,
// collect.py:64-131): the (17,7,10,9) float16 state stack of each sample and its file-mirrored
// twin (np.flip(state, axis=2), collect.py:125-127), and the dense float64 visit distribution
// with its mirror (mcts_prob[flip_map], collect.py:117-128).  Rows [0,n) are the originals,
// rows [n,2n) the mirrored copies (collect.py:131: data + data_flip).
#pragma once
#include "ccz_movegen.cuh"

namespace ccz {

// one thread per 32-bit word (two float16 elements) of the [2n,10710] output
__global__ void __launch_bounds__(256)
replay_states_kernel(const uint8_t *__restrict__ hist /*[n,8,96]*/, const uint8_t *__restrict__ turn_plane, int n,
                     uint32_t *__restrict__ out) {
    const long long wid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (wid >= 2ll * n * WORDS_PER_POS) return;
    const int row = (int)(wid / WORDS_PER_POS);
    const int off = (int)(wid - (long long)row * WORDS_PER_POS);
    const bool mirrored = row >= n;
    const int i = mirrored ? row - n : row;
    const int p = off / 45, d = off - 45 * p;
    const int play = p / 7, ch = p - 7 * play;
    uint32_t v = 0u;
    if (play == 16) {
        v = turn_plane[i] ? 0x3C003C00u : 0u; // collect.py:78-81
    } else {
        const uint8_t *B = hist + ((size_t)i * 8 + (play & 7)) * BOARD_BYTES;
        const uint32_t code = (uint32_t)(ch + 1) | (play >= 8 ? 8u : 0u);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int e = 2 * d + h;
            const int r = e / 9, f = e - 9 * r;
            const int s = mirrored ? r * 9 + (8 - f) : e;
            if (B[s] == code) v |= 0x3C00u << (16 * h);
        }
    }
    out[wid] = v;
}

// one warp per sample: scatter the sparse (acts, probs) into the dense rows (pre-zeroed)
__global__ void __launch_bounds__(256)
replay_pi_kernel(const int16_t *__restrict__ acts, const double *__restrict__ probs,
                 const int16_t *__restrict__ counts, int n, double *__restrict__ pi) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int c = counts[i];
    for (int k = lane; k < c; k += 32) {
        const int a = acts[(size_t)i * MAX_MOVES + k];
        if (a < 0 || a >= N_ACTIONS) continue;
        const double p = probs[(size_t)i * MAX_MOVES + k];
        pi[(size_t)i * N_ACTIONS + a] = p;
        pi[(size_t)(n + i) * N_ACTIONS + d_flip_of[a]] = p;
    }
}

} // namespace ccz
