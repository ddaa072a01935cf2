// One-shot exchanges of the class-sharded head over peer-mapped memory (NVLink / NVSwitch), replacing the three
// NCCL collectives of a step -- all-gather of the local embeddings + labels, all-gather of the per-row softmax
// statistics, reduce-scatter of the embedding gradient -- whose messages (0.1 - 1 MB) are latency-bound:
// every rank STORES its slice straight into the receive buffer of every peer, raises a flag there, and waits for
// the peers' flags.  No ring, no staging: one NVLink store latency plus one flag round trip.
//
// Reference counterpart: the scatter / gather / reduce traffic of nn.DataParallel
// (nlp_classifier_train_daodian_v2_dist.py:85), see sharded.py.
//
// Memory: `peer_bufs[r]` / `peer_flags[r]` are rank r's receive buffer and flag array as mapped into THIS
// process (torch.distributed._symmetric_memory on the host side).  Receive slot of sender s in a buffer:
// base + s * slot_stride.  Flags: uint32 [channels][ARCFACE_B200_MAX_RANKS], monotonically increasing call numbers.
// `sync` (local device memory, zeroed once): [0] call number of this channel, [1] CTA arrival counter.
// Buffer reuse is safe without a second handshake because a step runs the three exchanges in order on
// different channels: a rank can only start exchange k of step t + 1 after every peer has passed exchange
// k + 1 of step t, i.e. after every peer's consumer of exchange k (stream order) has been launched and,
// for the data to be overwritten, completed -- the consumer kernels precede the next exchange on the stream.
#include "host_util.h"

#include <stdint.h>

namespace ab {

constexpr int P2P_MAX_RANKS = 16;
// ~ seconds.  A peer that never shows up (crashed rank, mismatched call sequence) must not hang the GPU, and must not
// kill this process's CUDA context either: the waiting thread gives up, raises the caller's error word (nullable; may
// be mapped pinned host memory) and the kernel finishes normally.  The step's results are then meaningless; the host
// side (p2p.PeerExchange) examines the word before its next exchange and raises.
constexpr unsigned P2P_SPIN_LIMIT = 400000000u;

struct P2PParams {
    unsigned long long bufs[P2P_MAX_RANKS];
    unsigned long long flags[P2P_MAX_RANKS];
    const unsigned char* src;
    unsigned long long bytes_per_peer;   // multiple of 16
    unsigned long long src_stride;       // 0: the same bytes go to every peer (all-gather); else peer r gets src + r * stride
    unsigned long long slot_stride;
    unsigned long long split;            // all-gather only, 0 = off: the first `split` bytes of every rank's message land
                                         // contiguously in rank order ([world][split]), the remaining bytes likewise
                                         // behind them -- the gathered (x | labels) arrive as one x matrix and one
                                         // label vector, no unpacking copies
    int rank, world, channel;
    unsigned* sync;                      // this channel's [call number, arrival counter]
    unsigned* error_word;                // nullable: set to 1 + channel when a peer's flag never arrives
};

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) p2p_exchange_kernel(const P2PParams p) {
    __shared__ unsigned s_last;
    const unsigned call = p.sync[0] + 1u;  // every CTA reads it before the last one bumps it (see below)
    // ---- copy: (peer, 16-byte chunk) pairs spread over the grid; the local slot is written like any other
    const unsigned long long chunks = p.bytes_per_peer >> 4;
    const unsigned long long total = chunks * static_cast<unsigned long long>(p.world);
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += stride) {
        const int peer = static_cast<int>(i / chunks);
        const unsigned long long c = i - static_cast<unsigned long long>(peer) * chunks;
        const uint4 v = *reinterpret_cast<const uint4*>(p.src + static_cast<unsigned long long>(peer) * p.src_stride + (c << 4));
        const unsigned long long o = c << 4;
        unsigned long long dst = static_cast<unsigned long long>(p.rank) * p.slot_stride + o;
        if (p.split != 0ull)
            dst = o < p.split ? static_cast<unsigned long long>(p.rank) * p.split + o
                              : static_cast<unsigned long long>(p.world) * p.split +
                                    static_cast<unsigned long long>(p.rank) * (p.bytes_per_peer - p.split) + (o - p.split);
        *reinterpret_cast<uint4*>(p.bufs[peer] + dst) = v;
    }
    __threadfence_system();  // this thread's peer stores are performed before anything it does next
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(p.sync + 1, 1u);
        s_last = (prev == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last == 0u) return;
    // ---- last CTA: every CTA's stores are out.  Raise our flag on every rank, then wait for every rank's flag here.
    __threadfence_system();
    if (threadIdx.x < p.world) {
        unsigned* remote = reinterpret_cast<unsigned*>(p.flags[threadIdx.x]) + p.channel * P2P_MAX_RANKS + p.rank;
        st_release_sys_u32(remote, call);
        const unsigned* local = reinterpret_cast<const unsigned*>(p.flags[p.rank]) + p.channel * P2P_MAX_RANKS + threadIdx.x;
        unsigned spins = 0;
        while (static_cast<int>(ld_acquire_sys_u32(local) - call) < 0) {
            if (++spins > P2P_SPIN_LIMIT) {
                if (p.error_word != nullptr) *reinterpret_cast<volatile unsigned*>(p.error_word) = 1u + p.channel;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.sync[1] = 0u;
        p.sync[0] = call;
    }
}

// dx[b] = (g[b] - (xhat[b] . g[b]) xhat[b]) * inv_nx[b] with g[b] = sum over parts of parts[r][b] (fixed rank order:
// deterministic), xhat = x * inv_nx in fp32.  The reduce half of the reduce-scatter, fused into the normalise backward.
__global__ void __launch_bounds__(256)
normalize_bwd_x_sum_kernel(const float* __restrict__ x, const float* __restrict__ inv_nx, const float* __restrict__ parts,
                           int n_parts, int64_t part_stride, int B, int D, float* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const float inv = inv_nx[b];
    const float* xp = x + static_cast<int64_t>(b) * D;
    float* op = dx + static_cast<int64_t>(b) * D;
    float r = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < n_parts; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(parts + q * part_stride + static_cast<int64_t>(b) * D + d);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        *reinterpret_cast<float4*>(op + d) = g;  // parked in the output row, finished below
        const float4 a = *reinterpret_cast<const float4*>(xp + d);
        r += (a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w);
    }
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    r *= inv;  // xhat . g
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(xp + d);
        const float4 g = *reinterpret_cast<const float4*>(op + d);  // written by this lane above
        float4 o;
        o.x = (g.x - r * (a.x * inv)) * inv;
        o.y = (g.y - r * (a.y * inv)) * inv;
        o.z = (g.z - r * (a.z * inv)) * inv;
        o.w = (g.w - r * (a.w * inv)) * inv;
        *reinterpret_cast<float4*>(op + d) = o;
    }
}

}  // namespace ab

using namespace ab;

extern "C" int32_t arcface_b200_p2p_exchange(const void* src, size_t bytes_per_peer, size_t src_stride,
                                             const uint64_t* peer_bufs, const uint64_t* peer_flags, int32_t rank,
                                             int32_t world, size_t slot_stride, int32_t channel, uint32_t* sync_dev,
                                             uint32_t* error_word, void* stream) {
    return arcface_b200_p2p_gather_split(src, bytes_per_peer, src_stride, 0, peer_bufs, peer_flags, rank, world,
                                         slot_stride, channel, sync_dev, error_word, stream);
}

extern "C" int32_t arcface_b200_p2p_gather_split(const void* src, size_t bytes_per_peer, size_t src_stride, size_t split,
                                                 const uint64_t* peer_bufs, const uint64_t* peer_flags, int32_t rank,
                                                 int32_t world, size_t slot_stride, int32_t channel, uint32_t* sync_dev,
                                                 uint32_t* error_word, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(split % 16 == 0 && split <= bytes_per_peer && (split == 0 || src_stride == 0), ARCFACE_B200_E_LAYOUT,
               "p2p_gather_split: split must be a multiple of 16 bytes inside an all-gather message");
    AB_REQUIRE(src && peer_bufs && peer_flags && sync_dev, ARCFACE_B200_E_ARG, "p2p_exchange: null pointer");
    AB_REQUIRE(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world && channel >= 0 && channel < 8,
               ARCFACE_B200_E_ARG, "p2p_exchange: bad rank / world / channel");
    AB_REQUIRE(bytes_per_peer > 0 && bytes_per_peer % 16 == 0 && src_stride % 16 == 0 && slot_stride % 16 == 0 &&
                   slot_stride >= bytes_per_peer && aligned16(src),
               ARCFACE_B200_E_LAYOUT, "p2p_exchange: sizes and strides must be multiples of 16 bytes");
    P2PParams p;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) {
        p.bufs[r] = r < world ? peer_bufs[r] : 0ull;
        p.flags[r] = r < world ? peer_flags[r] : 0ull;
        AB_REQUIRE(r >= world || (p.bufs[r] != 0 && p.flags[r] != 0 && (p.bufs[r] & 15ull) == 0), ARCFACE_B200_E_ARG,
                   "p2p_exchange: peer %d has no mapped buffer", r);
    }
    p.src = static_cast<const unsigned char*>(src);
    p.bytes_per_peer = bytes_per_peer;
    p.src_stride = src_stride;
    p.slot_stride = slot_stride;
    p.split = split;
    p.rank = rank; p.world = world; p.channel = channel;
    p.sync = sync_dev + 2 * channel;
    p.error_word = error_word;
    const unsigned long long total16 = (bytes_per_peer >> 4) * static_cast<unsigned long long>(world);
    unsigned long long want = (total16 + 255ull) / 256ull;   // one 16-byte chunk per thread ...
    const int grid = static_cast<int>(want < 1 ? 1 : (want > 64 ? 64 : want));  // ... up to 64 CTAs
    p2p_exchange_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_normalize_bwd_x_sum(const float* x, const float* inv_nx, const float* parts,
                                                    int32_t n_parts, int64_t part_stride, int32_t B, int32_t D,
                                                    float* dx, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(x && inv_nx && parts && dx, ARCFACE_B200_E_ARG, "normalize_bwd_x_sum: null pointer");
    AB_REQUIRE(B >= 0 && D >= 8 && D % 8 == 0 && n_parts >= 1 && part_stride >= static_cast<int64_t>(B) * D &&
                   part_stride % 4 == 0,
               ARCFACE_B200_E_SHAPE, "normalize_bwd_x_sum: bad shape");
    AB_REQUIRE(aligned16(x) && aligned16(parts) && aligned16(dx), ARCFACE_B200_E_LAYOUT,
               "normalize_bwd_x_sum: pointers must be 16-byte aligned");
    if (B == 0) return ARCFACE_B200_OK;
    // one warp per row; small local batches (64 rows per rank at 8 GPUs) get two-warp CTAs so that the rows spread
    // over 32 SMs instead of 8
    const int wpb = B <= 256 ? 2 : 8;
    normalize_bwd_x_sum_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        x, inv_nx, parts, n_parts, part_stride, B, D, dx);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}
