// chunks.cuh — ESVO chunk serialization on the GPU (the step in front of the ray-cast path and the reference's own
// traced CPU hotspot, `serialize_chunk` in src/systems/worldsvo.rs:93-97).
//
// Behavioural contract = SerializedChunk::new + serialize_octant over the chunk octree that Chunk::fill_with builds
// (src/world/hds/esvo.rs:353-383, 439-512; src/world/chunk.rs:125-131): from a dense 32^3 BlockId array (index x + 32*(y + 32*z),
// 0 = air) and a LOD, the 12-word octant records in depth-first pre-order with relative child pointers, byte-identical to the
// host serializer (host/esvo.cpp, itself pinned on the reference's known-answer tests).
//
// Not a transliteration of the recursive serializer: one CTA per chunk builds the occupancy pyramid bottom-up in shared memory
// (child masks per cell for cell edges 2..32), the LOD representatives (pick_leaf_for_lod order 2,3,6,7,0,1,4,5,
// src/world/hds/internal.rs:461-485), subtree sizes bottom-up and pre-order offsets top-down, then every record is written by
// its own thread. Chunks claim their output range with one atomicAdd on a bump pointer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vx {

struct ChunkOut {        // = VxChunkInfo
    unsigned long long offset_bytes, length_bytes;
    uint8_t child_mask, leaf_mask, depth, pad[5];
};

#define VX_CHUNK_THREADS 256

// cells of edge 2^k: n_k = 32 >> k per axis. Shared arrays are laid out level after level; LVL_OFF[k] = first cell of level k
// in a flat array holding levels 1..5 (4096 + 512 + 64 + 8 + 1 = 4681 cells).
__device__ __forceinline__ uint32_t lvl_off(uint32_t k) { return k == 1 ? 0u : k == 2 ? 4096u : k == 3 ? 4608u : k == 4 ? 4672u : 4680u; }
__device__ __forceinline__ uint32_t cell_index(uint32_t k, uint32_t x, uint32_t y, uint32_t z) { const uint32_t n = 32u >> k; return x + n * (y + n * z); }
__device__ __forceinline__ void cell_coords(uint32_t k, uint32_t c, uint32_t& x, uint32_t& y, uint32_t& z) {
    const uint32_t n = 32u >> k; x = c % n; y = (c / n) % n; z = c / (n * n);
}
// child i of cell (x,y,z) at level k is cell (2x + (i&1), 2y + ((i>>1)&1), 2z + ((i>>2)&1)) at level k-1 (octree.rs:21-23)
__device__ __forceinline__ uint32_t child_cell(uint32_t k, uint32_t x, uint32_t y, uint32_t z, uint32_t i) {
    return cell_index(k - 1, 2 * x + (i & 1u), 2 * y + ((i >> 1) & 1u), 2 * z + ((i >> 2) & 1u));
}

struct ChunkSmem {
    uint32_t occ0[1024];      // voxel occupancy bits: bit x of word y + 32 z
    uint8_t cm[4681];         // child mask of every cell of levels 1..5 (bit i = child i occupied)
    uint8_t _pad[3];
    uint16_t size[4681];      // records in the subtree rooted at the cell (0 = empty / below the record levels)
    uint16_t off[4681];       // pre-order index of the cell's record
    uint32_t rep[4681];       // pick_leaf_for_lod representative of the cell (only the levels the LOD needs)
    unsigned long long base;  // output offset of this chunk in words
};

__global__ void __launch_bounds__(VX_CHUNK_THREADS) serialize_chunks_kernel(const uint32_t* __restrict__ blocks, const uint8_t* __restrict__ lods, uint32_t n_chunks,
                                                                            uint32_t* __restrict__ out, unsigned long long out_cap_words,
                                                                            unsigned long long* bump, ChunkOut* infos, unsigned int* overflow) {
    extern __shared__ unsigned char chunk_smem_raw[];
    ChunkSmem& sm = *reinterpret_cast<ChunkSmem*>(chunk_smem_raw);
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n_chunks) return;
    const uint32_t* b = blocks + (size_t)chunk * 32768u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t lod = lods ? lods[chunk] : 0u;
    const uint32_t D = (lod == 0u || lod > 5u) ? 5u : lod;   // record levels: cell levels 5 .. 6 - D; their children at level 5 - D are the leaves
    const uint32_t kr = 6u - D, kl = 5u - D;

    // 1. voxel occupancy: one warp per row of 32 voxels, a ballot per row; 8 rows are in flight per warp (the 128 KB of block
    //    ids are the bulk of the kernel's HBM traffic: one outstanding 128-byte load per warp left the memory system idle)
    for (uint32_t row0 = warp * 8u; row0 < 1024u; row0 += (VX_CHUNK_THREADS / 32) * 8u) {
        uint32_t v[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) v[j] = __ldcs(b + (row0 + j) * 32u + lane);
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            const unsigned m = __ballot_sync(0xffffffffu, v[j] != 0u);
            if (lane == 0) sm.occ0[row0 + j] = m;
        }
    }
    __syncthreads();
    // 2. child masks, level 1 from the voxel bits, levels 2..5 from the level below
    for (uint32_t c = tid; c < 4096u; c += VX_CHUNK_THREADS) {
        uint32_t x, y, z; cell_coords(1, c, x, y, z);
        uint32_t m = 0;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            const uint32_t vx_ = 2 * x + (i & 1u), vy = 2 * y + ((i >> 1) & 1u), vz = 2 * z + ((i >> 2) & 1u);
            m |= ((sm.occ0[vy + 32u * vz] >> vx_) & 1u) << i;
        }
        sm.cm[c] = (uint8_t)m;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= 5; ++k) {
        const uint32_t nk = (32u >> k) * (32u >> k) * (32u >> k);
        for (uint32_t c = tid; c < nk; c += VX_CHUNK_THREADS) {
            uint32_t x, y, z; cell_coords(k, c, x, y, z);
            uint32_t m = 0;
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) m |= (uint32_t)(sm.cm[lvl_off(k - 1) + child_cell(k, x, y, z, i)] != 0) << i;
            sm.cm[lvl_off(k) + c] = (uint8_t)m;
        }
        __syncthreads();
    }
    // 3. LOD representatives up to the leaf level kl (pick_leaf_for_lod: first occupied child in the order 2,3,6,7,0,1,4,5)
    for (uint32_t k = 1; k <= kl; ++k) {
        const uint32_t nk = (32u >> k) * (32u >> k) * (32u >> k);
        for (uint32_t c = tid; c < nk; c += VX_CHUNK_THREADS) {
            const uint32_t m = sm.cm[lvl_off(k) + c];
            uint32_t v = 0;
            if (m) {
                const uint32_t order[8] = {2, 3, 6, 7, 0, 1, 4, 5};
                uint32_t pick = 0;
#pragma unroll
                for (int j = 7; j >= 0; --j) if ((m >> order[j]) & 1u) pick = order[j];
                uint32_t x, y, z; cell_coords(k, c, x, y, z);
                if (k == 1) v = __ldg(b + (2 * x + (pick & 1u)) + 32u * ((2 * y + ((pick >> 1) & 1u)) + 32u * (2 * z + ((pick >> 2) & 1u))));
                else v = sm.rep[lvl_off(k - 1) + child_cell(k, x, y, z, pick)];
            }
            sm.rep[lvl_off(k) + c] = v;
        }
        __syncthreads();
    }
    // 4. subtree sizes bottom-up over the record levels kr .. 5
    for (uint32_t k = kr; k <= 5; ++k) {
        const uint32_t nk = (32u >> k) * (32u >> k) * (32u >> k);
        for (uint32_t c = tid; c < nk; c += VX_CHUNK_THREADS) {
            const uint32_t m = sm.cm[lvl_off(k) + c];
            uint32_t s = 0;
            if (m) {
                s = 1;
                if (k > kr) {
                    uint32_t x, y, z; cell_coords(k, c, x, y, z);
#pragma unroll
                    for (uint32_t i = 0; i < 8; ++i) s += sm.size[lvl_off(k - 1) + child_cell(k, x, y, z, i)];
                }
            }
            sm.size[lvl_off(k) + c] = (uint16_t)s;
        }
        __syncthreads();
    }
    // 5. pre-order offsets top-down; claim the output range
    const uint32_t total = sm.size[lvl_off(5)];
    if (tid == 0) {
        sm.off[lvl_off(5)] = 0;
        sm.base = total ? atomicAdd(bump, (unsigned long long)total * 12ull) : 0ull;
    }
    __syncthreads();
    for (uint32_t k = 5; k > kr; --k) {
        const uint32_t nk = (32u >> k) * (32u >> k) * (32u >> k);
        for (uint32_t c = tid; c < nk; c += VX_CHUNK_THREADS) {
            const uint32_t m = sm.cm[lvl_off(k) + c];
            if (!m) continue;
            uint32_t x, y, z; cell_coords(k, c, x, y, z);
            uint32_t running = (uint32_t)sm.off[lvl_off(k) + c] + 1u;
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) {
                const uint32_t cc = lvl_off(k - 1) + child_cell(k, x, y, z, i);
                sm.off[cc] = (uint16_t)running;
                running += sm.size[cc];
            }
        }
        __syncthreads();
    }
    const unsigned long long base = sm.base;
    const bool fits = base + (unsigned long long)total * 12ull <= out_cap_words;
    if (tid == 0) {
        ChunkOut o;
        o.offset_bytes = base * 4ull; o.length_bytes = (unsigned long long)total * 48ull;
        const uint32_t root = sm.cm[lvl_off(5)];
        o.child_mask = (uint8_t)root; o.leaf_mask = (uint8_t)(D == 1 ? root : 0); o.depth = (uint8_t)(total ? D : 0);
        for (int k = 0; k < 5; ++k) o.pad[k] = 0;
        infos[chunk] = o;
        if (!fits) atomicAdd(overflow, 1u);
    }
    if (!fits || !total) return;
    // 6. one thread per record: header masks of the child octants, relative pointers / leaf values (esvo.rs:74-101, 439-512)
    for (uint32_t k = kr; k <= 5; ++k) {
        const uint32_t nk = (32u >> k) * (32u >> k) * (32u >> k);
        for (uint32_t c = tid; c < nk; c += VX_CHUNK_THREADS) {
            const uint32_t m = sm.cm[lvl_off(k) + c];
            if (!m) continue;
            uint32_t x, y, z; cell_coords(k, c, x, y, z);
            const uint32_t self = sm.off[lvl_off(k) + c];
            uint32_t rec[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) rec[j] = 0;
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) {
                if (!((m >> i) & 1u)) continue;
                if (k == kr) {                                   // the children are the leaves of this LOD
                    uint32_t v;
                    if (kl == 0) v = __ldg(b + (2 * x + (i & 1u)) + 32u * ((2 * y + ((i >> 1) & 1u)) + 32u * (2 * z + ((i >> 2) & 1u))));
                    else v = sm.rep[lvl_off(kl) + child_cell(k, x, y, z, i)];
                    rec[4 + i] = v;
                } else {
                    const uint32_t cc = lvl_off(k - 1) + child_cell(k, x, y, z, i);
                    const uint32_t cmask = sm.cm[cc];
                    uint32_t mask = (cmask << 8) | (k - 1 == kr ? cmask : 0u);   // (child_mask << 8) | leaf_mask of that child
                    if (i & 1u) mask <<= 16;
                    rec[i >> 1] |= mask;
                    rec[4 + i] = (12u * ((uint32_t)sm.off[cc] - self) - 4u - i) | 0x80000000u;
                }
            }
            uint32_t* dst = out + base + 12ull * self;
#pragma unroll
            for (int j = 0; j < 12; ++j) dst[j] = rec[j];
        }
    }
}

}  // namespace vx
