// libvxgpu: chunk wire format on the device.  Replaces Chunk::encode / decode
// (voxels/Chunk.cpp:70-97) and Lightmap::encode / decode (lighting/Lightmap.cpp:
// 18-34) so that chunks move between region files / the lights cache and HBM
// without a CPU reformat (GlobalChunks::create, voxels/GlobalChunks.cpp:91-132;
// WorldRegions, world/files/WorldRegions.cpp:155-242).
//   wire voxels: u16 ids[65536] then u16 states[65536], little endian
//   wire lights: one byte per two voxels: S(2i) | S(2i+1) << 4
// All four kernels are pure streams: 16-byte loads and stores, grid (8, chunk).
#include <cstring>
#include <vector>

#include "vx_ctx.hpp"

using namespace vx;

namespace {

constexpr int ENC_CHUNK = VX_CHUNK_DATA_LEN + VX_LIGHTMAP_DATA_LEN;  // staging bytes per chunk

// 8 voxels per thread: one uint4 of ids + one uint4 of states -> two uint4 of voxels
__global__ void __launch_bounds__(256) k_decode_voxels(DevWorld W, const int32_t* pools,
                                                       const uint8_t* stage) {
    const int p = pools[blockIdx.y];
    const uint4* ids = reinterpret_cast<const uint4*>(stage + (size_t)blockIdx.y * ENC_CHUNK);
    const uint4* sts = ids + CVOL / 8;
    uint4* out = reinterpret_cast<uint4*>(W.vox_of(p));
    for (int u = blockIdx.x * 1024 + threadIdx.x; u < (blockIdx.x + 1) * 1024; u += 256) {
        const uint4 a = ids[u], b = sts[u];
        out[2 * u] = make_uint4((a.x & 0xFFFFu) | (b.x << 16), (a.x >> 16) | (b.x & 0xFFFF0000u),
                                (a.y & 0xFFFFu) | (b.y << 16), (a.y >> 16) | (b.y & 0xFFFF0000u));
        out[2 * u + 1] = make_uint4((a.z & 0xFFFFu) | (b.z << 16), (a.z >> 16) | (b.z & 0xFFFF0000u),
                                    (a.w & 0xFFFFu) | (b.w << 16), (a.w >> 16) | (b.w & 0xFFFF0000u));
    }
}

// 32 voxels per thread: one uint4 of packed sky nibbles -> four uint4 of lights
__global__ void __launch_bounds__(256) k_decode_lights(DevWorld W, const int32_t* pools,
                                                       const uint8_t* stage) {
    const int p = pools[blockIdx.y];
    const uint4* in = reinterpret_cast<const uint4*>(stage + (size_t)blockIdx.y * ENC_CHUNK + VX_CHUNK_DATA_LEN);
    uint4* out = reinterpret_cast<uint4*>(W.light_of(p));
    const int u = blockIdx.x * 256 + threadIdx.x;  // 2048 uint4 per chunk
    const uint4 q = in[u];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        // byte j of w[k] holds voxels 2j (low nibble) and 2j+1 (high nibble)
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = (w[k] >> (8 * j)) & 0xFFu;
            o[j] = ((b & 0xFu) << 12) | ((b & 0xF0u) << 24);
        }
        out[4 * u + k] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void __launch_bounds__(256) k_encode_voxels(DevWorld W, const int32_t* pools, uint8_t* stage) {
    const int p = pools[blockIdx.y];
    uint4* ids = reinterpret_cast<uint4*>(stage + (size_t)blockIdx.y * ENC_CHUNK);
    uint4* sts = ids + CVOL / 8;
    if (p < 0) {
        for (int u = blockIdx.x * 1024 + threadIdx.x; u < (blockIdx.x + 1) * 1024; u += 256)
            ids[u] = sts[u] = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint4* in = reinterpret_cast<const uint4*>(W.vox_of(p));
    for (int u = blockIdx.x * 1024 + threadIdx.x; u < (blockIdx.x + 1) * 1024; u += 256) {
        const uint4 a = in[2 * u], b = in[2 * u + 1];
        ids[u] = make_uint4((a.x & 0xFFFFu) | (a.y << 16), (a.z & 0xFFFFu) | (a.w << 16),
                            (b.x & 0xFFFFu) | (b.y << 16), (b.z & 0xFFFFu) | (b.w << 16));
        sts[u] = make_uint4((a.x >> 16) | (a.y & 0xFFFF0000u), (a.z >> 16) | (a.w & 0xFFFF0000u),
                            (b.x >> 16) | (b.y & 0xFFFF0000u), (b.z >> 16) | (b.w & 0xFFFF0000u));
    }
}

__global__ void __launch_bounds__(256) k_encode_lights(DevWorld W, const int32_t* pools, uint8_t* stage) {
    const int p = pools[blockIdx.y];
    uint4* out = reinterpret_cast<uint4*>(stage + (size_t)blockIdx.y * ENC_CHUNK + VX_CHUNK_DATA_LEN);
    const int u = blockIdx.x * 256 + threadIdx.x;
    if (p < 0) {
        out[u] = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint4* in = reinterpret_cast<const uint4*>(W.light_of(p));
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint4 q = in[4 * u + k];  // 8 lights -> 4 bytes
        const uint32_t v[4] = {q.x, q.y, q.z, q.w};
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            o |= (((v[j] >> 12) & 0xFu) | ((v[j] >> 24) & 0xF0u)) << (8 * j);
        w[k] = o;
    }
    out[u] = make_uint4(w[0], w[1], w[2], w[3]);
}

int grow_stage(vx_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->codec_stage_bytes) return 0;
    cudaFree(ctx->d_codec_stage);
    ctx->d_codec_stage = nullptr;
    ctx->codec_stage_bytes = 0;
    VX_CUDA(cudaMalloc((void**)&ctx->d_codec_stage, bytes + bytes / 4));
    ctx->codec_stage_bytes = bytes + bytes / 4;
    return 0;
}

}  // namespace

extern "C" {

int vx_chunks_put_encoded(vx_ctx* ctx, const vx_chunk_encoded* chunks, int32_t n) {
    if (!ctx) return -1;
    if (ctx->pool_cap == 0) {
        ctx->fail("vx_chunks_put_encoded: configure the window first");
        return -1;
    }
    if (n <= 0) return 0;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    std::vector<vx_chunk_key> keys(n);
    for (int i = 0; i < n; i++) keys[i] = vx_chunk_key {chunks[i].cx, chunks[i].cz};
    std::vector<uint8_t> has_light(n);
    for (int i = 0; i < n; i++) has_light[i] = chunks[i].lights != nullptr;
    if (grow_stage(ctx, (size_t)n * ENC_CHUNK)) return -1;
    for (int i = 0; i < n; i++) {
        uint8_t* dst = ctx->d_codec_stage + (size_t)i * ENC_CHUNK;
        VX_CUDA(cudaMemcpyAsync(dst, chunks[i].voxels, VX_CHUNK_DATA_LEN, cudaMemcpyHostToDevice, st));
        if (chunks[i].lights) {
            VX_CUDA(cudaMemcpyAsync(dst + VX_CHUNK_DATA_LEN, chunks[i].lights, VX_LIGHTMAP_DATA_LEN,
                                    cudaMemcpyHostToDevice, st));
        } else {
            VX_CUDA(cudaMemsetAsync(dst + VX_CHUNK_DATA_LEN, 0, VX_LIGHTMAP_DATA_LEN, st));
        }
    }
    return vx_put_staged(ctx, keys.data(), has_light.data(), n);
}

int vx_chunks_encode(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, uint8_t* voxels_out,
                     uint8_t* lights_out) {
    if (!ctx) return -1;
    if (n <= 0) return 0;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    cudaStream_t st = ctx->stream;
    if (vx_encode_staged(ctx, keys, n, voxels_out != nullptr, lights_out != nullptr)) return -1;
    for (int i = 0; i < n; i++) {
        const uint8_t* src = ctx->d_codec_stage + (size_t)i * ENC_CHUNK;
        if (voxels_out)
            VX_CUDA(cudaMemcpyAsync(voxels_out + (size_t)i * VX_CHUNK_DATA_LEN, src, VX_CHUNK_DATA_LEN,
                                    cudaMemcpyDeviceToHost, st));
        if (lights_out)
            VX_CUDA(cudaMemcpyAsync(lights_out + (size_t)i * VX_LIGHTMAP_DATA_LEN, src + VX_CHUNK_DATA_LEN,
                                    VX_LIGHTMAP_DATA_LEN, cudaMemcpyDeviceToHost, st));
    }
    VX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"

// Chunk::encode / Lightmap::encode of a batch into the staging area (chunk i at
// i * ENC_CHUNK: voxels, then lights); missing chunks encode as zeros.
int vx_encode_staged(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, bool voxels, bool lights) {
    cudaStream_t st = ctx->stream;
    std::vector<int> pools(n);
    for (int i = 0; i < n; i++) pools[i] = ctx->pool_of(keys[i].cx, keys[i].cz);
    if (n > ctx->pool_cap) {
        ctx->fail("vx_encode_staged: more keys than window slots");
        return -1;
    }
    if (grow_stage(ctx, (size_t)n * ENC_CHUNK)) return -1;
    if (lights && vx_materialize(ctx, pools)) return -1;
    // list slot 1 is free outside vx_light_build
    if (vx_upload_list(ctx, pools, 1)) return -1;
    const int32_t* d_pools = ctx->d_list + ctx->pool_cap;
    if (voxels)
        VX_LAUNCH(ctx, KID_ENCODE, k_encode_voxels<<<dim3(8, n), 256, 0, st>>>(ctx->W, d_pools, ctx->d_codec_stage));
    if (lights)
        VX_LAUNCH(ctx, KID_ENCODE, k_encode_lights<<<dim3(8, n), 256, 0, st>>>(ctx->W, d_pools, ctx->d_codec_stage));
    VX_CUDA(cudaGetLastError());
    return 0;
}

// Chunks::putChunk of n chunks whose wire-format bytes are in the staging area:
// Chunk::decode + Lightmap::decode on the device, then the derived tables.
int vx_put_staged(vx_ctx* ctx, const vx_chunk_key* keys, const uint8_t* has_light, int32_t n) {
    cudaStream_t st = ctx->stream;
    std::vector<int> pools;
    if (vx_put_slots(ctx, keys, has_light, n, pools)) return -1;
    if (vx_upload_list(ctx, pools, 0)) return -1;
    VX_LAUNCH(ctx, KID_DECODE, k_decode_voxels<<<dim3(8, n), 256, 0, st>>>(ctx->W, ctx->d_list, ctx->d_codec_stage));
    VX_LAUNCH(ctx, KID_DECODE, k_decode_lights<<<dim3(8, n), 256, 0, st>>>(ctx->W, ctx->d_list, ctx->d_codec_stage));
    VX_CUDA(cudaGetLastError());
    // border planes of the decoded lights + derived tables (a zero lightmap decodes to zeros)
    if (vx_derive_chunks(ctx, pools, true)) return -1;
    VX_CUDA(cudaStreamSynchronize(st));
    return 0;
}
