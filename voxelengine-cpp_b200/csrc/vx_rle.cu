// libvxgpu: the region-file compression of the voxels / lights layers on the
// device.  WorldRegions stores the Chunk::encode bytes with extrle::encode16
// and the Lightmap::encode bytes with extrle::encode (world/files/
// WorldRegions.cpp:60-72, coders/compression.cpp:58-100, coders/rle.cpp:93-205),
// so region-file blobs can go to HBM as they are (a terrain chunk is a few KiB
// instead of 256 KiB on PCIe) and come back compressed.
//
//   extrle16 run: [len & 0x3F | wide << 6 | ext << 7] [len >> 6 if ext] [value lo] [value hi if wide]
//   extrle8  run: [len & 0x7F | ext << 7] [len >> 7 if ext] [value]
//   a run repeats `value` len + 1 times; len <= 0x3FFF (16) / 0x7FFF (8)
//
// Decoding is the only sequential-looking step: where a run starts depends on
// every run before it.  A run is 2..4 bytes, so a 32-byte segment entered at
// byte offset ("phase") 0..3 is parsed speculatively for all four phases; the
// phase -> exit-phase maps compose associatively, a block scan of the maps
// gives every segment its true phase, a second scan the output offsets, and the
// segment is parsed once more to write.  Long runs are queued in shared memory
// and filled by whole warps with 16-byte stores.
// Encoding needs no speculation: a run ends where the value changes or at every
// max-length multiple from the last change, so each element knows from a
// max-scan (last change) whether it closes a run and how long that run is; a
// sum-scan of the encoded sizes gives the byte offsets.
#include <cstring>
#include <string>
#include <vector>

#include "vx_ctx.hpp"

using namespace vx;

namespace {

constexpr int ENC_CHUNK = VX_CHUNK_DATA_LEN + VX_LIGHTMAP_DATA_LEN;  // staging bytes per chunk (vx_codec.cu)

struct RleStream {
    unsigned long long src;  // byte offset of the input in the source buffer
    unsigned long long dst;  // byte offset of the output in the destination buffer
    uint32_t len;            // input bytes
    uint32_t expect;         // decode: decoded bytes the caller expects
};

// ---------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------
constexpr int SEG = 32;              // bytes per speculative parse segment
constexpr int WIN_SEGS = 256;        // segments per window: one per thread
constexpr int WIN = SEG * WIN_SEGS;  // 8 KiB of compressed input per window
constexpr int LONG_RUN = 8;          // runs above this go to the warp-fill queue

template <bool WIDE>
__device__ __forceinline__ void parse_run(const uint8_t* b, int pos, int& size, uint32_t& count,
                                          uint32_t& value) {
    const uint32_t h = b[pos];
    int i = pos + 1;
    uint32_t len;
    if (WIDE) {  // extrle::decode16, coders/rle.cpp:137-151
        len = h & 0x3Fu;
        if (h & 0x80u) len |= (uint32_t)b[i++] << 6;
        value = b[i++];
        if (h & 0x40u) value |= (uint32_t)b[i++] << 8;
    } else {  // extrle::decode, coders/rle.cpp:96-102
        len = h & 0x7Fu;
        if (h & 0x80u) len |= (uint32_t)b[i++] << 7;
        value = b[i++];
    }
    size = i - pos;
    count = len + 1;
}

// phase maps: 2 bits of exit phase per entry phase
__device__ __forceinline__ uint32_t map_apply(uint32_t m, uint32_t p) { return (m >> (2 * p)) & 3u; }
__device__ __forceinline__ uint32_t map_then(uint32_t first, uint32_t second) {
    uint32_t r = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) r |= map_apply(second, map_apply(first, p)) << (2 * p);
    return r;
}
constexpr uint32_t MAP_ID = 0xE4u;  // 3,2,1,0

template <typename T>
__device__ __forceinline__ void warp_fill(T* p, uint32_t count, uint32_t value, int lane) {
    constexpr uint32_t PER = 16 / sizeof(T);
    uint32_t head = (uint32_t)(((16u - (uint32_t)((uintptr_t)p & 15u)) & 15u) / sizeof(T));
    if (head > count) head = count;
    if ((uint32_t)lane < head) p[lane] = (T)value;
    p += head;
    count -= head;
    const uint32_t v32 = sizeof(T) == 2 ? value * 0x10001u : value * 0x01010101u;
    const uint4 q = make_uint4(v32, v32, v32, v32);
    const uint32_t nq = count / PER;
    uint4* p4 = reinterpret_cast<uint4*>(p);
    for (uint32_t i = lane; i < nq; i += 32) p4[i] = q;
    p += nq * PER;
    count -= nq * PER;
    if ((uint32_t)lane < count) p[lane] = (T)value;
}

// grid (n_streams), 256 threads.  status[2 * s] = 0 ok / 1 malformed or wrong
// size, status[2 * s + 1] = decoded bytes (what the reference's error message names).
template <bool WIDE>
__global__ void __launch_bounds__(256) k_rle_decode(const uint8_t* in, uint8_t* out,
                                                    const RleStream* streams, uint32_t* status) {
    using T = typename std::conditional<WIDE, uint16_t, uint8_t>::type;
    const RleStream S = streams[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* src = in + S.src;
    T* dst = reinterpret_cast<T*>(out + S.dst);
    const uint32_t expect = S.expect / sizeof(T);  // elements
    __shared__ uint8_t buf[WIN + 16];
    __shared__ uint32_t words[WIN_SEGS][4];
    __shared__ uint2 queue[WIN / 2];
    __shared__ uint32_t s_wmap[8], s_wsum[8];
    __shared__ uint32_t s_qn, s_bad;
    if (tid == 0) s_bad = 0;
    uint32_t phase_in = 0;   // phase entering the window (uniform)
    uint32_t done = 0;       // elements decoded so far (uniform)
    for (uint32_t w0 = 0; w0 < S.len; w0 += WIN) {
        const int wn = (int)min((uint32_t)WIN, S.len - w0);  // bytes of this window
        __syncthreads();
        for (int i = tid; i < WIN + 16; i += 256) buf[i] = (w0 + i < S.len) ? src[w0 + i] : (uint8_t)0;
        if (tid == 0) s_qn = 0;
        __syncthreads();
        // ---- speculative parse of segment `tid` from all four phases
        const int lo = tid * SEG, hi = min(lo + SEG, wn);
        uint32_t m = MAP_ID;
        if (lo < wn) {
            m = 0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                int pos = lo + p;
                uint32_t cnt = 0;
                while (pos < hi) {
                    int size;
                    uint32_t c, v;
                    parse_run<WIDE>(buf, pos, size, c, v);
                    cnt += c;
                    pos += size;
                }
                words[tid][p] = cnt;
                m |= (uint32_t)min(pos - hi, 3) << (2 * p);
            }
        } else {
            words[tid][0] = words[tid][1] = words[tid][2] = words[tid][3] = 0;
        }
        // ---- scan of the phase maps -> true entry phase of every segment
        uint32_t inc = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc = map_then(o, inc);
        }
        if (lane == 31) s_wmap[warp] = inc;
        __syncthreads();
        uint32_t pre = MAP_ID;  // composition of the warps before this one
        for (int k = 0; k < warp; k++) pre = map_then(pre, s_wmap[k]);
        uint32_t exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
        if (lane == 0) exc = MAP_ID;
        const uint32_t phase = map_apply(map_then(pre, exc), phase_in);
        uint32_t all = MAP_ID;
        for (int k = 0; k < 8; k++) all = map_then(all, s_wmap[k]);
        // ---- output offsets
        const uint32_t mine = words[tid][phase];
        uint32_t sum = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, sum, d);
            if (lane >= d) sum += o;
        }
        if (lane == 31) s_wsum[warp] = sum;
        __syncthreads();
        uint32_t wpre = 0, wall = 0;
        for (int k = 0; k < 8; k++) {
            if (k < warp) wpre += s_wsum[k];
            wall += s_wsum[k];
        }
        uint32_t o = done + wpre + sum - mine;
        // ---- the real parse: short runs are written here, long ones queued
        if (lo < wn) {
            int pos = lo + (int)phase;
            while (pos < hi) {
                int size;
                uint32_t c, v;
                parse_run<WIDE>(buf, pos, size, c, v);
                pos += size;
                if (o + c > expect || o + c < o) {  // would run past the caller's buffer
                    s_bad = 1;
                    c = o < expect ? expect - o : 0;
                }
                if (c <= LONG_RUN) {
                    for (uint32_t j = 0; j < c; j++) dst[o + j] = (T)v;
                } else {
                    queue[atomicAdd(&s_qn, 1u)] = make_uint2(o, ((c - 1) << 16) | v);
                }
                o += c;
            }
        }
        __syncthreads();
        const uint32_t qn = s_qn;
        for (uint32_t k = warp; k < qn; k += 8) {
            const uint2 e = queue[k];
            warp_fill<T>(dst + e.x, (e.y >> 16) + 1, e.y & 0xFFFFu, lane);
        }
        phase_in = map_apply(all, phase_in);
        done += wall;  // unclipped: a stream that decodes to too much is reported with its real size
    }
    __syncthreads();
    if (tid == 0) {
        // phase_in != 0: the last run's header reaches past the end of the stream
        const bool bad = s_bad || phase_in != 0 || done != expect;
        status[2 * blockIdx.x] = bad ? 1u : 0u;
        status[2 * blockIdx.x + 1] = done * (uint32_t)sizeof(T);
    }
}

// ---------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------
constexpr int EPT = 16;            // elements per thread
constexpr int TILE = 256 * EPT;    // elements per tile

template <bool WIDE>
__device__ __forceinline__ uint32_t enc_size(uint32_t counter, uint32_t c) {
    if (WIDE) return 1u + (counter >= 0x40u) + 1u + (c > 255u);  // coders/rle.cpp:171-183
    return 1u + (counter >= 0x80u) + 1u;                          // :118-125
}

template <bool WIDE>
__device__ __forceinline__ void enc_write(uint8_t* d, uint32_t counter, uint32_t c) {
    if (WIDE) {
        const uint32_t big = c > 255u ? 0x40u : 0u;
        if (counter >= 0x40u) {
            *d++ = (uint8_t)(0x80u | big | (counter & 0x3Fu));
            *d++ = (uint8_t)(counter >> 6);
        } else {
            *d++ = (uint8_t)(counter | big);
        }
        *d++ = (uint8_t)c;
        if (big) *d++ = (uint8_t)(c >> 8);
    } else {
        if (counter >= 0x80u) {
            *d++ = (uint8_t)(0x80u | (counter & 0x7Fu));
            *d++ = (uint8_t)(counter >> 7);
        } else {
            *d++ = (uint8_t)counter;
        }
        *d++ = (uint8_t)c;
    }
}

// grid (n_streams), 256 threads.  write == 0: sizes only.
template <bool WIDE>
__global__ void __launch_bounds__(256) k_rle_encode(const uint8_t* in, uint8_t* out,
                                                    const RleStream* streams, uint32_t* sizes,
                                                    int write) {
    using T = typename std::conditional<WIDE, uint16_t, uint8_t>::type;
    constexpr uint32_t MAXRUN = WIDE ? 0x4000u : 0x8000u;  // max_sequence + 1
    const RleStream S = streams[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* src = reinterpret_cast<const T*>(in + S.src);
    uint8_t* dst = out + S.dst;
    const uint32_t N = S.len / sizeof(T);
    __shared__ int s_wmax[8];
    __shared__ uint32_t s_wsum[8];
    int carry_start = 0;       // index of the last value change so far (uniform)
    uint32_t carry_bytes = 0;  // bytes written so far (uniform)
    for (uint32_t t0 = 0; t0 < N; t0 += TILE) {
        const uint32_t i0 = t0 + tid * EPT;
        uint32_t e[EPT + 1];
        uint32_t prev = 0;
        if (i0 + EPT <= N) {  // streams are 16-byte aligned and a multiple of 16 elements in practice
            const uint4* p4 = reinterpret_cast<const uint4*>(src + i0);
            if (WIDE) {
                const uint4 a = p4[0], b = p4[1];
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    e[2 * j] = w[j] & 0xFFFFu;
                    e[2 * j + 1] = w[j] >> 16;
                }
            } else {
                const uint4 a = p4[0];
                const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int j = 0; j < 16; j++) e[j] = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            }
        } else {
#pragma unroll
            for (int j = 0; j < EPT; j++) e[j] = i0 + j < N ? src[i0 + j] : 0u;
        }
        e[EPT] = i0 + EPT < N ? src[i0 + EPT] : 0u;
        if (i0 > 0 && i0 < N) prev = src[i0 - 1];
        // ---- last value change at or before each element: max-scan
        int last = -1;
#pragma unroll
        for (int j = 0; j < EPT; j++) {
            const uint32_t i = i0 + j;
            const uint32_t pv = j == 0 ? prev : e[j - 1];
            if (i < N && (i == 0 || e[j] != pv)) last = (int)i;
        }
        int inc = last;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc = max(inc, o);
        }
        __syncthreads();  // s_wmax / s_wsum of the previous tile are consumed
        if (lane == 31) s_wmax[warp] = inc;
        __syncthreads();
        int start = carry_start, tile_start = carry_start;
        for (int k = 0; k < 8; k++) {
            if (k < warp) start = max(start, s_wmax[k]);
            tile_start = max(tile_start, s_wmax[k]);
        }
        int exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
        if (lane > 0) start = max(start, exc);
        // ---- runs that end in this thread's elements and their encoded sizes
        uint32_t bytes = 0;
        {
            int rs = start;
#pragma unroll
            for (int j = 0; j < EPT; j++) {
                const uint32_t i = i0 + j;
                if (i >= N) break;
                const uint32_t pv = j == 0 ? prev : e[j - 1];
                if (i == 0 || e[j] != pv) rs = (int)i;
                const uint32_t counter = (i - (uint32_t)rs) % MAXRUN;
                const bool end = i == N - 1 || e[j + 1] != e[j] || counter == MAXRUN - 1;
                if (end) bytes += enc_size<WIDE>(counter, e[j]);
            }
        }
        uint32_t sum = bytes;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, sum, d);
            if (lane >= d) sum += o;
        }
        if (lane == 31) s_wsum[warp] = sum;
        __syncthreads();
        uint32_t wpre = 0, wall = 0;
        for (int k = 0; k < 8; k++) {
            if (k < warp) wpre += s_wsum[k];
            wall += s_wsum[k];
        }
        if (write) {
            uint8_t* d = dst + carry_bytes + wpre + sum - bytes;
            int rs = start;
#pragma unroll
            for (int j = 0; j < EPT; j++) {
                const uint32_t i = i0 + j;
                if (i >= N) break;
                const uint32_t pv = j == 0 ? prev : e[j - 1];
                if (i == 0 || e[j] != pv) rs = (int)i;
                const uint32_t counter = (i - (uint32_t)rs) % MAXRUN;
                const bool end = i == N - 1 || e[j + 1] != e[j] || counter == MAXRUN - 1;
                if (end) {
                    enc_write<WIDE>(d, counter, e[j]);
                    d += enc_size<WIDE>(counter, e[j]);
                }
            }
        }
        carry_start = tile_start;
        carry_bytes += wall;
    }
    if (tid == 0) sizes[blockIdx.x] = carry_bytes;
}

int grow_dev(vx_ctx* ctx, uint8_t** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t bytes = need + need / 4 + 256;
    VX_CUDA(cudaMalloc((void**)p, bytes));
    *cap = bytes;
    return 0;
}

inline size_t pad16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

extern "C" {

int vx_chunks_put_compressed(vx_ctx* ctx, const vx_chunk_compressed* chunks, int32_t n) {
    if (!ctx) return -1;
    if (ctx->pool_cap == 0) {
        ctx->fail("vx_chunks_put_compressed: configure the window first");
        return -1;
    }
    if (n <= 0) return 0;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    // ---- gather the blobs into one pinned buffer: voxel streams, then light streams
    std::vector<RleStream> sv(n), sl;
    std::vector<int> light_of(n, -1);
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        if (!chunks[i].voxels) {
            ctx->fail("vx_chunks_put_compressed: chunk without voxel data");
            return -1;
        }
        sv[i] = RleStream {total, (unsigned long long)i * ENC_CHUNK, chunks[i].voxels_len, VX_CHUNK_DATA_LEN};
        total += pad16(chunks[i].voxels_len);
    }
    for (int i = 0; i < n; i++) {
        if (!chunks[i].lights) continue;
        light_of[i] = (int)sl.size();
        sl.push_back(RleStream {total, (unsigned long long)i * ENC_CHUNK + VX_CHUNK_DATA_LEN, chunks[i].lights_len,
                                VX_LIGHTMAP_DATA_LEN});
        total += pad16(chunks[i].lights_len);
    }
    const size_t n_streams = sv.size() + sl.size();
    const size_t desc_bytes = pad16(n_streams * sizeof(RleStream));
    const size_t stat_bytes = pad16(n_streams * 2 * sizeof(uint32_t));
    if (!vx_host_stage(ctx, total + desc_bytes + stat_bytes)) return -1;
    if (grow_dev(ctx, &ctx->d_rle, &ctx->d_rle_bytes, total + desc_bytes + stat_bytes)) return -1;
    if (grow_dev(ctx, &ctx->d_codec_stage, &ctx->codec_stage_bytes, (size_t)n * ENC_CHUNK)) return -1;
    for (int i = 0; i < n; i++) {
        std::memcpy(ctx->h_rle + sv[i].src, chunks[i].voxels, chunks[i].voxels_len);
        if (light_of[i] >= 0) std::memcpy(ctx->h_rle + sl[light_of[i]].src, chunks[i].lights, chunks[i].lights_len);
    }
    RleStream* h_desc = reinterpret_cast<RleStream*>(ctx->h_rle + total);
    std::memcpy(h_desc, sv.data(), sv.size() * sizeof(RleStream));
    if (!sl.empty()) std::memcpy(h_desc + sv.size(), sl.data(), sl.size() * sizeof(RleStream));
    VX_CUDA(cudaMemcpyAsync(ctx->d_rle, ctx->h_rle, total + desc_bytes, cudaMemcpyHostToDevice, st));
    const RleStream* d_desc = reinterpret_cast<const RleStream*>(ctx->d_rle + total);
    uint32_t* d_stat = reinterpret_cast<uint32_t*>(ctx->d_rle + total + desc_bytes);
    uint32_t* h_stat = reinterpret_cast<uint32_t*>(ctx->h_rle + total + desc_bytes);
    // ---- decompress into the wire-format staging area (WorldRegions::getVoxels / getLights)
    VX_LAUNCH(ctx, KID_RLE_DECODE, k_rle_decode<true><<<(unsigned)sv.size(), 256, 0, st>>>(
                                       ctx->d_rle, ctx->d_codec_stage, d_desc, d_stat));
    if (!sl.empty())
        VX_LAUNCH(ctx, KID_RLE_DECODE, k_rle_decode<false><<<(unsigned)sl.size(), 256, 0, st>>>(
                                           ctx->d_rle, ctx->d_codec_stage, d_desc + sv.size(),
                                           d_stat + 2 * sv.size()));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(h_stat, d_stat, n_streams * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) {
        // compression::decompress throws for a voxel layer of the wrong size (coders/
        // compression.cpp:91-99); a truncated stream reads out of bounds there, an error here
        const int li = light_of[i];
        const bool bad_v = h_stat[2 * i] != 0;
        const bool bad_l = li >= 0 && h_stat[2 * (sv.size() + li)] != 0;
        if (bad_v || bad_l) {
            const uint32_t got = bad_v ? h_stat[2 * i + 1] : h_stat[2 * (sv.size() + li) + 1];
            ctx->fail("vx_chunks_put_compressed: chunk (" + std::to_string(chunks[i].cx) + ", " +
                      std::to_string(chunks[i].cz) + ") " + (bad_v ? "voxels" : "lights") +
                      ": expected decompressed size " +
                      std::to_string(bad_v ? VX_CHUNK_DATA_LEN : VX_LIGHTMAP_DATA_LEN) + " got " +
                      std::to_string(got));
            return -1;
        }
    }
    // ---- Chunk::decode / Lightmap::decode from the staging area
    std::vector<uint8_t> has_light(n);
    for (int i = 0; i < n; i++) {
        has_light[i] = light_of[i] >= 0;
        if (!has_light[i])
            VX_CUDA(cudaMemsetAsync(ctx->d_codec_stage + (size_t)i * ENC_CHUNK + VX_CHUNK_DATA_LEN, 0,
                                    VX_LIGHTMAP_DATA_LEN, st));
    }
    std::vector<vx_chunk_key> keys(n);
    for (int i = 0; i < n; i++) keys[i] = vx_chunk_key {chunks[i].cx, chunks[i].cz};
    return vx_put_staged(ctx, keys.data(), has_light.data(), n);
}

int vx_chunks_compress(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, int32_t layer, uint8_t* out,
                       uint64_t out_capacity, uint64_t* offsets) {
    if (!ctx) return -1;
    if (layer != VX_LAYER_VOXELS && layer != VX_LAYER_LIGHTS) {
        ctx->fail("vx_chunks_compress: layer must be VX_LAYER_VOXELS or VX_LAYER_LIGHTS");
        return -1;
    }
    if (!offsets) {
        ctx->fail("vx_chunks_compress: offsets is required");
        return -1;
    }
    offsets[0] = 0;
    if (n <= 0) return 0;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    cudaStream_t st = ctx->stream;
    const bool wide = layer == VX_LAYER_VOXELS;
    if (vx_encode_staged(ctx, keys, n, wide, !wide)) return -1;  // Chunk::encode / Lightmap::encode
    const uint32_t raw = wide ? VX_CHUNK_DATA_LEN : VX_LIGHTMAP_DATA_LEN;
    const size_t desc_bytes = pad16((size_t)n * sizeof(RleStream)), size_bytes = pad16((size_t)n * 4);
    if (!vx_host_stage(ctx, desc_bytes + size_bytes)) return -1;
    RleStream* h_desc = reinterpret_cast<RleStream*>(ctx->h_rle);
    uint32_t* h_sizes = reinterpret_cast<uint32_t*>(ctx->h_rle + desc_bytes);
    for (int i = 0; i < n; i++)
        h_desc[i] = RleStream {(unsigned long long)i * ENC_CHUNK + (wide ? 0 : VX_CHUNK_DATA_LEN), 0ull, raw, 0u};
    // descriptors and sizes live in front of the output bytes
    if (grow_dev(ctx, &ctx->d_rle, &ctx->d_rle_bytes, desc_bytes + size_bytes)) return -1;
    RleStream* d_desc = reinterpret_cast<RleStream*>(ctx->d_rle);
    uint32_t* d_sizes = reinterpret_cast<uint32_t*>(ctx->d_rle + desc_bytes);
    VX_CUDA(cudaMemcpyAsync(d_desc, h_desc, (size_t)n * sizeof(RleStream), cudaMemcpyHostToDevice, st));
    if (wide)
        VX_LAUNCH(ctx, KID_RLE_ENCODE, k_rle_encode<true><<<n, 256, 0, st>>>(ctx->d_codec_stage, nullptr, d_desc, d_sizes, 0));
    else
        VX_LAUNCH(ctx, KID_RLE_ENCODE, k_rle_encode<false><<<n, 256, 0, st>>>(ctx->d_codec_stage, nullptr, d_desc, d_sizes, 0));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(h_sizes, d_sizes, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    uint64_t total = 0;
    std::vector<uint32_t> sizes(h_sizes, h_sizes + n);  // h_rle may move below
    for (int i = 0; i < n; i++) {
        offsets[i] = total;
        total += sizes[i];
    }
    offsets[n] = total;
    if (!out || total > out_capacity) {
        ctx->fail("vx_chunks_compress: output needs " + std::to_string(total) + " bytes, capacity is " +
                  std::to_string(out ? out_capacity : 0));
        return -1;
    }
    const size_t head = desc_bytes + size_bytes;
    if (ctx->d_rle_bytes < head + total) {  // growing frees the buffer: descriptors are uploaded again below
        if (grow_dev(ctx, &ctx->d_rle, &ctx->d_rle_bytes, head + total)) return -1;
        d_desc = reinterpret_cast<RleStream*>(ctx->d_rle);
        d_sizes = reinterpret_cast<uint32_t*>(ctx->d_rle + desc_bytes);
    }
    if (!vx_host_stage(ctx, head)) return -1;
    h_desc = reinterpret_cast<RleStream*>(ctx->h_rle);
    for (int i = 0; i < n; i++)
        h_desc[i] = RleStream {(unsigned long long)i * ENC_CHUNK + (wide ? 0 : VX_CHUNK_DATA_LEN), head + offsets[i], raw, 0u};
    VX_CUDA(cudaMemcpyAsync(d_desc, h_desc, (size_t)n * sizeof(RleStream), cudaMemcpyHostToDevice, st));
    if (wide)
        VX_LAUNCH(ctx, KID_RLE_ENCODE, k_rle_encode<true><<<n, 256, 0, st>>>(ctx->d_codec_stage, ctx->d_rle, d_desc, d_sizes, 1));
    else
        VX_LAUNCH(ctx, KID_RLE_ENCODE, k_rle_encode<false><<<n, 256, 0, st>>>(ctx->d_codec_stage, ctx->d_rle, d_desc, d_sizes, 1));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(out, ctx->d_rle + head, total, cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
